import cProfile, pstats, io
import numpy as np, torch
from mdhelper_b200 import synthetic
from mdhelper_b200.analysis.structure import StructureFactor, RadialDistributionFunction
u2, cat, an = synthetic.electrolyte(20_000, 200, seed=2)
rdf = RadialDistributionFunction(cat, an, n_bins=201, range=(0.0, 14.5), verbose=False, batch_frames=100)
rdf.run(start=0, stop=100); rdf.run(start=0, stop=100)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): rdf.run(start=0, stop=100)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3800])
u = synthetic.lj_fluid(50_000, 128, seed=1)
L = float(u.trajectory.unitcells[0, 0])
sf = StructureFactor([u.atoms], n_points=32, q_max=2 * np.pi * 16 / L, verbose=False, batch_frames=128)
sf.run(start=0, stop=128); sf.run(start=0, stop=128)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): sf.run(start=0, stop=128)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3200])
