#!/usr/bin/env python
"""Fixed cost of one run() call (host side + synchronisation): runs over ONE frame and over
a few frame counts of the cfg4 / cfg2 / cfg3 systems, so that slope = per-frame cost and
intercept = per-call overhead.  Tuning aid for the strong-scaling passes of bench.py."""
import cProfile
import json
import pstats
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import torch
from mdhelper_b200 import synthetic
from mdhelper_b200.analysis.structure import RadialDistributionFunction, StructureFactor


def t(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


out = {}
u = synthetic.lj_fluid(50_000, 256, seed=1)
L = float(u.trajectory.unitcells[0, 0])
sf = StructureFactor([u.atoms], n_points=32, q_max=2 * np.pi * 16 / L, verbose=False)
for nf in (1, 8, 32, 128):
    out[f"sq_run_{nf}_frames_ms"] = t(lambda: sf.run(start=0, stop=nf))
u2, cat, an = synthetic.electrolyte(20_000, 64, seed=2)
rdf = RadialDistributionFunction(cat, an, n_bins=201, range=(0.0, 14.5), verbose=False)
for nf in (1, 8, 64):
    out[f"rdf_cfg2_run_{nf}_frames_ms"] = t(lambda: rdf.run(start=0, stop=nf))
u3 = synthetic.lj_fluid(500_000, 16, seed=3)
rdf3 = RadialDistributionFunction(u3.atoms, n_bins=100, range=(0.0, 2.5), verbose=False)
for nf in (1, 4, 16):
    out[f"rdf_cfg3_run_{nf}_frames_ms"] = t(lambda: rdf3.run(start=0, stop=nf))
print(json.dumps(out))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    sf.run(start=0, stop=1)
pr.disable()
pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(25)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    rdf.run(start=0, stop=1)
pr.disable()
pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(25)
