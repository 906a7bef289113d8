"""Max / median relative error of S(q) of one cfg4 frame (N = 50,000, N_q = 2,446) for each
GPU kernel against the output of the reference's numba kernel (tests/golden/sq_cfg4_frame.npz)."""
import sys; sys.path.insert(0, ".")
import numpy as np
from mdhelper_b200 import synthetic
from mdhelper_b200.analysis.structure import StructureFactor
g = np.load("tests/golden/sq_cfg4_frame.npz")
u = synthetic.lj_fluid(int(g["n"]), 1, seed=int(g["seed"]))
for k in ("lattice_dmma", "lattice_fp64", "general_fp64"):
    s = StructureFactor([u.atoms], n_points=32, q_max=float(g["q_max"]), sort=False, unique=False,
                        verbose=False, kernel=k).run()
    rel = np.abs(s.results.ssf - g["ssf_raw"]) / g["ssf_raw"]
    print(k, "max rel err vs the reference's numba output: %.2e  median %.2e" % (rel.max(), np.median(rel)))
