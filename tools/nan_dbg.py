import numpy as np
from mdhelper_b200.analysis import structure as S
rng = np.random.default_rng(3)
dims = np.array([6.0, 6.0, 6.0, 90, 90, 90], np.float32)
p = (rng.random((700, 3)) * 6).astype(np.float32)
for name, idx, val in [("nan5", 5, np.nan), ("inf5", 5, np.inf), ("nan300", 300, np.nan), ("big", 5, 1e30)]:
    q = p.copy(); q[idx] = val
    st = {}
    got = S.radial_histogram(q, q, 50, (0.0, 3.0), dims, stats=st)
    print(name, st["declined_frames"], st["deferred_entries"], got.sum())
