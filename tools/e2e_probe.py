"""Where does the end-to-end time of one S(q) / RDF step go?  (GPU box, PYTHONPATH=.)"""
import time
import numpy as np
import torch
from mdhelper_b200 import _lib, synthetic
from mdhelper_b200.analysis.structure import StructureFactor, RadialDistributionFunction


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


u = synthetic.lj_fluid(50_000, 256, seed=1)
L = float(u.trajectory.unitcells[0, 0])
sf = StructureFactor([u.atoms], n_points=32, q_max=2 * np.pi * 16 / L, verbose=False,
                     batch_frames=128)
print("sq run(128 frames) ms:", t(lambda: sf.run(start=0, stop=128)))
ctx = _lib.Context(0)
N = 50_000
cfg = lambda: ctx.sq_configure(N, [0, N], sf._wavevectors, [(-1, -1)], lattice_n=sf._lattice_n,
                               lattice_b=sf._lattice_b, mode="auto")
print("sq configure ms:", t(cfg))
c = u.trajectory.coordinates
print("pinned:", torch.from_numpy(c).is_pinned() if hasattr(torch.from_numpy(c), "is_pinned") else None)
acc = lambda: (ctx.sq_accumulate(c.ctypes.data, 3 * N, 128), ctx.sync())
print("sq accumulate(host,128)+sync ms:", t(acc), "kernel ms", ctx.last_kernel_ms()[1])
dev = torch.from_numpy(c).cuda()
accd = lambda: (ctx.sq_accumulate(dev.data_ptr(), 3 * N, 128, device=True), ctx.sync())
print("sq accumulate(device,128)+sync ms:", t(accd), "kernel ms", ctx.last_kernel_ms()[1])
print("sq fetch ms:", t(lambda: ctx.sq_fetch()))
h = torch.from_numpy(c[:128])
d = torch.empty_like(h, device="cuda")
print("torch H2D 76.8MB ms:", t(lambda: d.copy_(h, non_blocking=True)))

u2, cat, an = synthetic.electrolyte(20_000, 200, seed=2)
rdf = RadialDistributionFunction(cat, an, n_bins=201, range=(0.0, 14.5), verbose=False,
                                 batch_frames=100)
print("rdf run(100 frames) ms:", t(lambda: rdf.run(start=0, stop=100)))
