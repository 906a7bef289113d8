#!/usr/bin/env python
"""Device-resident S(q) rate of cfg4 (N = 50,000, n_max = 16) for kernel experiments:
MDH_B200_LIB=<variant .so> python tools/sq_speed.py [kernel] [frames_per_step] [n_max]"""
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import torch
from mdhelper_b200 import _lib, synthetic
from mdhelper_b200.analysis.structure import StructureFactor

kernel = sys.argv[1] if len(sys.argv) > 1 else "lattice_dmma"
fps = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n_max = int(sys.argv[3]) if len(sys.argv) > 3 else 16
N = 50_000
u = synthetic.lj_fluid(N, 256, seed=20260004)
L = float(u.trajectory.unitcells[0, 0])
kw = dict(n_points=32, verbose=False, batch_frames=fps, kernel=kernel)
if n_max < 32:
    kw["q_max"] = 2 * np.pi * n_max / L
sf = StructureFactor([u.atoms], **kw)
n_q = len(sf._wavenumbers)
dev = torch.from_numpy(u.trajectory.coordinates).cuda()
ctx = _lib.Context(0)
ctx.sq_configure(N, [0, N], sf._wavevectors, [(-1, -1)], lattice_n=sf._lattice_n,
                 lattice_b=sf._lattice_b, mode=kernel)
base = dev.data_ptr()


def step(s):
    f0 = (s * fps) % 256
    nf = min(fps, 256 - f0)
    ctx.sq_accumulate(base + 12 * N * f0, 3 * N, nf, device=True)
    return nf


t0 = time.time()
while time.time() - t0 < 0.5:
    step(0)
ctx.sync()
ctx.kernel_time(reset=True)
frames = 0
for s in range(10):
    frames += step(s)
ctx.sync()
_, _, ms, calls = ctx.kernel_time(reset=True)
rate = frames / (ms * 1e-3)
terms = N * n_q
print(f"{_lib.LIB_PATH.name} {ctx.sq_kernel()} n_q={n_q} fps={fps}: {rate:.0f} frames/s, "
      f"{ms / calls:.3f} ms/launch, frac of 63.7 FMA/clk/SM = "
      f"{rate * terms * 4 / (63.7 * 148 * 1965e6):.3f}  tiling={ctx.sq_tiling()}")
