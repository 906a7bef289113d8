#!/usr/bin/env python
"""Cell-list pair kernel on device-resident frames (tuning aid): us per frame of the whole
cell-list pipeline (CUDA events of mdh_kernel_time) for the cfg3 fluid and the cfg5 melt.
MDH_TUNE="cdbg=1" adds the per-stage times the library prints itself."""
import json
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from mdhelper_b200 import _lib, synthetic  # noqa: E402
from mdhelper_b200.analysis._binning import squared_thresholds  # noqa: E402


def run(name, u, frames_per_call, calls):
    N = u.atoms.n_atoms
    coords = u.trajectory.coordinates
    F = coords.shape[0]
    dev = torch.from_numpy(coords).cuda()
    boxes = np.ascontiguousarray(u.trajectory.unitcells[:, :3])
    ctx = _lib.Context(0)
    ctx.rdf_configure(N, N, True, squared_thresholds(100, (0.0, 2.5)), 0.0, 2.5, mode="cells")

    def call(k):
        f0 = (k * frames_per_call) % (F - frames_per_call + 1)
        ctx.rdf_accumulate(dev.data_ptr() + 12 * N * f0, 3 * N, None, 0,
                           boxes[f0:f0 + frames_per_call], frames_per_call, device=True)
    for k in range(2):
        call(k)
    ctx.sync()
    ctx.rdf_reset()
    ctx.kernel_time(reset=True)
    t0 = time.perf_counter()
    for k in range(calls):
        call(k)
    ctx.sync()
    wall = time.perf_counter() - t0
    ms, n, _, _ = ctx.kernel_time(reset=True)
    frames = frames_per_call * calls
    ev = ctx.rdf_pair_evaluations()
    st = ctx.rdf_filter_stats()
    print(json.dumps({"config": name, "tune": os.environ.get("MDH_TUNE", ""),
                      "frames": frames, "kernel_us_per_frame": 1e3 * ms / frames,
                      "wall_us_per_frame": 1e6 * wall / frames,
                      "pair_evaluations_per_frame": ev / frames,
                      "evaluations_per_s": ev / (ms * 1e-3),
                      "canonical_half_stencil_evals_per_s":
                          N * (N / float(np.prod(boxes[0], dtype=np.float64))) * 27 * 2.5 ** 3 / 2
                          / (1e-3 * ms / frames),
                      "deferred_per_frame": st["deferred_entries"] / frames,
                      "inline_per_frame": st["inline_entries"] / frames}), flush=True)
    ctx.close()
    del dev


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg5"]
    if "cfg3" in which:
        run("cfg3 fluid 500k", synthetic.lj_fluid(500_000, 32, seed=20260003, pinned=False), 16, 8)
    if "cfg5" in which:
        run("cfg5 melt 1M", synthetic.polymer_melt(10_000, 100, 8, seed=20260005, pinned=False),
            8, 6)
