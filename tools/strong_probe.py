#!/usr/bin/env python
"""Where a strong-scaling pass of cfg4 spends its time on every rank (tuning aid): host
time of the stages of StructureFactor.run(), device time of the kernels, the all-reduce.
Launch like bench.py (torchrun for N > 1)."""
import json
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from mdhelper_b200 import synthetic  # noqa: E402
from mdhelper_b200.analysis import structure  # noqa: E402
from mdhelper_b200.universe import SyntheticUniverse  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    C = bench.CFG4
    pos, L, keep = synthetic.fluid_positions(C["n"], 250, seed=C["seed"])
    u = SyntheticUniverse(pos, np.array([L, L, L, 90, 90, 90], np.float32),
                          n_frames=int(os.environ.get("PROBE_FRAMES", C["n_frames"])))
    sf = structure.StructureFactor([u.atoms], n_points=C["n_points"],
                                   q_max=2 * np.pi * C["n_max"] / float(L), verbose=False)
    marks = {}

    def timed(obj, name):
        fn = getattr(obj, name)

        def wrapper(*a, **k):
            t = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                marks[name] = marks.get(name, 0.0) + time.perf_counter() - t
        setattr(obj, name, wrapper)

    for name in ("_setup_frames", "_prepare", "_begin", "_consume", "_finish", "_conclude"):
        timed(sf, name)
    timed(structure, "all_reduce_sum")

    def one_pass():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sf.run()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    one_pass()
    one_pass()
    marks.clear()
    sf._ctx.kernel_time(reset=True)
    reps = 10
    times = [one_pass() for _ in range(reps)]
    out = {"rank": rank, "world": world, "pass_ms": [round(1e3 * t, 3) for t in times],
           "mean_ms": 1e3 * float(np.mean(times)),
           "kernel_ms": sf._ctx.kernel_time()[2] / reps}
    out.update({k + "_ms": 1e3 * v / reps for k, v in marks.items()})
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, out)
    else:
        gathered = [out]
    if rank == 0:
        for g in gathered:
            print(json.dumps(g), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
