#!/usr/bin/env python
"""
bench.py -- hot-path benchmark (contract: one JSON line on stdout from rank 0).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W      # CPU reference arm

Workload (BASELINE.json configs[1], SURVEY.md section 8(d) "cfg2"): two-group
cation-anion partial RDF of a 20,000-ion electrolyte, n_bins=201, range (0, 14.5),
2,000 synthetic frames per GPU.  One "step" = one batch of ``--frames-per-step``
frames (default 200) through the pair-histogram hot path.

* ``value``      pairs binned / s, coordinates already resident in HBM, device time
                 (CUDA events on the launching stream), max over ranks.
* ``e2e``        the same metric through the public class
                 ``RadialDistributionFunction(cations, anions).run(start, stop)``
                 from pinned HOST memory (H2D of every frame and D2H of the
                 counts inside the timed region).
* ``roofline``   the kernel that runs by default is the fp32-filter pair kernel
                 (rdf_filter.cu: counts identical to the reference's fp64 arithmetic,
                 uncertain pairs re-evaluated in fp64): pair evaluations/s x 16 FP32
                 operations per evaluation over the measured packed-FP32 rate x SM
                 count x the SM clock sampled during the run; beside it
                 ``fp64_pipe_equivalent_frac``, the same rate against the bound of a
                 kernel that executes the reference's 21 FP64 instructions per pair.
                 ``--arith off`` benches that kernel (FP64-pipe roofline).  (Neither
                 HBM- nor tensor-bound; the HBM figure is reported for the record.)
* ``cpu_baseline`` the restated reference CPU path (oracle/: C distances + real
                 numpy.histogram), serial and frame-parallel over all host cores,
                 on a bounded sample of the same frames.
* ``secondary``  the S(q) half of the metric: frames/s of the direct-sum structure
                 factor for N=50,000, N_q=2,446 (configs[3]) with its own roofline
                 and CPU baseline.

Multi-GPU: frames shard over ranks (weak scaling: every rank processes its own
2,000-frame trajectory), no data-path collective; one NCCL all-reduce of the
int64 counts at the end, inside the timed region.
"""

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG2 = dict(n_ions=20_000, n_frames=2_000, n_bins=201, range=(0.0, 14.5), seed=20260002)
CFG4 = dict(n=50_000, n_frames=1_000, n_points=32, n_max=16, seed=20260004)
FP64_OPS_PER_PAIR = 21       # FP64-pipe instructions per pair evaluation
FP32_OPS_PER_PAIR = 16       # FP32 instructions per pair of the filter
                             # kernel (3 sub, 3 fma, 3 sub, 3 fma, mul + 2 fma, 1 fma)
FP64_OPS_PER_TERM = 4        # DFMA per (q, r) term of the lattice kernel


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=200)
    ap.add_argument("--sq-frames-per-step", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--hist", default="auto")
    ap.add_argument("--sq-kernel", default="lattice_dmma",
                    choices=["lattice_dmma", "lattice_fp64"],
                    help="lattice_dmma: FP64 matrix unit (default); lattice_fp64: scalar DFMA")
    ap.add_argument("--arith", default="auto", choices=["auto", "off"],
                    help="auto: fp32 filter + exact fp64 re-evaluation; off: fp64 for "
                         "every pair (the round-1 kernel)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_ready(self, timeout=30.0):
        """Blocks until nvidia-smi has printed its first sample: its start-up (NVML /
        driver initialisation, seconds on a fresh box) stalls kernel launches of other
        processes and must not overlap the timed region."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.05)

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm),
                "reasons": sorted(reasons)}


def profiled_traffic(kernel_tag, frames_in_capture, frames_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full
    capture (profiles/r01_<tag>_metrics.csv), scaled from the capture's frame count to
    this run's (the kernel streams every coordinate once, so traffic is linear in
    frames).  None if the profile is missing."""
    p = ROOT / "profiles" / f"r01_{kernel_tag}_metrics.csv"
    if not p.exists():
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    try:
        for line in p.read_text().splitlines():
            f = line.split(",")
            if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[2]) * unit[f[1]]
    except (ValueError, KeyError, IndexError):
        return None
    return tot * frames_per_launch / frames_in_capture


def measured_peaks():
    out = {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)",
           "fp64_per_clk_sm": 64.0, "sfu_per_clk_sm": 16.0, "fp32_per_clk_sm": 128.0,
           "fp32_source": "nominal (no profiles/microbench2_r01.json)",
           "pipe_source": "nominal (no profiles/microbench_r01.json)"}
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            out["hbm_gbs"] = float(json.loads(p.read_text())["hbm_gbs"])
            out["hbm_source"] = "MEASURED_PEAKS.json"
        except (ValueError, KeyError):
            pass
    p = ROOT / "profiles" / "microbench_r01.json"
    if p.exists():
        try:
            # GPU-wide instruction rates timed with CUDA events at the 1965 MHz the
            # run held (profiles/microbench_r01_clocks.csv) -> per clock and SM
            mb = json.loads(p.read_text())
            per = 1e9 / (mb["sms"] * mb["clock_khz"] * 1e3)
            out["fp64_per_clk_sm"] = float(mb["dfma"]["gops_per_s"]) * per
            out["sfu_per_clk_sm"] = float(mb["mufu_sin"]["gops_per_s"]) * per
            out["pipe_source"] = "measured: profiles/microbench_r01.json (tools/microbench.cu)"
        except (ValueError, KeyError):
            pass
    p = ROOT / "profiles" / "microbench2_r01.json"
    if p.exists():
        try:
            # FP32 lanes per clock and SM: packed FFMA2 (two lanes per instruction)
            mb = json.loads(p.read_text())
            per = 1e9 / (mb["sms"] * mb["clock_khz"] * 1e3)
            out["fp32_per_clk_sm"] = max(float(mb["ffma"]["gops_per_s"]),
                                         2 * float(mb["ffma2"]["gops_per_s"])) * per
            out["fp32_source"] = "measured: profiles/microbench2_r01.json (tools/microbench2.cu)"
        except (ValueError, KeyError):
            pass
    return out


# ---------------------------------------------------------------------------------
# CPU reference path (restated): oracle distances + real numpy.histogram
# ---------------------------------------------------------------------------------

_CPU_STATE = {}


def _cpu_rdf_frame(f):
    from oracle import reference_port as rp
    u, cat, an = _CPU_STATE["u"]
    ts = u.trajectory[int(f)]
    # mirrors _single_frame_parallel (structure.py:793-835): counts || volume
    c = rp.radial_histogram(cat.positions, an.positions, CFG2["n_bins"], CFG2["range"],
                            ts.dimensions, method="bruteforce")
    return np.concatenate((c, [ts.volume]))


def cpu_rdf(frames, n_jobs):
    """Pairs binned per second by the restated reference path on `n_jobs` processes."""
    import multiprocessing as mp
    t0 = time.perf_counter()
    if n_jobs == 1:
        res = [_cpu_rdf_frame(f) for f in frames]
    else:
        with mp.get_context("fork").Pool(n_jobs) as pool:      # base.py:477-501
            res = pool.map(_cpu_rdf_frame, frames, chunksize=1)
    dt = time.perf_counter() - t0
    tot = np.vstack(res).sum(axis=0)                           # structure.py:842
    return float(tot[:-1].sum()) / dt, dt, tot[:-1].astype(np.int64)


def cpu_sq(u, wavevectors, n_frames, n_threads):
    from oracle import reference_port as rp
    pos = [u.trajectory.coordinates[f].astype(np.float64) for f in range(n_frames)]
    rp.delta_fourier_transform_sum(wavevectors[:64], pos[0][:1000], n_threads)   # warm-up
    t0 = time.perf_counter()
    for p in pos:
        rho = rp.delta_fourier_transform_sum(wavevectors, p, n_threads)
        _ = (rho * rho.conj()).real
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def make_cpu_sample(n_frames):
    from mdhelper_b200 import synthetic
    _CPU_STATE["u"] = synthetic.electrolyte(CFG2["n_ions"], n_frames, seed=CFG2["seed"],
                                            pinned=False)


# ---------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = len(os.sched_getaffinity(0))
    fps = cores                                     # one frame per core per step
    n_need = fps * (args.steps + args.warmup)
    make_cpu_sample(min(n_need, 64))
    n_have = len(_CPU_STATE["u"][0].trajectory)
    for s in range(args.warmup):
        cpu_rdf([(s * fps + i) % n_have for i in range(fps)], cores)
    t0 = time.perf_counter()
    binned = 0
    for s in range(args.warmup, args.warmup + args.steps):
        _, _, c = cpu_rdf([(s * fps + i) % n_have for i in range(fps)], cores)
        binned += int(c.sum())
    dt = time.perf_counter() - t0
    value = binned / dt
    sample = (f"{fps} frames per step ({fps * args.steps} of the workload's 2,000), "
              f"multiprocessing fork pool of {cores} processes over frames "
              "(mirrors base.py:477-501)")
    line = {
        "impl": "reference", "metric": "rdf_pairs_binned_per_s", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(fps),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "restated reference CPU path: MDAnalysis is not installable here, so "
                "capped_distance is the C restatement in oracle/ feeding the real "
                "numpy.histogram",
    }
    emit(line)


def prewarm(step, ctx, seconds=0.75):
    """Untimed: run steps until `seconds` of wall time have passed, so the timed
    region starts with the SM clocks already raised."""
    t0 = time.perf_counter()
    s = 0
    while time.perf_counter() - t0 < seconds:
        step(s)
        ctx.sync()
        s += 1


def workload_config(frames_per_step):
    return {"workload": "cfg2: two-group cation-anion partial RDF, 20,000-ion electrolyte "
                        "(10,000 x 10,000 ordered pairs per frame), n_bins=201, "
                        "range=(0, 14.5), L=29.2402, 2,000 synthetic frames per GPU",
            "frames_per_step": frames_per_step,
            "pairs_per_frame": 100_000_000,
            "cache": "inputs larger than L2: 480 MB of coordinates cycle through HBM, "
                     "each step reads a different 24 MB batch",
            "parallelism": "frames sharded over GPUs, one all-reduce at the end"}


# ---------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------

def run_ours(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cores = len(os.sched_getaffinity(0))

    # ---- CPU baselines first (fork pool before any CUDA context exists) ----
    cpu = None
    cpu_sq_res = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        make_cpu_sample(2 * cores)
        v1, dt1, _ = cpu_rdf([0], 1)
        vp, dtp, _ = cpu_rdf(list(range(2 * cores)), cores)
        cpu = {"value": vp, "unit": "pairs/s", "cores": cores, "kind": "port",
               "serial_value": v1,
               "sample": f"serial: 1 frame in {dt1:.1f} s; parallel: {2 * cores} frames on a "
                         f"fork pool of {cores} processes in {dtp:.1f} s (same cfg2 frames; "
                         "restated reference path: C capped_distance + numpy.histogram)"}

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mdhelper_b200 import _lib, synthetic
    from mdhelper_b200.analysis._binning import squared_thresholds
    from mdhelper_b200.analysis.structure import RadialDistributionFunction, StructureFactor

    K, W, fps = args.steps, args.warmup, args.frames_per_step
    n_frames = CFG2["n_frames"]
    u, cat, an = synthetic.electrolyte(CFG2["n_ions"], n_frames, seed=CFG2["seed"] + 1000 * rank)
    n1, n2, N = cat.n_atoms, an.n_atoms, CFG2["n_ions"]
    coords = u.trajectory.coordinates
    boxes = np.ascontiguousarray(u.trajectory.unitcells[:, :3])

    # ---- value: coordinates resident in HBM ----
    dev = torch.from_numpy(coords).cuda(non_blocking=True)
    torch.cuda.synchronize()
    ctx = _lib.Context(local)
    thr = squared_thresholds(CFG2["n_bins"], CFG2["range"])
    ctx.rdf_set_filter(args.arith)
    ctx.rdf_configure(n1, n2, False, thr, *CFG2["range"], hist=args.hist)
    base = dev.data_ptr()

    def step(s):
        f0 = (s * fps) % n_frames
        nf = min(fps, n_frames - f0)
        ctx.rdf_accumulate(base + 4 * 3 * N * f0, 3 * N, base + 4 * 3 * (N * f0 + n1), 3 * N,
                           boxes[f0:f0 + nf], nf, device=True)
        return nf

    # the clock sampler starts first: nvidia-smi's own start-up (NVML init) must be
    # over before the timed region; only samples inside [wall0, wall1] are used
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    if world > 1:
        dist.barrier()
    prewarm(step, ctx, 1.5)                   # bring the clocks up (untimed, on top of W)
    for s in range(W):
        step(s)
    # the tail of the timed region once, untimed: the first fetch -> device copy ->
    # all-reduce of a process pays one-off initialisation (tens of ms on a fresh box)
    warm = torch.from_numpy(ctx.rdf_fetch()).cuda()
    if world > 1:
        dist.all_reduce(warm)
    torch.cuda.synchronize()
    ctx.sync()
    ctx.rdf_reset()
    ctx.kernel_time(reset=True)
    launches0 = ctx.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    ev0.record()
    frames_done, kernel_ms = 0, 0.0
    dbg = os.environ.get("MDH_BENCH_DEBUG")
    dbg_ev, dbg_host = [], []
    for s in range(W, W + K):
        if dbg:
            e = torch.cuda.Event(enable_timing=True); e.record(); dbg_ev.append(e)
            dbg_host.append(time.perf_counter())
        frames_done += step(s)
    if dbg:
        e = torch.cuda.Event(enable_timing=True); e.record(); dbg_ev.append(e)
        dbg_host.append(time.perf_counter())
    counts = torch.from_numpy(ctx.rdf_fetch()).cuda()
    if world > 1:
        dist.all_reduce(counts)
    ev1.record()
    torch.cuda.synchronize()
    wall1 = time.time()
    if dbg:
        print("per-step GPU ms:", [round(a.elapsed_time(b), 2) for a, b in
                                   zip(dbg_ev[:-1], dbg_ev[1:])], file=sys.stderr)
        print("per-step host enqueue ms:", [round(1e3 * (b - a), 2) for a, b in
                                            zip(dbg_host[:-1], dbg_host[1:])], file=sys.stderr)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = ctx.launch_count() - launches0
    binned_total = int(counts.sum().item())
    evals_local = ctx.rdf_pair_evaluations()
    fstats = ctx.rdf_filter_stats()
    filtered = bool(fstats["eligible"]) and args.arith != "off" and args.hist in ("auto", "warp_atomic")
    # per-launch duration of the pair kernel: CUDA events recorded around every
    # launch of the timed region, on the launching stream (mdh_kernel_time)
    kern_total_ms, kern_calls, _, _ = ctx.kernel_time(reset=True)
    kern_ms = kern_total_ms / max(kern_calls, 1)
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    value = binned_total / (ms * 1e-3)

    # ---- e2e: public class, pinned host memory, H2D + D2H inside the timed region ----
    rdf = RadialDistributionFunction(cat, an, n_bins=CFG2["n_bins"], range=CFG2["range"],
                                     verbose=False, batch_frames=fps, hist=args.hist,
                                     arith=args.arith)
    # weak scaling like `value`: run() shards the frames it is given over the ranks
    # (np.array_split), so a step hands it world * fps frames -- fps per GPU, each rank
    # reading its own trajectory -- and the counts come back summed over the ranks
    span = fps * world
    for s in range(W):
        f0 = (s * span) % n_frames
        rdf.run(start=f0, stop=min(n_frames, f0 + span))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_binned = 0
    for s in range(W, W + K):
        f0 = (s * span) % n_frames
        rdf.run(start=f0, stop=min(n_frames, f0 + span))
        e2e_binned += int(rdf.results.counts.sum())       # already summed over ranks
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = e2e_binned / float(e2e_s.item())

    # ---- secondary: S(q), cfg4 ----
    secondary = None
    if not args.no_secondary:
        secondary = bench_sq(args, rank, world, local, cores, dist, torch)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    sm_mhz = clocks["sm_mhz"] or clocks.get("sm_max_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    fp64_peak = peaks["fp64_per_clk_sm"] * n_sm * sm_mhz * 1e6        # instr/s
    fp32_peak = peaks["fp32_per_clk_sm"] * n_sm * sm_mhz * 1e6        # FP32 lanes/s
    evals_per_launch = fps * n1 * n2
    alg_bytes = fps * (n1 + n2) * 16 + CFG2["n_bins"] * 8            # float4 in, counts out
    hbm = {"achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9,
           "peak_gbs": peaks["hbm_gbs"], "source": peaks["hbm_source"],
           "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    fp64_equiv = evals_per_launch * FP64_OPS_PER_PAIR / (kern_ms * 1e-3)
    if filtered:
        # the kernel that runs: fp32 filter (packed FFMA2/FADD2 on the FP32 pipe);
        # about 1 pair in 1,500 is re-evaluated in fp64 (counted in fstats)
        achieved = evals_per_launch * FP32_OPS_PER_PAIR / (kern_ms * 1e-3)
        roofline = {
            "kernel": "rdf_filter_kernel", "bound": "fp32_pipe",
            "achieved": achieved / 1e9, "peak": fp32_peak / 1e9, "unit": "Ginstr/s",
            "frac": achieved / fp32_peak,
            "traffic": profiled_traffic("filter", 20, fps),
            "traffic_note": "dram__bytes_read+write of profiles/r01_filter_metrics.csv (a "
                            "20-frame launch) scaled to this launch's frames; bytes",
            "per_unit": f"{FP32_OPS_PER_PAIR} FP32 operations per pair evaluation "
                        "(minimum image, squared distance, bin coordinate; issued as "
                        "packed f32x2 instructions)",
            "units_per_launch": evals_per_launch, "launch_ms": kern_ms,
            "peak_source": f"{peaks['fp32_source']}; {peaks['fp32_per_clk_sm']:.1f} FP32 "
                           f"lanes/clk/SM x {n_sm} SMs x {sm_mhz:.0f} MHz (sampled)",
            "fp64_pipe_equivalent_frac": fp64_equiv / fp64_peak,
            "fp64_pipe_equivalent_note": "the same pair rate expressed against the bound "
                                         "of a kernel that evaluates the reference's 21 "
                                         "FP64 instructions for every pair (the --arith "
                                         "off kernel reaches 0.62 of it)",
            "filter": {"deferred_entries": fstats["deferred_entries"],
                       "inline_entries": fstats["inline_entries"],
                       "declined_frames": fstats["declined_frames"]},
            "hbm": hbm,
        }
    else:
        roofline = {
            "kernel": "rdf_allpairs_kernel", "bound": "fp64_pipe",
            "achieved": fp64_equiv / 1e9, "peak": fp64_peak / 1e9, "unit": "Ginstr/s",
            "frac": fp64_equiv / fp64_peak,
            "traffic": profiled_traffic("pair", 20, fps),
            "traffic_note": "dram__bytes_read+write of profiles/r01_pair_metrics.csv (a "
                            "20-frame launch) scaled to this launch's frames; bytes",
            "per_unit": f"{FP64_OPS_PER_PAIR} FP64-pipe instructions per pair evaluation "
                        "(no FMA fusion allowed)",
            "units_per_launch": evals_per_launch, "launch_ms": kern_ms,
            "peak_source": f"{peaks['pipe_source']}; {peaks['fp64_per_clk_sm']:.1f} "
                           f"instr/clk/SM x {n_sm} SMs x {sm_mhz:.0f} MHz (sampled)",
            "hbm": hbm,
        }
    line = {
        "metric": "rdf_pairs_binned_per_s", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(fps),
        "pairs_evaluated_per_s": evals_local * world / (ms * 1e-3) * (K / (K + 0.0)),
        "frames_per_s": frames_done * world / (ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pairs/s",
                "h2d_bytes_per_step": world * (fps * (n1 + n2) * 12 + fps * 48),
                "d2h_bytes_per_step": world * CFG2["n_bins"] * 8,
                "api": "RadialDistributionFunction(cations, anions, n_bins=201, "
                       "range=(0, 14.5)).run(start, stop) per step"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "secondary": secondary,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def bench_sq(args, rank, world, local, cores, dist, torch):
    """S(q) frames/s for cfg4 (N=50,000, N_q=2,446), device-resident and end to end."""
    from mdhelper_b200 import _lib, synthetic
    from mdhelper_b200.analysis.structure import StructureFactor
    K, W, fps = args.steps, args.warmup, args.sq_frames_per_step
    ring = max(256, fps * world)                  # distinct frames (>= 154 MB > L2)
    u = synthetic.lj_fluid(CFG4["n"], ring, seed=CFG4["seed"] + 1000 * rank)
    L = float(u.trajectory.unitcells[0, 0])
    q_max = 2 * np.pi * CFG4["n_max"] / L
    sf = StructureFactor([u.atoms], n_points=CFG4["n_points"], q_max=q_max, verbose=False,
                         batch_frames=fps, kernel=args.sq_kernel)
    dmma = args.sq_kernel == "lattice_dmma"
    n_q = len(sf._wavenumbers)
    N = CFG4["n"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v1, dt1 = cpu_sq(u, sf._wavevectors, 1, 1)
        vp, dtp = cpu_sq(u, sf._wavevectors, 3, cores)
        cpu = {"value": vp, "unit": "frames/s", "cores": cores, "kind": "port",
               "serial_value": v1,
               "sample": f"serial: 1 frame in {dt1:.1f} s; {cores} OpenMP threads over "
                         f"wavevectors: 3 frames in {dtp:.1f} s (C restatement of "
                         "accelerated.py:81-165)"}

    dev = torch.from_numpy(u.trajectory.coordinates).cuda()
    ctx = _lib.Context(local)
    ctx.sq_configure(N, [0, N], sf._wavevectors, [(-1, -1)], lattice_n=sf._lattice_n,
                     lattice_b=sf._lattice_b, mode=args.sq_kernel)
    assert ctx.sq_kernel() == args.sq_kernel, ctx.sq_kernel()
    tiling = ctx.sq_tiling()
    base = dev.data_ptr()

    def step(s):
        f0 = (s * fps) % ring
        nf = min(fps, ring - f0)
        ctx.sq_accumulate(base + 4 * 3 * N * f0, 3 * N, nf, device=True)
        return nf

    prewarm(step, ctx, 0.4)
    for s in range(W):
        step(s)
    warm = torch.from_numpy(ctx.sq_fetch()).cuda()
    if world > 1:
        dist.all_reduce(warm)
    torch.cuda.synchronize()
    ctx.sync()
    ctx.sq_reset()
    ctx.kernel_time(reset=True)
    l0 = ctx.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    frames = 0
    for s in range(W, W + K):
        frames += step(s)
    acc = torch.from_numpy(ctx.sq_fetch()).cuda()
    if world > 1:
        dist.all_reduce(acc)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = ctx.launch_count() - l0
    _, _, kern_total_ms, kern_calls = ctx.kernel_time(reset=True)
    kern_ms = kern_total_ms / max(kern_calls, 1) * (fps * K / max(frames, 1))

    span = fps * world                            # fps frames per GPU and step
    for s in range(W):
        f0 = (s * span) % ring
        sf.run(start=f0, stop=min(ring, f0 + span))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_frames = 0
    for s in range(W, W + K):
        f0 = (s * span) % ring
        sf.run(start=f0, stop=min(ring, f0 + span))
        e2e_frames += sf.n_frames                 # all ranks together
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)

    peaks = measured_peaks()
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    terms = fps * N * n_q
    achieved = terms * FP64_OPS_PER_TERM / (kern_ms * 1e-3)
    peak = peaks["fp64_per_clk_sm"] * n_sm * 1965e6
    return {
        "metric": "sq_frames_per_s", "value": frames * world / (ms * 1e-3), "unit": "frames/s",
        "config": {"workload": f"cfg4: direct-sum S(q), N=50,000, n_points=32, "
                               f"q_max=2*pi*16/L -> N_q={n_q}, mode=None, form=exp, fp64",
                   "frames_per_step": fps, "terms_per_frame": N * n_q,
                   "cache": f"ring of {ring} distinct frames ({ring * N * 12 / 1e6:.0f} MB > L2)"},
        "ms_per_step": ms / K, "gpu_launches": launches,
        "e2e": {"value": e2e_frames / float(e2e_s.item()), "unit": "frames/s",
                "h2d_bytes_per_step": world * fps * N * 12,
                "d2h_bytes_per_step": world * n_q * 8,
                "api": "StructureFactor([atoms], n_points=32, q_max=...).run(start, stop)"},
        # the DMMA kernel runs on the FP64 matrix unit: the contract's "tensor" bound, in
        # TFLOP/s (2 flop per FMA); the scalar kernel is bound by the FP64 FMA pipe
        "roofline": {"kernel": "sq_lattice_mma_kernel" if dmma else "sq_lattice_kernel<double>",
                     "bound": "tensor" if dmma else "fp64_pipe",
                     "achieved": achieved * 2 / 1e12 if dmma else achieved / 1e9,
                     "peak": peak * 2 / 1e12 if dmma else peak / 1e9,
                     "unit": "TFLOP/s" if dmma else "Ginstr/s",
                     "frac": achieved / peak,
                     "bound_note": ("fp64 matrix unit (mma.m8n8k4.f64 = SASS DMMA.8x8x4); "
                                    "algorithmic work = 4 fp64 FMA = 8 flop per (q, r) term")
                     if dmma else None,
                     "traffic": profiled_traffic("sq", 128, fps),
                     "traffic_note": "dram__bytes_read+write of profiles/r01_sq_metrics.csv "
                                     "(a 128-frame launch) scaled to this launch's frames; "
                                     "bytes",
                     "nominal_vs_reachable":
                         ("DMMA.8x8x4 sustains the nominal 63.6 FMA/clk/SM "
                          "(profiles/microbench3_r01.json); the (column group x nz tile) "
                          "tiling of the |q| <= q_max sphere carries padding slots that "
                          "are executed but not counted as algorithmic work") if dmma else
                         ("a DFMA with three register operands sustains 42.6/clk/SM "
                          "(profiles/microbench2_r01.json), 0.67 of the nominal rate used "
                          "as peak"),
                     "per_unit": f"{FP64_OPS_PER_TERM} fp64 FMA per (q, r) term (complex "
                                 "multiply-accumulate)" + (", issued as DMMA m8n8k4 "
                                 "(256 FMA per warp instruction)" if dmma else ""),
                     "units_per_launch": terms, "launch_ms": kern_ms,
                     "tiling": (dict(tiling, executed_over_algorithmic=tiling["tiles"] * 64 / n_q,
                                     note="(column group x nz tile) pairs of 64 accumulator "
                                          "slots; the DMMAs of padding slots are executed but "
                                          "not counted in `achieved`") if dmma else None),
                     "peak_source": f"{peaks['pipe_source']}; {peaks['fp64_per_clk_sm']:.1f} "
                                    f"instr/clk/SM x {n_sm} SMs x 1965 MHz (max clock)",
                     "sfu_equivalent_frac": (terms * 2 / (kern_ms * 1e-3))
                     / (peaks["sfu_per_clk_sm"] * n_sm * 1965e6)},
        "cpu_baseline": cpu,
    }


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    # Libraries print to stdout behind our back (NCCL: "NCCL version ..." at communicator
    # creation): everything but the result line goes to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
