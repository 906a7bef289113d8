#!/usr/bin/env python
"""
One pass over the five BASELINE.json configurations on one GPU (bounded frame
counts; SURVEY.md section 8(d) sizes), through the public classes.  Prints one JSON
object per configuration: end-to-end rates (pinned host -> results) and the device time
of the hot kernels (CUDA events, mdh_kernel_time).  Used for profiles/configs_rNN.json.
"""
import json
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from mdhelper_b200 import synthetic  # noqa: E402
from mdhelper_b200.analysis.structure import (IntermediateScatteringFunction,  # noqa: E402
                                              RadialDistributionFunction,
                                              StructureFactor)


def timed(obj, **run_kw):
    obj.run(**run_kw)                      # warm-up (allocations, clocks)
    obj._ctx.kernel_time(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    obj.run(**run_kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rdf_ms, rdf_n, sq_ms, sq_n = obj._ctx.kernel_time(reset=True)
    return dt, rdf_ms, sq_ms


def rdf_report(name, rdf, n_pairs_frame, **run_kw):
    dt, kms, _ = timed(rdf, **run_kw)
    ev0 = rdf._pair_evaluations
    out = {
        "config": name, "frames": rdf.n_frames, "e2e_s": dt,
        "e2e_frames_per_s": rdf.n_frames / dt,
        "e2e_pairs_binned_per_s": float(rdf.results.counts.sum()) / dt,
        "kernel_ms": kms,
        "pair_evaluations": ev0,
        "kernel_evaluations_per_s": ev0 / (kms * 1e-3) if kms else None,
        "ordered_pairs_considered_per_s_kernel":
            n_pairs_frame * rdf.n_frames / (kms * 1e-3) if kms else None,
        "fp64_pipe_frac": (ev0 * 21 / (kms * 1e-3)) / 18529.6e9 if kms else None,
        "kernel_us_per_frame": 1e3 * kms / rdf.n_frames if kms else None,
        "filter_stats": rdf._filter_stats,
        "counts_sum": int(rdf.results.counts.sum()),
    }
    print(json.dumps(out), flush=True)


def main():
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    if "cfg1" in which:
        u = synthetic.lj_fluid(1000, 200, seed=20260001)
        rdf = RadialDistributionFunction(u.atoms, n_bins=201, range=(0.0, 5.375),
                                         verbose=False)
        rdf_report("cfg1: RDF 1,000-particle LJ, 200 frames", rdf, 1000 * 1000)
    if "cfg2" in which:
        u, cat, an = synthetic.electrolyte(20_000, 400, seed=20260002)
        rdf = RadialDistributionFunction(cat, an, n_bins=201, range=(0.0, 14.5),
                                         verbose=False, batch_frames=100)
        rdf_report("cfg2: cation-anion RDF, 20k ions, 400 of 2,000 frames", rdf, 10 ** 8)
    if "cfg3" in which:
        u = synthetic.lj_fluid(500_000, 64, seed=20260003)
        for mode in ("cells",):
            rdf = RadialDistributionFunction(u.atoms, n_bins=100, range=(0.0, 2.5),
                                             verbose=False, mode=mode, batch_frames=32)
            rdf_report(f"cfg3: RDF cut-off 2.5, 500k LJ, 64 frames, mode={mode}", rdf,
                       500_000 ** 2)
    if "cfg4" in which:
        u = synthetic.lj_fluid(50_000, 256, seed=20260004)
        L = float(u.dimensions[0])
        sf = StructureFactor([u.atoms], n_points=32, q_max=2 * np.pi * 16 / L,
                             verbose=False, batch_frames=128)
        dt, _, sms = timed(sf)
        nq = len(sf._wavenumbers)
        print(json.dumps({"config": "cfg4: S(q) 50k particles, n_max=16, 256 frames",
                          "n_q": nq, "e2e_frames_per_s": sf.n_frames / dt,
                          "kernel_ms": sms, "kernel_frames_per_s": sf.n_frames / (sms * 1e-3),
                          "fp64_pipe_frac": 50_000 * nq * sf.n_frames * 4 / (sms * 1e-3)
                          / 18529.6e9}), flush=True)
    if "cfg5" in which:
        u = synthetic.polymer_melt(10_000, 100, 16, seed=20260005)
        rdf = RadialDistributionFunction(u.atoms, n_bins=100, range=(0.0, 2.5),
                                         verbose=False, batch_frames=8)
        rdf_report("cfg5a: RDF cut-off 2.5, 1M-bead melt, 16 frames", rdf, 10 ** 12)
        L = float(u.dimensions[0])
        sf = StructureFactor([u.atoms], n_points=32, q_max=2 * np.pi * 16 / L,
                             verbose=False, batch_frames=8)
        dt, _, sms = timed(sf)
        nq = len(sf._wavenumbers)
        print(json.dumps({"config": "cfg5b: S(q) 1M beads, n_max=16, 16 frames", "n_q": nq,
                          "e2e_frames_per_s": sf.n_frames / dt, "kernel_ms": sms,
                          "kernel_frames_per_s": sf.n_frames / (sms * 1e-3),
                          "fp64_pipe_frac": 1e6 * nq * sf.n_frames * 4 / (sms * 1e-3)
                          / 18529.6e9}), flush=True)
        # the configuration as BASELINE.json names it: RDF + S(q) in ONE pass over each
        # uploaded frame
        from mdhelper_b200.analysis import CombinedAnalysis
        both = CombinedAnalysis(rdf, sf, batch_frames=8)
        both.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        both.run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"config": "cfg5: combined RDF + S(q) pass, 1M beads, 16 frames",
                          "e2e_frames_per_s": rdf.n_frames / dt, "e2e_s": dt}), flush=True)


def isf():
    """Scope table 8(f) rank 1: F(q, t) and F_s(q, t) of the cfg4 system, 96 frames, 32
    lags -- 96 rho(q, t) sums and 2,576 displacement sums of 50,000 x 2,446 terms."""
    u = synthetic.lj_fluid(50_000, 96, seed=20260011)
    L = float(u.dimensions[0])
    kw = dict(n_points=32, q_max=2 * np.pi * 16 / L, n_lags=32, verbose=False,
              batch_frames=32, kernel=os.environ.get("MDH_BENCH_SQ_KERNEL"))
    for inc in (False, True):
        f = IntermediateScatteringFunction([u.atoms], incoherent=inc, **kw)
        dt, _, sms = timed(f)
        nq = len(f._wavenumbers)
        sums = 96 + (sum(min(32, t + 1) for t in range(96)) if inc else 0)
        print(json.dumps({"config": f"isf: 50k particles, 96 frames, 32 lags, "
                                    f"incoherent={inc}", "kernel": f._ctx.sq_kernel(),
                          "n_q": nq, "direct_sums": sums,
                          "e2e_s": dt, "kernel_ms": sms,
                          "sums_per_s_kernel": sums / (sms * 1e-3),
                          "fp64_pipe_frac": 50_000 * nq * sums * 4 / (sms * 1e-3)
                          / 18529.6e9}), flush=True)


def scsf():
    """Scope table 8(f) rank 3: single-chain structure factor of the cfg5 melt (10,000
    chains x 100 beads), full 32^3 wavevector grid, 2 frames."""
    from mdhelper_b200.analysis.polymer import SingleChainStructureFactor
    u = synthetic.polymer_melt(10_000, 100, 2, seed=20260005)
    s = SingleChainStructureFactor(u.atoms, n_points=32, n_chains=10_000, n_monomers=100,
                                   verbose=False, batch_frames=2,
                                   kernel=os.environ.get("MDH_BENCH_SQ_KERNEL"))
    dt, _, sms = timed(s)
    nq = len(s._wavenumbers)
    print(json.dumps({"config": "scsf: 10,000 chains x 100 beads, 32^3 wavevectors, 2 frames",
                      "kernel": s._ctx.sq_kernel(), "n_q": nq, "e2e_s": dt, "e2e_frames_per_s": s.n_frames / dt,
                      "kernel_ms": sms, "terms_per_s_kernel": 1e6 * nq * s.n_frames
                      / (sms * 1e-3),
                      "fp64_pipe_frac": 1e6 * nq * s.n_frames * 4 / (sms * 1e-3)
                      / 18529.6e9}), flush=True)


if __name__ == "__main__":
    if "scsf" in sys.argv[1:]:
        scsf()
    elif "isf" in sys.argv[1:]:
        isf()
    else:
        main()
