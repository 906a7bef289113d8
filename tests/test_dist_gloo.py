"""CPU, world_size 2, gloo: frame sharding + the single all-reduce of the accumulators."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from _fake import FakeISF, FakeRDF, FakeSSF
        from mdhelper_b200.universe import SyntheticUniverse
        g = dict(np.load(GOLDEN / "rdf_lj1000.npz"))
        u = SyntheticUniverse(g["positions"], g["dims"])
        r = FakeRDF(u.atoms, n_bins=int(g["n_bins"]), range=tuple(g["range"]),
                    verbose=False).run()
        # 5 frames over 2 ranks -> 3 + 2, as np.array_split does
        assert r.n_local_frames == (3 if rank == 0 else 2)
        assert r.n_frames == 5
        assert np.array_equal(r.results.counts, g["counts"])
        np.testing.assert_allclose(r.results.rdf, g["rdf"], rtol=1e-6)
        # the in-run parity check of bench.py: a rank recomputes the whole frame list alone
        # (no sharding, no collective) and must find the all-reduced result
        from mdhelper_b200.analysis.base import single_rank, world
        if rank == 0:
            with single_rank():
                assert world() == (0, 1)
                alone = FakeRDF(u.atoms, n_bins=int(g["n_bins"]), range=tuple(g["range"]),
                                verbose=False).run()
                assert alone.n_local_frames == 5
            assert np.array_equal(alone.results.counts, r.results.counts)
        assert world() == (rank, 2)
        # the collective itself: exact int64 sums, float64 sums, the caller's array untouched
        from mdhelper_b200.analysis.base import all_reduce_sum
        mine = np.arange(6, dtype=np.int64).reshape(2, 3) * (rank + 1) + (1 << 40)
        keep = mine.copy()
        total = all_reduce_sum(mine)
        assert np.array_equal(mine, keep) and total.shape == (2, 3)
        assert np.array_equal(total, np.arange(6).reshape(2, 3) * 3 + (1 << 41))
        assert all_reduce_sum(np.array([0.25 * (rank + 1)]))[0] == 0.75
        # a ring trajectory: 12 frames cycling through the 5 in memory, split 6 + 6
        ur = SyntheticUniverse(g["positions"], g["dims"], n_frames=12)
        rr = FakeRDF(ur.atoms, n_bins=int(g["n_bins"]), range=tuple(g["range"]),
                     verbose=False).run()
        assert rr.n_local_frames == 6 and rr.n_frames == 12
        per_frame = [FakeRDF(u.atoms, n_bins=int(g["n_bins"]), range=tuple(g["range"]),
                             verbose=False) for _ in range(1)][0]
        with single_rank():
            want = sum(per_frame.run(start=f % 5, stop=f % 5 + 1).results.counts
                       for f in range(12))
        assert np.array_equal(rr.results.counts, want)
        # a strided selection and more ranks than frames on one side
        r2 = FakeRDF(u.atoms, n_bins=int(g["n_bins"]), range=tuple(g["range"]),
                     verbose=False).run(start=4)
        assert r2.n_local_frames == (1 if rank == 0 else 0)

        g = dict(np.load(GOLDEN / "sq_small.npz"))
        u = SyntheticUniverse(g["positions"], g["dims"])
        n = int(g["n_cat"])
        cat, an = u.select(slice(0, n)), u.select(slice(n, u.atoms.n_atoms))
        s = FakeSSF([cat, an], mode="partial", n_points=int(g["n_points"]),
                    q_max=float(g["q_max"]), verbose=False).run()
        assert s.n_local_frames == 1
        np.testing.assert_allclose(s.results.ssf, g["ssf_partial_exp"], rtol=1e-9,
                                   atol=1e-12)
        # intermediate scattering function: wavevector columns are sharded, every
        # rank streams all frames, one all-reduce assembles the columns
        g = dict(np.load(GOLDEN / "isf_small.npz"))
        u = SyntheticUniverse(g["positions"], g["dims"])
        n = int(g["n_cat"])
        cat, an = u.select(slice(0, n)), u.select(slice(n, u.atoms.n_atoms))
        f = FakeISF([cat, an], mode="partial", n_points=int(g["n_points"]),
                    q_max=float(g["q_max"]), n_lags=int(g["n_lags"]), incoherent=True,
                    dt=float(g["dt"]), verbose=False).run()
        n_q = len(f._wavenumbers)
        assert len(f._local_cols) == len(np.array_split(np.arange(n_q), 2)[rank])
        assert 0 < len(f._local_cols) < n_q and f.n_local_frames == 14
        np.testing.assert_allclose(f.results.cisf, g["cisf_partial_exp"], rtol=1e-9,
                                   atol=1e-10)
        np.testing.assert_allclose(f.results.iisf, g["iisf_partial_exp"], rtol=1e-9,
                                   atol=1e-10)
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), r.results.counts)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "ok0.npy")
    b = np.load(tmp_path / "ok1.npy")
    assert np.array_equal(a, b)          # every rank ends with the full result
