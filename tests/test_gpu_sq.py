"""GPU: structure-factor kernels through the C ABI vs the oracle / golden fixtures
(tolerance: 1e-6 relative per BASELINE.json; the fp64 kernels are held to 1e-9)."""
import hashlib

import numpy as np
import pytest

from conftest import universe_from

pytestmark = pytest.mark.gpu


def _S():
    from mdhelper_b200.analysis import structure
    return structure


def _groups(u, g):
    n = int(g["n_cat"])
    return u.select(slice(0, n)), u.select(slice(n, u.atoms.n_atoms))


@pytest.mark.parametrize("kernel", [None, "lattice_dmma", "lattice_fp64", "general_fp64"])
@pytest.mark.parametrize("mode", [None, "pair", "partial"])
def test_class_matches_golden(golden, mode, kernel):
    g = golden("sq_small")
    u = universe_from(g)
    cat, an = _groups(u, g)
    for form in ("exp", "trig"):
        s = _S().StructureFactor([cat, an], mode=mode, form=form,
                                 n_points=int(g["n_points"]), q_max=float(g["q_max"]),
                                 kernel=kernel, verbose=False).run()
        np.testing.assert_allclose(s.results.ssf, g[f"ssf_{mode}_{form}"], rtol=1e-9,
                                   atol=1e-12)
        np.testing.assert_allclose(s.results.wavenumbers,
                                   g[f"wavenumbers_{mode}_{form}"], rtol=1e-13)
    assert len(s.results.pairs) == (3 if mode == "partial" else 1)
    # default: a lattice kernel (the matrix-unit one for all but tiny wavevector sets)
    assert s._ctx.sq_kernel() == kernel or (kernel is None and s._ctx.sq_kernel() in
                                            ("lattice_dmma", "lattice_fp64"))


def test_raw_grid_order_and_rho(golden):
    """sort=False, unique=False exposes the meshgrid ordering (Appendix B); rho(q)
    of the last frame is seam #2 itself."""
    g = golden("sq_small")
    u = universe_from(g)
    s = _S().StructureFactor([u.atoms], n_points=int(g["n_points"]),
                             q_max=float(g["q_max"]), sort=False, unique=False,
                             verbose=False).run()
    assert np.array_equal(s._wavevectors, g["wavevectors_raw"])
    np.testing.assert_allclose(s.results.ssf, g["ssf_raw"], rtol=1e-9, atol=1e-12)
    rho = s._ctx.sq_fetch_rho()[0]
    np.testing.assert_allclose(rho, g["rho_last"], rtol=0, atol=1e-9)
    # S(q = 0) = N exactly
    assert s.results.ssf[0, 0] == pytest.approx(u.atoms.n_atoms, rel=1e-14)


def test_off_lattice_wavevectors(golden):
    g = golden("sq_small")
    u = universe_from(g)
    cat, an = _groups(u, g)
    s = _S().StructureFactor([u.atoms], n_points=6, n_surfaces=3, n_surface_points=8,
                             verbose=False).run()
    np.testing.assert_allclose(s.results.ssf, g["ssf_surfaces"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(s.results.wavenumbers, g["wavenumbers_surfaces"],
                               rtol=1e-13)
    s = _S().StructureFactor([cat, an], mode="partial",
                             wavevectors=g["wavevectors_user"], sort=False, unique=False,
                             verbose=False).run()
    np.testing.assert_allclose(s.results.ssf, g["ssf_user"], rtol=1e-9, atol=1e-12)


def test_noncubic(golden):
    g = golden("sq_noncubic")
    u = universe_from(g)
    for kernel in (None, "lattice_dmma", "lattice_fp64", "general_fp64"):
        s = _S().StructureFactor([u.atoms], n_points=int(g["n_points"]),
                                 q_max=float(g["q_max"]), kernel=kernel,
                                 verbose=False).run()
        np.testing.assert_allclose(s.results.ssf, g["ssf"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(s.results.wavenumbers, g["wavenumbers"], rtol=1e-13)


def test_config4_frame(golden):
    """N = 50,000, n_max = 16 (N_q = 2,446): one frame of the bench workload against
    the reference's numba kernel output."""
    from mdhelper_b200 import synthetic
    g = golden("sq_cfg4_frame")
    u = synthetic.lj_fluid(int(g["n"]), 1, seed=int(g["seed"]))
    sha = hashlib.sha256(np.ascontiguousarray(u.trajectory.coordinates).tobytes())
    assert sha.hexdigest() == str(g["positions_sha256"]), "synthetic generator drifted"
    s = _S().StructureFactor([u.atoms], n_points=32, q_max=float(g["q_max"]),
                             sort=False, unique=False, verbose=False).run()
    assert s.results.ssf.shape == (1, 2446)
    assert s._ctx.sq_kernel() == "lattice_dmma"
    np.testing.assert_allclose(s.results.ssf, g["ssf_raw"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(s.results.ssf, g["ssf_raw"], rtol=1e-9, atol=1e-12)
    sd = _S().StructureFactor([u.atoms], n_points=32, q_max=float(g["q_max"]),
                              sort=False, unique=False, kernel="lattice_fp64",
                              verbose=False).run()
    assert sd._ctx.sq_kernel() == "lattice_fp64"
    np.testing.assert_allclose(sd.results.ssf, g["ssf_raw"], rtol=1e-9, atol=1e-12)
    su = _S().StructureFactor([u.atoms], n_points=32, q_max=float(g["q_max"]),
                              verbose=False).run()
    np.testing.assert_allclose(su.results.ssf, g["ssf_unique"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(su.results.wavenumbers, g["wavenumbers_unique"],
                               rtol=1e-13)
    # approximate FP32 mode: report, and hold to a loose bound
    sf = _S().StructureFactor([u.atoms], n_points=32, q_max=float(g["q_max"]),
                              sort=False, unique=False, precision="fp32",
                              verbose=False).run()
    rel = np.abs(sf.results.ssf - g["ssf_raw"]) / g["ssf_raw"]
    print("fp32 mode: max rel err", rel.max(), "median", np.median(rel))
    assert rel.max() < 1e-3


def test_default_grid_n_points_32_no_qmax():
    """32,768 wavevectors (nz up to 31: two column segments per (nx, ny))."""
    from mdhelper_b200 import synthetic
    from oracle import reference_port as rp
    u = synthetic.lj_fluid(300, 2, seed=21)
    want = rp.ssf_run(u, [u.atoms], sort=False, unique=False, n_threads=8)
    for kernel in (None, "lattice_fp64"):
        s = _S().StructureFactor([u.atoms], sort=False, unique=False, kernel=kernel,
                                 verbose=False).run()
        assert s.results.ssf.shape == (1, 32768)
        np.testing.assert_allclose(s.results.ssf, want["ssf"], rtol=1e-9, atol=1e-11)


def test_mma_tiling_odd_shapes():
    """DMMA kernel: grids whose column count is not a multiple of 8, a single column,
    nz ranges that need a partial last tile and more than four tiles, anisotropic
    n_max, two groups with a cross term -- against the general (sincos) kernel."""
    from mdhelper_b200 import synthetic
    u = synthetic.lj_fluid(777, 3, seed=31)
    L = float(u.dimensions[0])
    cases = [dict(n_points=3), dict(n_points=5, q_max=2 * np.pi * 3.3 / L),
             dict(n_points=40, q_max=2 * np.pi * 5.01 / L), dict(n_points=37, q_max=None)]
    for kw in cases:
        kw = {k: v for k, v in kw.items() if v is not None}
        a = _S().StructureFactor([u.atoms], sort=False, unique=False, verbose=False,
                                 kernel="lattice_dmma", **kw).run()
        assert a._ctx.sq_kernel() == "lattice_dmma"
        b = _S().StructureFactor([u.atoms], sort=False, unique=False, verbose=False,
                                 kernel="general_fp64", **kw).run()
        np.testing.assert_allclose(a.results.ssf, b.results.ssf, rtol=1e-9, atol=1e-10)
    # user wavevectors on the lattice: one column (0, 0, nz) and a sparse set
    b3 = 2 * np.pi / np.asarray(u.dimensions[:3], dtype=np.float64)
    n = np.array([[0, 0, z] for z in range(0, 45, 3)] + [[7, 2, 1], [1, 30, 0], [9, 9, 9]])
    wv = n * b3
    g1, g2 = u.select(slice(0, 300)), u.select(slice(300, 777))
    a = _S().StructureFactor([g1, g2], mode="partial", wavevectors=wv, sort=False,
                             unique=False, kernel="lattice_dmma", verbose=False).run()
    assert a._ctx.sq_kernel() == "lattice_dmma"
    b = _S().StructureFactor([g1, g2], mode="partial", wavevectors=wv, sort=False,
                             unique=False, kernel="general_fp64", verbose=False).run()
    np.testing.assert_allclose(a.results.ssf, b.results.ssf, rtol=1e-9, atol=1e-10)


def test_frame_additivity_and_selection():
    from mdhelper_b200 import synthetic
    from oracle import reference_port as rp
    u = synthetic.lj_fluid(500, 7, seed=22)
    kw = dict(n_points=8, q_max=3.0)
    s = _S().StructureFactor([u.atoms], batch_frames=2, verbose=False, **kw).run(
        start=1, step=2)
    want = rp.ssf_run(u, [u.atoms], frames=[1, 3, 5], **kw)
    np.testing.assert_allclose(s.results.ssf, want["ssf"], rtol=1e-9, atol=1e-12)
    # residue centres of mass (host helper) on the general path
    um = synthetic.polymer_melt(30, 4, 2, seed=23)
    s = _S().StructureFactor([um.atoms], groupings="residues", mode="pair",
                             verbose=False, **kw).run()
    assert s.results.ssf.shape[0] == 1 and np.isfinite(s.results.ssf).all()


def test_auto_kernel_choice_and_tiling_report():
    """MDH_SQ_AUTO: small lattice sets run on the scalar kernel, larger ones on the matrix
    unit, off-lattice wavevectors on the general kernel; mdh_sq_tiling agrees with the
    device-free mdh_sq_plan."""
    from mdhelper_b200 import _lib, synthetic
    u = synthetic.lj_fluid(400, 1, seed=2)
    small = _S().StructureFactor([u.atoms], n_points=8, verbose=False).run()       # 8 pairs
    assert small._ctx.sq_kernel() == "lattice_fp64"
    assert small._ctx.sq_tiling()["tiles"] == 0
    big = _S().StructureFactor([u.atoms], n_points=12, verbose=False).run()        # 36 pairs
    assert big._ctx.sq_kernel() == "lattice_dmma"
    plan = _lib.sq_plan(big._lattice_n)
    tiling = big._ctx.sq_tiling()
    assert {k: plan[k] for k in tiling} == tiling and tiling["tiles"] == 36
    off = _S().StructureFactor([u.atoms], n_points=6, n_surfaces=2, n_surface_points=8,
                               verbose=False).run()
    assert off._ctx.sq_kernel() == "general_fp64"


def test_mma_pipeline_long_run():
    """DMMA kernel: thousands of work units per launch (every persistent block walks many
    units, the table ring and its mbarrier phases wrap hundreds of times), a particle count
    that leaves a partial last sub-chunk in every chunk, two groups of unequal size -- the
    scalar lattice kernel must agree, and so must a second run of the same object."""
    from mdhelper_b200 import synthetic
    u = synthetic.lj_fluid(1999, 130, seed=41)
    L = float(u.dimensions[0])
    g1, g2 = u.select(slice(0, 777)), u.select(slice(777, 1999))
    kw = dict(mode="partial", n_points=12, q_max=2 * np.pi * 9.5 / L, sort=False,
              unique=False, verbose=False, batch_frames=130)
    a = _S().StructureFactor([g1, g2], kernel="lattice_dmma", **kw).run()
    assert a._ctx.sq_kernel() == "lattice_dmma"
    first = a.results.ssf.copy()
    b = _S().StructureFactor([g1, g2], kernel="lattice_fp64", **kw).run()
    np.testing.assert_allclose(first, b.results.ssf, rtol=1e-10, atol=1e-9)
    a.run()
    np.testing.assert_allclose(a.results.ssf, first, rtol=1e-12, atol=1e-10)


# ---- intermediate scattering function ------------------

@pytest.mark.parametrize("kernel", [None, "lattice_dmma", "lattice_fp64", "general_fp64"])
@pytest.mark.parametrize("mode", [None, "pair", "partial"])
def test_isf_matches_golden(golden, mode, kernel):
    """GPU IntermediateScatteringFunction vs the fixtures of the reference's real class:
    coherent and incoherent parts, all modes, lattice and general kernels, a batch size
    that cuts the run into several accumulate calls (window hand-over)."""
    g = golden("isf_small")
    u = universe_from(g)
    cat, an = _groups(u, g)
    for form, bf in (("exp", 3), ("trig", None)):
        r = _S().IntermediateScatteringFunction(
            [cat, an], mode=mode, form=form, n_points=int(g["n_points"]),
            q_max=float(g["q_max"]), n_lags=int(g["n_lags"]), incoherent=True,
            dt=float(g["dt"]), kernel=kernel, batch_frames=bf, verbose=False).run()
        np.testing.assert_allclose(r.results.cisf, g[f"cisf_{mode}_{form}"], rtol=1e-9,
                                   atol=1e-10)
        np.testing.assert_allclose(r.results.iisf, g[f"iisf_{mode}_{form}"], rtol=1e-9,
                                   atol=1e-10)
        np.testing.assert_allclose(r.results.times, g[f"times_{mode}_{form}"])
        np.testing.assert_allclose(r.results.wavenumbers,
                                   g[f"wavenumbers_{mode}_{form}"], rtol=1e-13)


def test_isf_strided_frames_user_wavevectors_and_properties(golden):
    g = golden("isf_small")
    u = universe_from(g)
    cat, an = _groups(u, g)
    S = _S()
    r = S.IntermediateScatteringFunction(
        [u.atoms], n_points=int(g["n_points"]), q_max=float(g["q_max"]), incoherent=True,
        sort=False, unique=False, dt=float(g["dt"]), verbose=False,
        batch_frames=2).run(start=2, stop=14, step=3)
    np.testing.assert_allclose(r.results.cisf, g["cisf_strided"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r.results.iisf, g["iisf_strided"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r.results.times, g["times_strided"])
    # F_s(q, 0) = 1;  F(q, 0) = S(q) of the same frames
    np.testing.assert_allclose(r.results.iisf[0], 1.0, rtol=1e-12)
    s = S.StructureFactor([u.atoms], n_points=int(g["n_points"]), q_max=float(g["q_max"]),
                          sort=False, unique=False, verbose=False).run(start=2, stop=14,
                                                                       step=3)
    np.testing.assert_allclose(r.results.cisf[0], s.results.ssf, rtol=1e-9, atol=1e-12)
    # without the incoherent part; user wavevectors off the lattice
    r = S.IntermediateScatteringFunction(
        [cat, an], mode="partial", wavevectors=g["wavevectors_user"], n_lags=4,
        incoherent=True, sort=False, unique=False, dt=float(g["dt"]), verbose=False).run()
    np.testing.assert_allclose(r.results.cisf, g["cisf_user"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r.results.iisf, g["iisf_user"], rtol=1e-9, atol=1e-10)
    r2 = S.IntermediateScatteringFunction(
        [cat, an], mode="partial", wavevectors=g["wavevectors_user"], n_lags=4,
        sort=False, unique=False, dt=float(g["dt"]), verbose=False).run()
    assert r2.results.iisf is None
    np.testing.assert_allclose(r2.results.cisf, r.results.cisf, rtol=1e-13)
    with pytest.raises(ValueError):        # frames must be evenly spaced
        S.IntermediateScatteringFunction([u.atoms], n_points=4, verbose=False).run(
            frames=[0, 1, 3])


def test_isf_larger_system_against_oracle():
    """5,000 particles, 12 frames, 6 lags, lattice kernel vs the CPU oracle."""
    from mdhelper_b200 import synthetic
    from oracle import reference_port as rp
    u = synthetic.lj_fluid(5000, 12, seed=77)
    L = float(u.dimensions[0])
    kw = dict(n_points=10, q_max=2 * np.pi * 6 / L, n_lags=6, incoherent=True, dt=2.0)
    r = _S().IntermediateScatteringFunction([u.atoms], verbose=False, batch_frames=5,
                                            **kw).run()
    o = rp.isf_run(u, [u.atoms], n_threads=8, **kw)
    np.testing.assert_allclose(r.results.cisf, o["cisf"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r.results.iisf, o["iisf"], rtol=1e-9, atol=1e-10)


def test_device_centres_of_mass_equal_host():
    """groupings="residues" with the centres of mass built on the device vs the host
    helper (identical float32 centres -> S(q) to rounding), mixed with an atoms group."""
    from mdhelper_b200.universe import SyntheticUniverse
    rng = np.random.default_rng(5)
    sizes = rng.integers(1, 6, 500)
    res = np.repeat(np.arange(500), sizes)
    n = res.size
    dims = np.array([12.0, 12.0, 12.0, 90, 90, 90], np.float32)
    pos = (rng.random((3, n, 3)) * 12).astype(np.float32)
    u = SyntheticUniverse(pos, dims, resindices=res, masses=rng.uniform(1, 30, n))
    half = int(np.searchsorted(res, 250))
    g1, g2 = u.select(slice(0, half)), u.select(slice(half, n))
    kw = dict(mode="partial", n_points=6, verbose=False)
    for grp in (("residues", "residues"), ("residues", "atoms")):
        d = _S().StructureFactor([g1, g2], grp, **kw).run()
        h = _S().StructureFactor([g1, g2], grp, host_com=True, **kw).run()
        assert d._com is not None and h._com is None
        np.testing.assert_allclose(d.results.ssf, h.results.ssf, rtol=1e-12, atol=1e-13)


# ---- single-chain structure factor ---------------------

def test_scsf_matches_golden(golden):
    """GPU SingleChainStructureFactor vs the fixtures of the reference's real class:
    wrapped coordinates, host unwrapping, a strided frame selection, chains taken from
    the segment information."""
    from mdhelper_b200.analysis.polymer import SingleChainStructureFactor
    from mdhelper_b200 import synthetic
    g = golden("scsf_small")
    u = universe_from(g)
    kw = dict(n_points=int(g["n_points"]), n_chains=int(g["n_chains"]),
              n_monomers=int(g["n_monomers"]), verbose=False)
    for unwrap, kernel in ((False, None), (True, "lattice_dmma"), (False, "lattice_dmma"),
                           (False, "lattice_fp64")):
        r = SingleChainStructureFactor(u.atoms, unwrap=unwrap, batch_frames=4,
                                       kernel=kernel, **kw).run()
        assert kernel is None or r._ctx.sq_kernel() == kernel
        np.testing.assert_allclose(r.results.scsf, g[f"scsf_unwrap{int(unwrap)}"],
                                   rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(r.results.wavenumbers, g["wavenumbers"], rtol=1e-13)
    r = SingleChainStructureFactor(u.atoms, **kw).run(start=1, stop=6, step=2)
    np.testing.assert_allclose(r.results.scsf, g["scsf_strided"], rtol=1e-9, atol=1e-10)
    # chain bookkeeping from the universe's segments (polymer_melt sets them)
    u2 = synthetic.polymer_melt(12, 20, 6, seed=20260012)
    r2 = SingleChainStructureFactor(u2.atoms, n_points=int(g["n_points"]),
                                    verbose=False).run()
    np.testing.assert_allclose(r2.results.scsf, g["scsf_unwrap0"], rtol=1e-9, atol=1e-10)
    with pytest.raises(ValueError):
        SingleChainStructureFactor(u.atoms, grouping="segments", **kw)


def test_scsf_many_chains_against_oracle():
    """More chains than one grid dimension takes at once is not needed here, but chain
    lengths that are not a multiple of the 32-particle sub-chunk and a non-cubic box
    are: 300 chains x 37 monomers vs the CPU oracle."""
    from mdhelper_b200.analysis.polymer import SingleChainStructureFactor
    from mdhelper_b200.universe import SyntheticUniverse
    from oracle import reference_port as rp
    rng = np.random.default_rng(12)
    dims = np.array([21.0, 24.0, 27.0, 90, 90, 90], np.float32)
    pos = (rng.random((2, 300 * 37, 3)) * dims[:3]).astype(np.float32)
    u = SyntheticUniverse(pos, dims)
    o = rp.scsf_run(u, u.atoms, n_points=7, n_chains=300, n_monomers=37)
    for kernel in ("lattice_dmma", "lattice_fp64"):
        r = SingleChainStructureFactor(u.atoms, n_points=7, n_chains=300, n_monomers=37,
                                       kernel=kernel, verbose=False).run()
        assert r._ctx.sq_kernel() == kernel
        np.testing.assert_allclose(r.results.scsf, o["scsf"], rtol=1e-9, atol=1e-10)


def test_combined_pass_equals_separate_runs():
    """CombinedAnalysis (one upload per batch shared by RDF and S(q), BASELINE cfg5):
    identical counts, S(q) to rounding; strided frames; fallback for scattered groups."""
    from mdhelper_b200 import synthetic
    from mdhelper_b200.analysis import CombinedAnalysis
    u, cat, an = synthetic.electrolyte(6000, 11, seed=8)
    L = float(u.dimensions[0])
    S = _S()
    mk_r = lambda a, b: S.RadialDistributionFunction(a, b, n_bins=50, range=(0.0, 4.0),  # noqa
                                                     verbose=False)
    mk_s = lambda g: S.StructureFactor(g, mode="partial" if len(g) > 1 else None,  # noqa
                                       n_points=8, q_max=2 * np.pi * 4 / L, verbose=False)
    for run_kw in (dict(), dict(start=1, stop=11, step=3)):
        r0, s0 = mk_r(cat, an).run(**run_kw), mk_s([cat, an]).run(**run_kw)
        r1, s1 = mk_r(cat, an), mk_s([cat, an])
        CombinedAnalysis(r1, s1, batch_frames=4).run(**run_kw)
        assert np.array_equal(r1.results.counts, r0.results.counts)
        np.testing.assert_allclose(r1.results.rdf, r0.results.rdf, rtol=1e-13)
        np.testing.assert_allclose(s1.results.ssf, s0.results.ssf, rtol=1e-12)
    # scattered selection -> per-analysis fallback, same results
    g = u.select(np.random.default_rng(0).permutation(6000)[:2000])
    r0 = mk_r(g, None).run()
    r1, s1 = mk_r(g, None), mk_s([u.atoms])
    CombinedAnalysis(r1, s1).run()
    assert np.array_equal(r1.results.counts, r0.results.counts)
    np.testing.assert_allclose(s1.results.ssf, mk_s([u.atoms]).run().results.ssf, rtol=1e-12)


def test_host_batches_in_overlapped_pieces():
    """S(q) host batches above ~24 MB go through the copy stream in pieces (uneven
    split: 45 frames of 600 kB -> 2 pieces of 23 + 22); same sums as small batches."""
    from mdhelper_b200 import synthetic
    u = synthetic.lj_fluid(50_000, 45, seed=3)
    L = float(u.dimensions[0])
    kw = dict(n_points=8, q_max=2 * np.pi * 4 / L, sort=False, unique=False, verbose=False)
    a = _S().StructureFactor([u.atoms], batch_frames=45, **kw).run()
    b = _S().StructureFactor([u.atoms], batch_frames=7, **kw).run()
    np.testing.assert_allclose(a.results.ssf, b.results.ssf, rtol=1e-12)


def test_argument_errors():
    from mdhelper_b200 import _lib
    ctx = _lib.Context(0)
    wv = np.eye(3)
    with pytest.raises(RuntimeError):
        ctx.sq_accumulate(np.zeros((4, 3), np.float32), 12, 1)
    with pytest.raises(ValueError):
        ctx.sq_configure(4, [0, 3], wv, [(-1, -1)])          # offsets do not reach n
    with pytest.raises(ValueError):
        ctx.sq_configure(4, [0, 4], wv, [(0, 1)])            # missing group
    with pytest.raises(ValueError):
        ctx.sq_configure(4, [0, 4], wv, [(-1, -1)], mode="lattice_fp64")  # no lattice
    ctx.close()


@pytest.mark.parametrize("tag", ["equal", "unequal"])
def test_residue_structure_factor_against_the_reference(golden, tag):
    """StructureFactor(groupings="residues") vs the reference's own class on its own
    ``center_of_mass`` (tests/golden/com_ref.npz).  The centres stay float64 up to the
    Fourier sums, as in the reference's position buffer (structure.py:1468): centres
    rounded to float32 would leave phase errors of q |r| 6e-8, ~1e-6 of S(q)."""
    from mdhelper_b200.universe import SyntheticUniverse
    g = golden("com_ref")
    u = SyntheticUniverse(g[f"{tag}_positions"], g[f"{tag}_dims"],
                          resindices=g[f"{tag}_resindices"],
                          segindices=g[f"{tag}_segindices"], masses=g[f"{tag}_masses"])
    n_a, n = int(g[f"{tag}_n_a"]), u.atoms.n_atoms
    a, b = u.select(slice(0, n_a)), u.select(slice(n_a, n))
    for host_com in (False, True):
        s = _S().StructureFactor([a, b], groupings="residues", mode="partial", n_points=5,
                                 verbose=False, host_com=host_com).run()
        np.testing.assert_allclose(s.results.wavenumbers, g[f"{tag}_ssf_wavenumbers"],
                                   rtol=1e-12)
        np.testing.assert_allclose(s.results.ssf, g[f"{tag}_ssf_res"], rtol=1e-9, atol=1e-10)


def test_structure_factor_of_a_triclinic_universe_uses_the_edge_lengths():
    """The reference builds its wavevectors from ``dimensions[:3]`` alone
    (structure.py:1366-1381), angles ignored: same result as the orthorhombic cell."""
    from mdhelper_b200.universe import SyntheticUniverse
    rng = np.random.default_rng(2)
    pos = (rng.random((2, 500, 3)) * 9).astype(np.float32)
    ut = SyntheticUniverse(pos, np.array([9.0, 10.0, 11.0, 80, 70, 60], np.float32))
    uo = SyntheticUniverse(pos, np.array([9.0, 10.0, 11.0, 90, 90, 90], np.float32))
    st = _S().StructureFactor([ut.atoms], n_points=6, verbose=False).run()
    so = _S().StructureFactor([uo.atoms], n_points=6, verbose=False).run()
    assert np.array_equal(st.results.wavenumbers, so.results.wavenumbers)
    np.testing.assert_allclose(st.results.ssf, so.results.ssf, rtol=1e-12)


def test_grids_too_large_for_shared_memory_tables_use_the_general_kernel():
    """n_points = 96: the phase-factor tables of the lattice kernels (3 KB x n_points per
    sub-chunk) exceed the 227 KB of shared memory; AUTO then takes the general kernel
    (the reference has no such limit), an explicit lattice kernel is refused."""
    from mdhelper_b200 import _lib
    from mdhelper_b200.universe import SyntheticUniverse
    rng = np.random.default_rng(5)
    L = np.float32(9.0)
    pos = (rng.random((1, 60, 3)) * float(L)).astype(np.float32)
    u = SyntheticUniverse(pos, np.array([L, L, L, 90, 90, 90], np.float32))
    q_cut = 2 * np.pi * 14.5 / float(L)
    # a small sphere of a 96-point grid is fine (tables sized by the largest index used)
    s = _S().StructureFactor([u.atoms], n_points=96, q_max=q_cut, verbose=False).run()
    assert s._ctx.sq_kernel() in ("lattice_dmma", "lattice_fp64")
    # a thin shell reaching index 95 along each axis needs tables of 96 entries per axis
    idx = np.array([[95, 0, 0], [0, 95, 0], [0, 0, 95], [95, 95, 95], [1, 2, 3], [0, 0, 0]])
    wv = 2 * np.pi * idx / float(L)
    s = _S().StructureFactor([u.atoms], wavevectors=wv, unique=False, sort=False,
                             verbose=False).run()
    assert s._ctx.sq_kernel() == "general_fp64"
    want = np.abs(np.exp(1j * pos[0].astype(np.float64) @ wv.T).sum(axis=0)) ** 2 / 60
    np.testing.assert_allclose(s.results.ssf[0], want, rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError):
        _S().StructureFactor([u.atoms], wavevectors=wv, unique=False, sort=False,
                             verbose=False, kernel="lattice_fp64").run()


# ---- float64 coordinates (centres of mass, unwrapped positions) ------------------------

@pytest.mark.parametrize("kernel", ["lattice_dmma", "lattice_fp64", "general_fp64"])
def test_float64_coordinates_are_not_rounded_to_float32(kernel):
    """mdh_sq_accumulate_f64 vs a numpy direct sum over the same float64 coordinates, far
    from the origin (|r| ~ 1e3, where a float32 copy is off by 6e-5 and S(q) by far more
    than the bound): host and device input, more frames than one piece."""
    import torch
    from mdhelper_b200 import _lib
    rng = np.random.default_rng(20260641)
    n, F, L = 777, 5, 17.0
    pos = 1000.0 + rng.random((F, n, 3)) * L
    nvec = np.array([(a, b, c) for a in range(4) for b in range(4) for c in range(5)][1:],
                    dtype=np.int32)
    b = 2 * np.pi / np.array([L, L, L])
    wv = nvec * b
    want = np.zeros(len(wv))
    for f in range(F):
        rho = np.exp(1j * (pos[f] @ wv.T)).sum(axis=0)
        want += (rho * rho.conj()).real
    ctx = _lib.Context(0)
    lat = {} if kernel == "general_fp64" else dict(lattice_n=nvec, lattice_b=b)
    ctx.sq_configure(n, [0, n], wv, [(-1, -1)], mode=kernel, **lat)
    ctx.sq_accumulate(pos, 3 * n, F, f64=True)
    got = ctx.sq_fetch()[0]
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-7)
    ctx.sq_reset()
    dev = torch.from_numpy(pos).cuda()
    ctx.sq_accumulate(dev.data_ptr(), 3 * n, F, device=True, f64=True)
    np.testing.assert_allclose(ctx.sq_fetch()[0], got, rtol=1e-12, atol=1e-9)
    # the float32 path on the rounded coordinates is what this entry point avoids
    ctx.sq_reset()
    ctx.sq_accumulate(pos.astype(np.float32), 3 * n, F)
    rounded = ctx.sq_fetch()[0]
    assert np.abs(rounded - want).max() > 1e3 * np.abs(got - want).max()
    # float32-valued doubles: the same phases as the float32 entry point
    ctx.sq_reset()
    ctx.sq_accumulate(pos.astype(np.float32).astype(np.float64), 3 * n, F, f64=True)
    np.testing.assert_allclose(ctx.sq_fetch()[0], rounded, rtol=1e-12, atol=1e-9)
    ctx.close()


class _Centres:
    """What the oracle's loops need from a group: the float64 centres of mass of the
    residues of ``group`` in the current frame."""

    def __init__(self, group):
        self._g = group
        _, self._inv = np.unique(group.resindices, return_inverse=True)
        self._m = np.asarray(group.masses, dtype=np.float64)
        self.n_atoms = int(self._inv.max()) + 1

    @property
    def positions(self):
        p = np.asarray(self._g.positions, dtype=np.float64)
        out = np.zeros((self.n_atoms, 3))
        np.add.at(out, self._inv, self._m[:, None] * p)
        return out / np.bincount(self._inv, weights=self._m)[:, None]


def test_isf_of_residue_centres_against_oracle():
    """IntermediateScatteringFunction(groupings="residues"): centres of mass and their
    displacements in float64 (structure.py:1927-1957) vs the oracle's sliding window on
    float64 centres, several batches so that the coordinate window is carried over."""
    from mdhelper_b200.universe import SyntheticUniverse
    from oracle import reference_port as rp
    rng = np.random.default_rng(20260642)
    sizes = rng.integers(1, 5, 300)
    res = np.repeat(np.arange(300), sizes)
    n, F, L = res.size, 9, 14.0
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    pos = (rng.random((1, n, 3)) * L + rng.normal(0, 0.3, (F, n, 3)).cumsum(axis=0)
           ).astype(np.float32)
    u = SyntheticUniverse(pos, dims, resindices=res, masses=rng.uniform(1, 30, n))
    half = int(np.searchsorted(res, 150))
    g1, g2 = u.select(slice(0, half)), u.select(slice(half, n))
    kw = dict(mode="partial", n_points=5, n_lags=4, incoherent=True, dt=1.0)
    r = _S().IntermediateScatteringFunction([g1, g2], groupings="residues", verbose=False,
                                            batch_frames=4, **kw).run()
    o = rp.isf_run(u, [_Centres(g1), _Centres(g2)], n_threads=4, **kw)
    np.testing.assert_allclose(r.results.cisf, o["cisf"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r.results.iisf, o["iisf"], rtol=1e-9, atol=1e-10)


def _drifting_chains(rng, n_chains, n_mono, F, L, per_bead=1):
    start = rng.random((n_chains, 1, 3)) * L
    chain = start + rng.normal(0, 0.5, (n_chains, n_mono, 3)).cumsum(axis=1)
    drift = np.arange(F)[:, None, None, None] * np.array([4.3, -3.1, 2.2])
    true = chain[None] + drift + rng.normal(0, 0.05, (F, n_chains, n_mono, 3))
    true = np.repeat(true.reshape(F, -1, 3), per_bead, axis=1)
    if per_bead > 1:
        true = true + rng.normal(0, 0.1, true.shape)
    wrapped = np.mod(true, L).astype(np.float32)
    wrapped[wrapped >= L] = 0.0
    return wrapped


def test_scsf_of_unwrapped_chains_far_from_the_origin():
    """SingleChainStructureFactor(unwrap=True) of chains that drift many box lengths.
    grouping="atoms": the reference unwraps the reader's float32 array in place
    (polymer.py:1079, 1091-1093), i.e. rounds the unwrapped coordinates to float32 -- so
    does this path, vs the oracle.  grouping="residues": centres of mass are float64 and
    are unwrapped in float64 (polymer.py:1080-1093); vs the same sums in numpy."""
    from mdhelper_b200.analysis.polymer import SingleChainStructureFactor
    from mdhelper_b200.universe import SyntheticUniverse
    from oracle import reference_port as rp
    rng = np.random.default_rng(20260643)
    n_chains, n_mono, F, L = 20, 16, 8, 10.0
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    u = SyntheticUniverse(_drifting_chains(rng, n_chains, n_mono, F, L), dims)
    kw = dict(n_points=5, n_chains=n_chains, n_monomers=n_mono, unwrap=True)
    r = SingleChainStructureFactor(u.atoms, verbose=False, batch_frames=3, **kw).run()
    o = rp.scsf_run(u, u.atoms, **kw)
    np.testing.assert_allclose(r.results.scsf, o["scsf"], rtol=1e-9, atol=1e-10)

    # three atoms per monomer, one residue each
    per = 3
    wrapped = _drifting_chains(rng, n_chains, n_mono, F, L, per_bead=per)
    n = wrapped.shape[1]
    masses = rng.uniform(1, 20, n)
    u = SyntheticUniverse(wrapped, dims, resindices=np.arange(n) // per, masses=masses)
    r = SingleChainStructureFactor(u.atoms, grouping="residues", verbose=False,
                                   batch_frames=3, **kw).run()
    m = masses.reshape(-1, per)
    wv = r._wavevectors
    scsf = np.zeros(len(wv))
    old = images = None
    box = dims[:3].astype(np.float64)
    for f in range(F):
        p = wrapped[f].astype(np.float64).reshape(-1, per, 3)
        com = (m[:, :, None] * p).sum(axis=1) / m.sum(axis=1, keepdims=True)
        if old is None:
            old, images = com.copy(), np.zeros(com.shape, dtype=int)
        d = com - old
        mask = np.abs(d) >= box / 2
        images[mask] -= np.sign(d[mask]).astype(int)
        old = com.copy()
        com = com + images * box
        for chain in com.reshape(n_chains, n_mono, 3):
            arg = chain @ wv.T
            scsf += np.sin(arg).sum(axis=0) ** 2 + np.cos(arg).sum(axis=0) ** 2
    scsf /= n_chains * n_mono * F
    want = np.array([scsf[np.isclose(q, r._wavenumbers)].mean()
                     for q in r.results.wavenumbers])
    np.testing.assert_allclose(r.results.scsf, want, rtol=1e-9, atol=1e-10)


def test_float64_paths_against_the_reference_classes(golden):
    """tests/golden/f64_ref.npz: the reference's REAL IntermediateScatteringFunction with
    groupings="residues" and SingleChainStructureFactor with grouping="residues" (wrapped,
    and unwrapped over several box lengths) -- positions the reference keeps in float64."""
    from mdhelper_b200.analysis.polymer import SingleChainStructureFactor
    from mdhelper_b200.universe import SyntheticUniverse
    g = golden("f64_ref")
    u = SyntheticUniverse(g["isf_positions"], g["isf_dims"], resindices=g["isf_resindices"],
                          masses=g["isf_masses"])
    n_a, n = int(g["isf_n_a"]), u.atoms.n_atoms
    a, b = u.select(slice(0, n_a)), u.select(slice(n_a, n))
    r = _S().IntermediateScatteringFunction([a, b], groupings="residues", mode="partial",
                                            n_points=5, n_lags=4, incoherent=True, dt=1.0,
                                            verbose=False, batch_frames=3).run()
    np.testing.assert_allclose(r.results.wavenumbers, g["isf_wavenumbers"], rtol=1e-12)
    np.testing.assert_allclose(r.results.cisf, g["isf_cisf"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r.results.iisf, g["isf_iisf"], rtol=1e-9, atol=1e-10)

    per = int(g["scsf_per"])
    pos = g["scsf_positions"]
    u = SyntheticUniverse(pos, g["scsf_dims"], resindices=np.arange(pos.shape[1]) // per,
                          masses=g["scsf_masses"])
    for unwrap in (False, True):
        s = SingleChainStructureFactor(u.atoms, grouping="residues", n_points=5,
                                       n_chains=int(g["scsf_n_chains"]),
                                       n_monomers=int(g["scsf_n_monomers"]), unwrap=unwrap,
                                       verbose=False, batch_frames=3).run()
        np.testing.assert_allclose(s.results.wavenumbers, g["scsf_wavenumbers"], rtol=1e-12)
        np.testing.assert_allclose(s.results.scsf, g[f"scsf_res_unwrap{int(unwrap)}"],
                                   rtol=1e-9, atol=1e-10)
