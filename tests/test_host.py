"""CPU: host-side logic (universe, frame feeder, frame selection, wavevectors, normalisation)."""
import numpy as np
import pytest

from mdhelper_b200 import synthetic
from mdhelper_b200.analysis import base, structure
from mdhelper_b200.analysis._binning import squared_thresholds
from mdhelper_b200.universe import SyntheticUniverse
from oracle import reference_port as rp


def test_thresholds_reproduce_numpy_histogram():
    rng = np.random.default_rng(3)
    for nb, rg in [(201, (0, 15.0)), (10, (0, 11)), (50, (1.0, 9.0)), (37, (2.0, 7.3))]:
        T = squared_thresholds(nb, rg)
        edges = np.linspace(*rg, nb + 1)
        d2 = np.concatenate([rng.random(200_000) * (rg[1] * 1.2) ** 2, edges ** 2,
                             np.nextafter(edges ** 2, np.inf),
                             np.nextafter(edges ** 2, -np.inf), T,
                             np.nextafter(T, -np.inf), np.nextafter(T, np.inf)])
        d2 = d2[d2 >= 0]
        d = np.sqrt(d2)
        keep = (d > rg[0] - np.finfo(float).eps) & (d <= rg[1])
        ref = np.histogram(d[keep], bins=nb, range=rg)[0]
        k = np.searchsorted(T, d2, side="right") - 1
        ok = (k >= 0) & (k < nb)
        assert np.array_equal(ref, np.bincount(k[ok], minlength=nb))


def test_thresholds_reject_bad_ranges():
    with pytest.raises(ValueError):
        squared_thresholds(10, (3.0, 1.0))
    with pytest.raises(ValueError):
        squared_thresholds(0, (0.0, 1.0))


def test_universe_protocol():
    u = synthetic.lj_fluid(100, 4, seed=2)
    tr = u.trajectory
    assert len(tr) == 4 and tr.n_atoms == 100
    ts = tr[2]
    assert ts.frame == 2 and ts.positions.dtype == np.float32
    assert u.atoms.positions.shape == (100, 3)
    L = float(ts.dimensions[0])
    assert ts.volume == pytest.approx(L ** 3, rel=1e-12)
    assert tr.check_slice_indices(None, None, None) == (0, 4, 1)
    assert [t.frame for t in tr[1:4:2]] == [1, 3]
    g = u.select(slice(10, 20))
    assert g.n_atoms == 10 and g == u.select(np.arange(10, 20))
    with pytest.raises(IndexError):
        tr[7]


def test_feeder_zero_copy_equals_staged():
    u = synthetic.lj_fluid(50, 7, seed=4)
    sets = [np.arange(5, 30)]
    frames = np.arange(1, 7, 2)
    zc = base.FrameFeeder(u.trajectory, sets, frames, 2)
    assert zc.zero_copy
    st = base.FrameFeeder(u.trajectory, [np.array([5, 7, 6] + list(range(8, 30)))],
                          frames, 2)
    assert not st.zero_copy
    import ctypes
    got = []
    for b in zc:
        for f in range(b.n_frames):
            addr = b.ptrs[0] + 4 * b.strides[0] * f
            a = np.ctypeslib.as_array((ctypes.c_float * 75).from_address(addr))
            got.append(a.reshape(25, 3).copy())
    want = [u.trajectory.coordinates[f, 5:30] for f in frames]
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    got = []
    for b in st:
        a = b.keepalive[0][0][:b.n_frames]
        got.extend(x.copy() for x in a)
    order = [5, 7, 6] + list(range(8, 30))
    want = [u.trajectory.coordinates[f, order] for f in frames]
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


def test_setup_frames_semantics():
    u = synthetic.lj_fluid(10, 9, seed=5)
    a = base.GpuAnalysisBase(u.trajectory)
    a._setup_frames(u.trajectory, 2, 9, 3)
    assert list(a._frame_list) == [2, 5, 8] and a.n_frames == 3
    a._setup_frames(u.trajectory, frames=[0, 4])
    assert list(a._frame_list) == [0, 4]
    a._setup_frames(u.trajectory, frames=np.array([True] + [False] * 7 + [True]))
    assert list(a._frame_list) == [0, 8]
    with pytest.raises(ValueError):
        a._setup_frames(u.trajectory, start=1, frames=[0])


def test_wavevector_grid_matches_reference_construction():
    u = synthetic.lj_fluid(20, 1, seed=6)
    for n_points, q_max in ((6, None), (9, 2.0)):
        s = structure.StructureFactor([u.atoms], n_points=n_points, q_max=q_max)
        wv, wn = rp.lattice_wavevectors(u.dimensions[:3].copy(), n_points, q_max)
        assert np.array_equal(s._wavevectors, wv)
        assert np.array_equal(s._wavenumbers, wn)
        n, b = s._lattice_n, s._lattice_b
        np.testing.assert_allclose(n * b, wv, rtol=4e-16, atol=0)
    # non-cubic
    u = SyntheticUniverse(np.zeros((1, 4, 3), np.float32),
                          np.array([9, 11.5, 14.25, 90, 90, 90], np.float32))
    s = structure.StructureFactor([u.atoms], n_points=5, q_max=3.0)
    wv, _ = rp.lattice_wavevectors(u.dimensions[:3].copy(), 5, 3.0)
    assert np.array_equal(s._wavevectors, wv)
    # user wavevectors on / off the lattice
    s2 = structure.StructureFactor([u.atoms], wavevectors=wv[::-1].copy())
    assert s2._lattice_n is not None
    s3 = structure.StructureFactor([u.atoms], wavevectors=wv + 0.01)
    assert s3._lattice_n is None


def test_constructor_errors_match_reference_messages():
    u = synthetic.lj_fluid(20, 1, seed=6)
    with pytest.raises(ValueError, match="Invalid grouping"):
        structure.RadialDistributionFunction(u.atoms, groupings="molecules")
    with pytest.raises(ValueError, match="Invalid axis to drop"):
        structure.RadialDistributionFunction(u.atoms, drop_axis=5)
    with pytest.raises(ValueError, match="exactly one or two groups"):
        structure.StructureFactor([u.atoms[:5], u.atoms[5:10], u.atoms[10:]], mode="pair")
    with pytest.raises(ValueError, match="do not contain all atoms"):
        structure.StructureFactor([u.atoms[:5]])
    with pytest.raises(ValueError, match="must have length 3"):
        structure.StructureFactor([u.atoms], dimensions=(1, 2))
    assert structure.RadialDistributionFunction(u.atoms, drop_axis="z")._drop_axis == 2


from _fake import FakeRDF as _FakeRDF, FakeSSF  # noqa: E402


@pytest.mark.parametrize("name", ["lj1000", "excl410", "dropz", "dropx_density",
                                  "noncubic_npt"])
def test_rdf_normalisation_matches_golden(golden, name):
    from test_oracle import rdf_groups, rdf_kwargs
    from conftest import universe_from
    g = golden(f"rdf_{name}")
    u = universe_from(g)
    ag1, ag2 = rdf_groups(u, g)
    kw = rdf_kwargs(g)
    if name == "dropx_density":
        kw["norm"] = "density"
    r = _FakeRDF(ag1, ag2, verbose=False, **kw).run()
    assert np.array_equal(r.results.counts, g["counts"])
    np.testing.assert_allclose(r.results.rdf, g["rdf"], rtol=1e-6)
    if "edges" in g:
        assert np.array_equal(r.results.edges, g["edges"])
        assert np.array_equal(r.results.bins, g["bins"])
    np.testing.assert_allclose(r._get_rdf() if kw.get("norm", "rdf") == "rdf"
                               else r.results.rdf, r.results.rdf)


def test_results_container_and_save(tmp_path):
    h = base.Hash()
    h.x = np.arange(3)
    assert h["x"] is h.x and h.missing is None
    u = synthetic.lj_fluid(30, 2, seed=8)
    r = _FakeRDF(u.atoms, n_bins=5, range=(0.0, 1.5), verbose=False).run()
    r.results.pop("units")
    r.save(str(tmp_path / "out"))
    z = np.load(tmp_path / "out.npz")
    assert np.array_equal(z["counts"], r.results.counts)


def test_centres_of_mass_helper():
    u = synthetic.polymer_melt(5, 4, 1, seed=9)
    pos = u.trajectory[0].positions
    com = structure._centers_of_mass(u.atoms, "residues", pos)
    want = pos.reshape(5, 4, 3).astype(np.float64).mean(axis=1)
    np.testing.assert_allclose(com, want, rtol=1e-12)
    assert u.atoms.n_residues == 5


def test_ssf_host_logic_matches_golden(golden):
    from conftest import universe_from
    g = golden("sq_small")
    u = universe_from(g)
    n = int(g["n_cat"])
    cat, an = u.select(slice(0, n)), u.select(slice(n, u.atoms.n_atoms))
    for mode in (None, "pair", "partial"):
        s = FakeSSF([cat, an], mode=mode, n_points=int(g["n_points"]),
                    q_max=float(g["q_max"]), verbose=False).run()
        np.testing.assert_allclose(s.results.ssf, g[f"ssf_{mode}_exp"], rtol=1e-9,
                                   atol=1e-12)
        np.testing.assert_allclose(s.results.wavenumbers, g[f"wavenumbers_{mode}_exp"],
                                   rtol=1e-13)
    s = FakeSSF([u.atoms], n_points=6, n_surfaces=3, n_surface_points=8,
                verbose=False).run()
    np.testing.assert_allclose(s.results.ssf, g["ssf_surfaces"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(s.results.wavenumbers, g["wavenumbers_surfaces"],
                               rtol=1e-13)


def test_rdf_postprocessing_matches_reference_functions(golden):
    """calculate_coordination_numbers / calculate_structure_factor /
    radial_fourier_transform: same numbers as the reference's own functions
    (tests/golden/rdf_post.npz, produced by them)."""
    from mdhelper_b200.analysis import structure as S
    g = golden("rdf_post")
    bins, rdf, rho = g["bins"], g["rdf"], float(g["rho"])
    cn = S.calculate_coordination_numbers(bins, rdf, rho, n_coord_nums=3)
    np.testing.assert_allclose(cn, g["coordination_numbers"], rtol=1e-13)
    q, s = S.calculate_structure_factor(bins, rdf, True, rho, n_q=64)
    np.testing.assert_allclose(q, g["q_fz"], rtol=1e-14)
    np.testing.assert_allclose(s, g["ssf_fz"], rtol=1e-12, atol=1e-12)
    q, s = S.calculate_structure_factor(bins, rdf, False, rho, 0.3, 0.7, n_q=32,
                                        formalism="AL")
    np.testing.assert_allclose(s, g["ssf_al"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(
        S.radial_fourier_transform(bins, rdf - 1, np.array([0.0, 0.5, 1.0])), g["rft"],
        rtol=1e-12)
    np.testing.assert_allclose(
        np.ravel(S.zeroth_order_hankel_transform(bins, rdf - 1, np.array([0.7])))[0],
        float(g["hankel_07"]), rtol=1e-12)
    # the 2-D transform also takes wavenumber arrays here (it raises in the reference)
    h = S.zeroth_order_hankel_transform(bins, rdf - 1, np.array([0.0, 0.7]))
    np.testing.assert_allclose(h[1], float(g["hankel_07"]), rtol=1e-12)
    with pytest.raises(ValueError):
        S.calculate_structure_factor(bins, rdf, False, rho, 0.3, 0.7, formalism="xyz")
    with pytest.raises(ValueError):
        S.calculate_coordination_numbers(bins, rdf, rho, n_dims=4)


def test_rdf_class_postprocessing_methods(golden):
    """results.coordination_numbers / pmf / wavenumbers / ssf of the class, driven on
    the host with the counts of the lj1000 fixture in place of the GPU pass."""
    from _fake import FakeRDF
    from conftest import universe_from
    from mdhelper_b200.analysis import _postprocess as P
    g = golden("rdf_lj1000")
    p = golden("rdf_post")
    u = universe_from(g)
    r = FakeRDF(u.atoms, n_bins=int(g["n_bins"]), range=tuple(g["range"]),
                verbose=False).run()
    r.calculate_coordination_numbers(0.8, n_coord_nums=3)
    np.testing.assert_allclose(r.results.coordination_numbers, p["coordination_numbers"],
                               rtol=1e-12)
    r.calculate_structure_factor(0.8, n_q=64)
    np.testing.assert_allclose(r.results.ssf, p["ssf_fz"], rtol=1e-10, atol=1e-10)
    r.calculate_pmf(300.0)
    ok = r.results.rdf > 0
    np.testing.assert_allclose(r.results.pmf[ok],
                               -8.31446261815324e-3 * 300.0 * np.log(r.results.rdf[ok]),
                               rtol=1e-12)
    assert P.thermal_energy(1.5, True) == 1.5


def test_isclose_members_equals_the_reference_scan():
    """The sorted-window grouping gives exactly np.isclose's index sets (also where
    neighbouring unique wavenumbers are closer than isclose's relative tolerance)."""
    rng = np.random.default_rng(0)
    g = 2 * np.pi * np.arange(12) / np.float32(17.3)
    wv = np.stack(np.meshgrid(g, g * 1.000004, g), -1).reshape(-1, 3)
    wn = np.linalg.norm(wv, axis=1)
    uq = np.unique(wn.round(11))
    got = structure._isclose_members(uq, wn)
    assert len(got) == len(uq)
    for q, m in zip(uq, got):
        assert np.array_equal(m, np.flatnonzero(np.isclose(q, wn)))
    assert any(len(m) > 1 for m in got)


def test_feeder_stages_float64_results_of_positions_fn_unrounded():
    """positions_fn results the reference keeps in float64 (centres of mass) are staged
    as float64 when the analysis asks for it, and as float32 otherwise."""
    u = synthetic.lj_fluid(40, 5, seed=9)
    shift = 1000.0 + 1e-9

    def fn(ts):
        return [ts.positions[:10].astype(np.float64) + shift]
    for dtype in (np.float64, np.float32):
        feeder = base.FrameFeeder(u.trajectory, [np.arange(10)], np.arange(5), 2, fn,
                                  dtype=dtype)
        assert not feeder.zero_copy
        n = 0
        for b in feeder:
            assert b.f64 == (dtype is np.float64) and b.strides == [30]
            arr = b.keepalive[0][0]
            assert arr.dtype == dtype and b.ptrs[0] == arr.ctypes.data
            for k in range(b.n_frames):
                want = u.trajectory.coordinates[n + k, :10].astype(np.float64) + shift
                assert np.array_equal(arr[k], want.astype(dtype))
            n += b.n_frames
        assert n == 5


def test_feeder_accepts_mdanalysis_memoryreader_layout():
    """A reader that keeps the trajectory in memory under MDAnalysis' MemoryReader names
    (coordinate_array / dimensions_array / stored_order) is fed without copies."""
    class MemoryReaderLike:
        def __init__(self, arr, dims):
            self.coordinate_array, self.dimensions_array = arr, dims
            self.stored_order = "fac"

        def __len__(self):
            return len(self.coordinate_array)

    arr = np.random.default_rng(0).random((6, 40, 3)).astype(np.float32)
    dims = np.tile(np.array([5, 5, 5, 90, 90, 90], np.float32), (6, 1))
    t = MemoryReaderLike(arr, dims)
    f = base.FrameFeeder(t, [np.arange(10, 30)], np.arange(0, 6, 2), 2)
    assert f.zero_copy
    b = next(iter(f))
    assert b.ptrs[0] == arr.ctypes.data + 12 * 10 and b.strides[0] == 2 * 40 * 3
    assert np.array_equal(b.dims, dims[[0, 2]])
    t.stored_order = "afc"
    assert not base.FrameFeeder(t, [np.arange(10, 30)], np.arange(6), 2).zero_copy


# ---- work decomposition of the DMMA S(q) kernel (host-side planning, no device) ---------

def _sphere(n_max, n_points=None):
    r = range((n_points or n_max + 1))
    return np.array([(x, y, z) for x in r for y in r for z in r
                     if x * x + y * y + z * z <= n_max * n_max], dtype=np.int32)


@pytest.mark.parametrize("case", ["cfg4", "n20", "full32", "tiny", "column", "sparse",
                                  "anisotropic", "tall"])
def test_sq_dmma_plan_covers_every_wavevector_once(case):
    """mdh_sq_plan: every wavevector is mapped to exactly one accumulator slot, the paired
    columns obey the bank rule of the table layout, the tables fit in shared memory and the
    schedulers' loads are level (include/mdh_b200.h; kernel: csrc/sq.cu)."""
    from mdhelper_b200 import _lib
    rng = np.random.default_rng(3)
    n = {"cfg4": _sphere(16), "n20": _sphere(20), "tiny": _sphere(2),
         "full32": np.stack(np.meshgrid(*[np.arange(32)] * 3), -1).reshape(-1, 3),
         "column": np.array([(0, 0, z) for z in range(0, 45, 3)]),
         "sparse": rng.permutation(_sphere(12))[:97],
         "anisotropic": np.array([(x, y, z) for x in range(3) for y in range(11)
                                  for z in range(37)]),
         "tall": np.array([(x, 0, z) for x in range(5) for z in range(70)])}[case]
    n = rng.permutation(np.asarray(n, dtype=np.int32))          # order must not matter
    plan = _lib.sq_plan(n)
    assert np.array_equal(plan["coverage"], np.ones(len(n), np.int32))
    assert plan["pair_rule_violations"] == 0
    assert plan["tiles"] * 64 >= len(n)
    assert 1 <= plan["warps_per_block"] <= 14
    assert plan["items"] <= plan["warps_per_block"] * plan["schedulers"] // 4
    if case == "cfg4":      # the bench workload: 216 columns -> 27 groups, 49 tiles, one block
        assert (plan["tiles"], plan["items"], plan["schedulers"]) == (49, 14, 4)
        assert plan["smem_bytes"] <= 200 * 1024
    if case == "full32":    # perfect tiling: 1,024 columns x 4 tiles
        assert plan["tiles"] * 64 == 32 ** 3


def test_sq_dmma_plan_rejects_bad_indices():
    from mdhelper_b200 import _lib
    with pytest.raises(ValueError):
        _lib.sq_plan(np.array([(0, 0, 1), (0, 0, 1)]))           # duplicate
    with pytest.raises(ValueError):
        _lib.sq_plan(np.array([(0, -1, 1)]))                     # negative index


def test_grouped_mean_is_bit_identical_to_the_per_group_loop():
    """structure._grouped_mean (one numpy call per group size) == the reference's loop of
    values[..., members].mean(axis=-1) (structure.py:1538-1543), bit for bit, for the
    unique-|q| groups of a real wavevector grid and for 1-, 2- and 3-dimensional values."""
    rng = np.random.default_rng(11)
    grid = 2 * np.pi * np.arange(18) / 39.685
    wv = np.stack(np.meshgrid(grid, grid, grid), -1).reshape(-1, 3)
    w = np.linalg.norm(wv, axis=1)
    w = w[w <= grid[16] + 1e-9]
    uq = np.unique(w.round(11))
    members = structure._isclose_members(uq, w)
    assert sorted(len(m) for m in members)[-1] > 24       # long groups: pairwise summation
    for shape in ((len(w),), (3, len(w)), (5, 2, len(w))):
        v = rng.standard_normal(shape) * 10.0 ** rng.integers(-3, 4, size=shape)
        want = np.stack([v[..., m].mean(axis=-1) for m in members], axis=-1)
        got = structure._grouped_mean(v, members)
        assert got.shape == want.shape and np.array_equal(got, want)


def test_lattice_factorisation_and_3m_product_accuracy():
    """The arithmetic of sq_lattice_mma_kernel (csrc/sq.cu) restated in numpy: per-axis phase
    tables from one sincos and four interleaved recurrences E(n + 4) = E(n) E(4), the column
    products A = E_x E_y, and the 3-multiplication complex contraction
    Re = P1 - P2, Im = P3 - P1 - P2 -- against the reference's direct sum
    (accelerated.py:81-122) for N = 4,000 particles and a 12^3 lattice."""
    rng = np.random.default_rng(5)
    N, n_max, L = 4000, 11, np.array([17.0, 19.5, 23.25])
    r = (rng.random((N, 3)) * L).astype(np.float32).astype(np.float64)
    b = 2 * np.pi / L

    def table(a):
        th = b[a] * r[:, a]
        c1, s1 = np.cos(th), np.sin(th)
        c2, s2 = c1 * c1 - s1 * s1, 2.0 * (c1 * s1)
        c4, s4 = c2 * c2 - s2 * s2, 2.0 * (c2 * s2)
        e = np.zeros((n_max + 1, N), dtype=np.complex128)
        e[0] = 1.0
        e[1] = c1 + 1j * s1
        e[2] = c2 + 1j * s2
        e[3] = (e[1].real * c2 - e[1].imag * s2) + 1j * (e[1].real * s2 + e[1].imag * c2)
        for n in range(4, n_max + 1):
            e[n] = (e[n - 4].real * c4 - e[n - 4].imag * s4) \
                + 1j * (e[n - 4].real * s4 + e[n - 4].imag * c4)
        return e

    ex, ey, ez = table(0), table(1), table(2)
    a = (ex[:, None, :] * ey[None, :, :]).reshape(-1, N)          # columns (nx, ny)
    p1 = a.real @ ez.real.T
    p2 = a.imag @ ez.imag.T
    p3 = (a.real + a.imag) @ (ez.real + ez.imag).T
    rho = (p1 - p2) + 1j * (p3 - p1 - p2)                          # [(nx, ny), nz]

    n = np.stack(np.meshgrid(*[np.arange(n_max + 1)] * 3, indexing="ij"), -1).reshape(-1, 3)
    want = np.exp(1j * (n * b) @ r.T).sum(axis=1).reshape(rho.shape)
    ssf, ssf_want = np.abs(rho) ** 2, np.abs(want) ** 2
    assert np.abs(rho - want).max() < 1e-10                        # |rho| is O(sqrt(N)) ~ 60
    np.testing.assert_allclose(ssf[ssf_want > 1.0], ssf_want[ssf_want > 1.0], rtol=1e-10)


def test_sq_dmma_plan_fuzz():
    """mdh_sq_plan on 80 random wavevector sets (boxes, spheres, random subsets, nz ranges of
    up to 25 tiles): every wavevector exactly once, pairing rule kept, block shape legal."""
    from mdhelper_b200 import _lib
    rng = np.random.default_rng(123)
    for _ in range(80):
        nx, ny, nz = (int(v) for v in rng.integers(1, 30, 3))
        if rng.random() < 0.3:
            nz = int(rng.integers(1, 200))
        pts = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz),
                                   indexing="ij"), -1).reshape(-1, 3)
        mode = rng.integers(0, 3)
        if mode == 0:
            pts = pts[rng.choice(len(pts), int(rng.integers(1, len(pts) + 1)), replace=False)]
        elif mode == 1:
            radius = rng.uniform(1, max(nx, ny, nz))
            pts = pts[(pts ** 2).sum(1) <= radius * radius]
        plan = _lib.sq_plan(pts)
        assert (plan["coverage"] == 1).all()
        assert plan["pair_rule_violations"] == 0
        assert 1 <= plan["warps_per_block"] <= 14 and plan["tiles"] * 64 >= len(pts)


def test_ring_trajectory_and_feeder_runs():
    """A RingTrajectory (frame f = ring[f % period]) is fed without copies: runs never
    straddle the wrap, pointers and strides address the right ring slots."""
    import ctypes
    from mdhelper_b200.analysis.base import FrameFeeder, _is_contiguous, _runs_in_ring
    from mdhelper_b200.universe import SyntheticUniverse
    assert list(_runs_in_ring(np.arange(20), 1, 8, 5)) == [(0, 5), (5, 3), (8, 5), (13, 3),
                                                            (16, 4)]
    assert list(_runs_in_ring(np.arange(20), 1, 0, 6)) == [(0, 6), (6, 6), (12, 6), (18, 2)]
    for first, cnt in _runs_in_ring(np.arange(3, 40, 3), 3, 8, 4):
        slots = (np.arange(3, 40, 3)[first:first + cnt]) % 8
        assert np.all(np.diff(slots) == 3)
    pos = np.arange(4 * 5 * 3, dtype=np.float32).reshape(4, 5, 3)
    u = SyntheticUniverse(pos, np.array([9, 9, 9, 90, 90, 90], np.float32), n_frames=11)
    traj = u.trajectory
    assert len(traj) == 11 and traj.unitcells.shape == (11, 6)
    assert np.array_equal(traj[9].positions, pos[1]) and traj[9].frame == 9
    feeder = FrameFeeder(traj, [np.arange(1, 4)], np.arange(2, 11), 3)
    assert feeder.zero_copy
    seen = []
    for b in feeder:
        assert b.strides == [15] and b.dims.shape == (b.n_frames, 6)
        for k in range(b.n_frames):
            a = np.ctypeslib.as_array(
                (ctypes.c_float * 9).from_address(b.ptrs[0] + 4 * 15 * k)).reshape(3, 3)
            seen.append(a.copy())
    want = [pos[f % 4, 1:4] for f in range(2, 11)]
    assert len(seen) == 9 and all(np.array_equal(a, w) for a, w in zip(seen, want))
    assert _is_contiguous([3, 4, 5]) and not _is_contiguous([0, 2, 1, 3])
    assert not _is_contiguous([1, 1, 2]) and not _is_contiguous([])


def test_combined_analysis_rejects_classes_with_their_own_frame_loop():
    from mdhelper_b200.analysis import CombinedAnalysis
    from mdhelper_b200.analysis.structure import (IntermediateScatteringFunction,
                                                  RadialDistributionFunction)
    from mdhelper_b200.universe import SyntheticUniverse
    pos = np.random.default_rng(0).random((3, 40, 3)).astype(np.float32) * 5
    u = SyntheticUniverse(pos, np.array([5, 5, 5, 90, 90, 90], np.float32))
    rdf = RadialDistributionFunction(u.atoms, n_bins=10, range=(0.0, 2.0), verbose=False)
    isf = IntermediateScatteringFunction([u.atoms], n_points=3, verbose=False)
    with pytest.raises(TypeError):
        CombinedAnalysis(rdf, isf)
    CombinedAnalysis(rdf)                      # the plain classes are accepted


def test_staged_feeder_reads_an_on_disk_style_reader(tmp_path):
    """A reader without a whole-trajectory array (frames seeked one by one from a file):
    the feeder gathers each batch into its staging buffers -- right frames, right atoms,
    right cells, two alternating buffers."""
    import ctypes
    from conftest import file_universe
    from mdhelper_b200.analysis.base import FrameFeeder
    rng = np.random.default_rng(3)
    pos = rng.random((7, 30, 3)).astype(np.float32)
    dims = np.concatenate([rng.uniform(5, 6, (7, 3)), np.full((7, 3), 90.0)], 1).astype(np.float32)
    u = file_universe(tmp_path, pos, dims)
    ix = [np.array([2, 3, 4, 9]), np.arange(10, 30)]
    feeder = FrameFeeder(u.trajectory, ix, np.array([6, 1, 3, 4, 0]), 2)
    assert not feeder.zero_copy
    seen = 0
    for b, frames in zip(feeder, ([6, 1], [3, 4], [0])):
        assert b.n_frames == len(frames) and b.strides == [12, 60]
        np.testing.assert_array_equal(b.dims, dims[frames])
        for k, f in enumerate(frames):
            for s, sel in enumerate(ix):
                n = len(sel)
                a = np.ctypeslib.as_array((ctypes.c_float * (3 * n)).from_address(
                    b.ptrs[s] + 4 * b.strides[s] * k)).reshape(n, 3)
                np.testing.assert_array_equal(a, pos[f][sel])
        seen += b.n_frames
    assert seen == 5 and u.trajectory.reads >= 5


def test_staging_plan_ramp_and_tail():
    """How host batches are cut into pieces (mdh_stage_plan): a short first copy, pieces
    doubling up to ~32 MB, every frame exactly once; a remainder below half a piece joins
    the last piece unless the measured copy/kernel ratio says its copy would outlast the
    kernels before it -- then it is cut in halves."""
    from mdhelper_b200 import _lib
    rng = np.random.default_rng(11)
    for _ in range(300):
        n = int(rng.integers(1, 3000))
        bpf = float(10 ** rng.uniform(2, 7.5))
        ratio = float(rng.choice([0.0, 0.1, 0.3, 0.6, 1.5]))
        p = _lib.stage_plan(n, bpf, ratio)
        assert sum(p) == n and min(p) >= 1
        assert p[0] < 1.5 * max(1, int(2e6 // bpf)) + 1          # the copy nothing can hide
        cap = max(1, int(32e6 // bpf))
        assert max(p) < 1.5 * cap + 1
        for a, b in zip(p[:-2], p[1:-1]):          # the ramp never more than doubles (+ rounding)
            assert b <= 2 * a + 1
    # cfg4's 125-frame call (600 kB per frame): merged tail by default and on a fast link,
    # halved tail when the copy is slow against the kernels
    assert _lib.stage_plan(125, 600e3) == [3, 6, 13, 26, 77]
    assert _lib.stage_plan(125, 600e3, 0.25) == [3, 6, 13, 26, 77]
    assert _lib.stage_plan(125, 600e3, 0.54) == [3, 6, 13, 26, 39, 38]
    assert _lib.stage_plan(0, 600e3) == []
