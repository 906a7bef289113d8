"""GPU: pair-histogram kernels through the C ABI vs the oracle / golden fixtures.
Integer bin counts must be bit-exact."""
import numpy as np
import pytest

from conftest import universe_from
from test_oracle import RDF_CASES, rdf_groups, rdf_kwargs

pytestmark = pytest.mark.gpu

HISTS = ["warp_atomic", "lane_private"]
# (hist, arith): the fp32-filter kernel, the exact fp64 kernel with either histogram
VARIANTS = [("warp_atomic", "auto"), ("warp_atomic", "off"), ("lane_private", "off")]


def _structure():
    from mdhelper_b200.analysis import structure
    return structure


def _oracle():
    from oracle import reference_port as rp
    return rp


def test_kat_radial_histogram(golden):
    """The reference's own known-answer test (tests/test_analysis_structure.py:21-40)."""
    g = golden("kat_radial_histogram")
    for hist, arith in VARIANTS + [("warp_atomic", "audit")]:
        st = {}
        got = _structure().radial_histogram(
            g["origin"], g["neighbors"], int(g["n_bins"]), tuple(g["range"]), g["dims"],
            hist=hist, arith=arith, stats=st)
        assert np.array_equal(got, g["expected_from_norms"])
        assert np.array_equal(got, g["reference_counts"])
        assert st["audit_violations"] == 0


@pytest.mark.parametrize("hist,arith", VARIANTS)
@pytest.mark.parametrize("name", RDF_CASES)
def test_class_matches_golden(golden, name, hist, arith):
    g = golden(f"rdf_{name}")
    u = universe_from(g)
    ag1, ag2 = rdf_groups(u, g)
    kw = rdf_kwargs(g)
    if name == "dropx_density":
        kw["norm"] = "density"
    r = _structure().RadialDistributionFunction(
        ag1, ag2, verbose=False, mode="allpairs", hist=hist, arith=arith, **kw).run()
    if arith == "auto":
        # every golden configuration is eligible for the filter
        assert r._filter_stats["eligible"] == 1
    assert r.results.counts.dtype == np.int64 or r.results.counts.dtype == int
    assert np.array_equal(r.results.counts, g["counts"])
    np.testing.assert_allclose(r.results.rdf, g["rdf"], rtol=1e-6)


@pytest.mark.parametrize("hist,arith", VARIANTS)
@pytest.mark.parametrize("name", ["lj1000", "twogroup", "excl11", "excl410",
                                  "noncubic_npt", "unwrapped"])
def test_cells_mode_matches_golden(golden, name, hist, arith):
    g = golden(f"rdf_{name}")
    u = universe_from(g)
    ag1, ag2 = rdf_groups(u, g)
    kw = rdf_kwargs(g)
    # the cell list needs >= 3 cells per axis: shorten the range where necessary
    # and compare with the oracle instead of the stored counts
    lmin = float(np.atleast_2d(g["dims"])[:, :3].min())
    if kw["range"][1] * 1.00001 * 3 > lmin:
        kw["range"] = (kw["range"][0], np.float32(lmin / 3.2).item())
        want = _oracle().rdf_run(u, ag1, ag2, **kw)["counts"]
    else:
        want = g["counts"]
    r = _structure().RadialDistributionFunction(
        ag1, ag2, verbose=False, mode="cells", hist=hist, arith=arith, **kw).run()
    assert np.array_equal(r.results.counts, want)
    if arith == "auto":
        assert r._filter_stats["eligible"] == 1


@pytest.mark.parametrize("n1,n2", [(1, 1), (1, 700), (513, 511), (512, 1024), (33, 1537)])
def test_ragged_sizes(n1, n2):
    rng = np.random.default_rng(n1 * 10007 + n2)
    dims = np.array([7.5, 8.25, 9.0, 90, 90, 90], np.float32)
    p1 = (rng.random((n1, 3)) * dims[:3]).astype(np.float32)
    p2 = (rng.random((n2, 3)) * dims[:3]).astype(np.float32)
    want = _oracle().radial_histogram(p1, p2, 47, (0.0, 3.7), dims)
    for hist, arith in VARIANTS:
        got = _structure().radial_histogram(p1, p2, 47, (0.0, 3.7), dims, hist=hist,
                                            arith=arith)
        assert np.array_equal(got, want)


def test_many_bins_falls_back_to_warp_atomics():
    rng = np.random.default_rng(5)
    dims = np.array([12, 12, 12, 90, 90, 90], np.float32)
    p = (rng.random((900, 3)) * 12).astype(np.float32)
    want = _oracle().radial_histogram(p, p, 3000, (0.0, 6.0), dims)
    got = _structure().radial_histogram(p, p, 3000, (0.0, 6.0), dims)
    assert np.array_equal(got, want)


# ---- the fp32 filter (rdf_filter.cu): audited against the exact arithmetic ----------

def _audit(p1, p2, n_bins, rng_, dims, exclusion=None):
    """Runs the filter kernel with the audit on; returns (counts, stats)."""
    st = {}
    got = _structure().radial_histogram(p1, p2, n_bins, rng_, dims, exclusion=exclusion,
                                        arith="audit", stats=st)
    assert st["eligible"] == 1
    assert st["audit_violations"] == 0, st
    return got, st


@pytest.mark.parametrize("n_bins,rng_", [(201, (0.0, 5.0)), (64, (0.0, 5.4)),
                                         (100, (1.0, 4.5)), (1000, (0.0, 2.5)),
                                         (3000, (0.0, 5.0)), (37, (2.25, 2.75))])
def test_filter_audit_random(n_bins, rng_):
    """Every pair the filter calls certain agrees with the fp64 arithmetic; counts are
    the oracle's (uniform random coordinates, non-cubic box, ranges with r_lo > 0)."""
    rng = np.random.default_rng(n_bins)
    dims = np.array([10.75, 11.5, 12.25, 90, 90, 90], np.float32)
    p1 = (rng.random((1500, 3)) * dims[:3]).astype(np.float32)
    p2 = (rng.random((2100, 3)) * dims[:3]).astype(np.float32)
    want = _oracle().radial_histogram(p1, p2, n_bins, rng_, dims)
    got, st = _audit(p1, p2, n_bins, rng_, dims)
    assert np.array_equal(got, want)
    # the uncertainty window is a small part of a bin
    assert st["audit_uncertain_pairs"] < 0.02 * len(p1) * len(p2)
    off = _structure().radial_histogram(p1, p2, n_bins, rng_, dims, arith="off")
    assert np.array_equal(off, want)


def test_filter_audit_exclusions_and_unwrapped():
    rng = np.random.default_rng(77)
    dims = np.array([9.5, 9.5, 9.5, 90, 90, 90], np.float32)
    # coordinates up to three boxes outside the cell
    p = ((rng.random((1800, 3)) * 7 - 3) * dims[:3]).astype(np.float32)
    want = _oracle().radial_histogram(p, p, 120, (0.0, 4.75), dims, exclusion=(3, 3))
    got, _ = _audit(p, p, 120, (0.0, 4.75), dims, exclusion=(3, 3))
    assert np.array_equal(got, want)


def test_filter_adversarial_lattice_overflows_deferred_list():
    """Simple-cubic lattice whose neighbour distances sit exactly on bin edges: most
    pairs are uncertain, the deferred lists overflow into the inline path, and the
    counts still equal the oracle's."""
    n = 16
    a = 0.5
    g = np.arange(n, dtype=np.float32) * a
    p = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    dims = np.array([n * a, n * a, n * a, 90, 90, 90], np.float32)
    want = _oracle().radial_histogram(p, p, 16, (0.0, 4.0), dims)
    st = {}
    got = _structure().radial_histogram(p, p, 16, (0.0, 4.0), dims, stats=st)
    assert np.array_equal(got, want)
    assert st["eligible"] == 1 and st["deferred_entries"] > 0
    assert st["inline_entries"] > 0
    got, _ = _audit(p, p, 16, (0.0, 4.0), dims)
    assert np.array_equal(got, want)


def test_filter_declines_frames_far_outside_the_box():
    """Coordinates millions of boxes away: the error bound is useless, the filter
    kernel hands the frame to the exact kernel, counts stay the oracle's."""
    rng = np.random.default_rng(3)
    dims = np.array([6.0, 6.0, 6.0, 90, 90, 90], np.float32)
    p = (rng.random((700, 3)) * 6).astype(np.float32)
    q = p.copy()
    q[::2, 0] += np.float32(6.0 * 2 ** 21)
    want = _oracle().radial_histogram(q, q, 50, (0.0, 3.0), dims)
    st = {}
    got = _structure().radial_histogram(q, q, 50, (0.0, 3.0), dims, stats=st)
    assert np.array_equal(got, want)
    assert st["declined_frames"] == 1
    # non-finite coordinates are declined as well (and never counted)
    q = p.copy()
    q[5] = np.nan
    want = _oracle().radial_histogram(q, q, 50, (0.0, 3.0), dims)
    st = {}
    got = _structure().radial_histogram(q, q, 50, (0.0, 3.0), dims, stats=st)
    assert np.array_equal(got, want)
    assert st["declined_frames"] == 1


@pytest.mark.parametrize("same", [True, False])
@pytest.mark.parametrize("exclusion", [None, (2, 2)])
def test_cells_filter_audit(same, exclusion):
    """The cell-list kernel with the fp32 filter, audited pair by pair against the fp64
    arithmetic: half stencil (same group) and full stencil (two groups), odd particle
    counts (the pair layout's filler slot), exclusions, r_lo > 0."""
    rng = np.random.default_rng(41 + same)
    dims = np.array([21.0, 23.5, 22.25, 90, 90, 90], np.float32)
    p1 = (rng.random((6001, 3)) * dims[:3]).astype(np.float32)
    p2 = p1 if same else (rng.random((7003, 3)) * dims[:3]).astype(np.float32)
    S = _structure()
    for n_bins, rng_ in [(100, (0.0, 2.5)), (75, (0.5, 2.75))]:
        want = _oracle().radial_histogram(p1, p2, n_bins, rng_, dims, exclusion=exclusion)
        st = {}
        got = S.radial_histogram(p1, p2, n_bins, rng_, dims, exclusion=exclusion,
                                 mode="cells", arith="audit", stats=st)
        assert st["eligible"] == 1 and st["audit_violations"] == 0, st
        assert st["deferred_entries"] > 0
        assert np.array_equal(got, want)
        for arith in ("auto", "off"):
            got = S.radial_histogram(p1, p2, n_bins, rng_, dims, exclusion=exclusion,
                                     mode="cells", arith=arith)
            assert np.array_equal(got, want)


def test_cells_dense_clusters_take_the_long_list_path():
    """Clusters far denser than the average: cells with more particles than one pass deals
    to the lanes (several particle chunks) and candidate lists longer than the per-warp
    buffer (read from global memory instead); same group, two groups, exclusions, audit."""
    rng = np.random.default_rng(2026)
    dims = np.array([30.0, 27.0, 33.0, 90, 90, 90], np.float32)
    p = (rng.random((5000, 3)) * dims[:3]).astype(np.float32)
    p[:1400] = (np.array([7.0, 8.0, 9.0]) + 0.8 * rng.standard_normal((1400, 3))).astype(np.float32)
    p[1400:1500] = (np.array([29.9, 0.1, 16.0])
                    + 0.2 * rng.standard_normal((100, 3))).astype(np.float32)   # across a face
    q = (rng.random((3100, 3)) * dims[:3]).astype(np.float32)
    q[:600] = (np.array([7.5, 8.5, 9.5]) + 0.5 * rng.standard_normal((600, 3))).astype(np.float32)
    S = _structure()
    for p2, exclusion in [(p, None), (p, (4, 4)), (q, None), (q, (2, 3))]:
        # a few cluster particles lie outside [0, L): both sides follow the reference's
        # method choice (grid search here: coordinates moved into the cell in float32)
        want = _oracle().radial_histogram(p, p2, 90, (0.0, 3.0), dims, exclusion=exclusion)
        for arith in ("auto", "audit", "off"):
            st = {}
            got = S.radial_histogram(p, p2, 90, (0.0, 3.0), dims, exclusion=exclusion,
                                     mode="cells", arith=arith, stats=st)
            assert np.array_equal(got, want), (exclusion, arith)
            assert st["audit_violations"] == 0
            if arith != "off":
                assert st["eligible"] == 1 and st["declined_frames"] == 0


@pytest.mark.parametrize("tune", ["cipt=2", "cchunk=1", "cchunk=64,cws=1"])
def test_cells_kernel_tunables_do_not_change_counts(tune, monkeypatch):
    """Two particles per lane instead of four, one cell or 64 cells per work item, one
    frame per sort group: identical counts."""
    from mdhelper_b200 import synthetic
    u = synthetic.lj_fluid(20_000, 3, seed=5)
    S = _structure()
    kw = dict(n_bins=100, range=(0.0, 2.5), norm=None, verbose=False, mode="cells")
    ref = S.RadialDistributionFunction(u.atoms, **kw).run().results.counts
    monkeypatch.setenv("MDH_TUNE", tune)
    got = S.RadialDistributionFunction(u.atoms, **kw).run().results.counts
    assert np.array_equal(got, ref)
    want = _oracle().rdf_run(u, u.atoms, n_bins=100, range=(0.0, 2.5), norm=None)["counts"]
    assert np.array_equal(ref, want)


def test_cells_minimal_grid_and_changing_boxes():
    """Three cells per axis (every stencil offset wraps) and a box that changes from
    frame to frame (the grid is rebuilt per frame)."""
    from mdhelper_b200.universe import SyntheticUniverse
    rng = np.random.default_rng(12)
    F, n = 5, 2500
    edges = np.array([[9.1, 9.4, 9.7], [9.3, 9.2, 10.4], [12.9, 9.05, 9.6], [9.02, 9.01, 9.03],
                      [15.5, 12.5, 9.9]], np.float32)
    pos = (rng.random((F, n, 3)) * edges[:, None, :]).astype(np.float32)
    dims = np.concatenate([edges, np.full((F, 3), 90, np.float32)], axis=1)
    u = SyntheticUniverse(pos, dims)
    S = _structure()
    for sel, excl in [((u.atoms, None), None), ((u.select(slice(0, 900)),
                                                 u.select(slice(900, n))), None),
                      ((u.atoms, None), (5, 5))]:
        kw = dict(n_bins=60, range=(0.0, 3.0), norm=None, exclusion=excl)
        want = _oracle().rdf_run(u, sel[0], sel[1], method="bruteforce", **kw)["counts"]
        for arith in ("auto", "off", "audit"):
            r = S.RadialDistributionFunction(sel[0], sel[1], verbose=False, mode="cells",
                                             arith=arith, **kw).run()
            assert np.array_equal(r.results.counts, want)
            assert r._filter_stats["audit_violations"] == 0


@pytest.mark.parametrize("exclusion", [(2, 3), (3, 1)])
def test_same_group_with_unequal_exclusion_blocks(exclusion):
    """ag1 is ag2 with exclusion[0] != exclusion[1]: i // e0 == j // e1 is not symmetric,
    so the pair symmetry must not be used (structure.py:100-102)."""
    from mdhelper_b200 import synthetic
    u = synthetic.lj_fluid(1200, 2, seed=17)
    L = float(u.dimensions[0])
    S = _structure()
    kw = dict(n_bins=50, range=(0.0, L / 2), norm=None, exclusion=exclusion)
    want = _oracle().rdf_run(u, u.atoms, **kw)["counts"]
    for mode, rng_ in [("allpairs", (0.0, L / 2)), ("cells", (0.0, L / 3.3))]:
        kw["range"] = rng_
        want = _oracle().rdf_run(u, u.atoms, **kw)["counts"]
        for arith in ("auto", "off", "audit"):
            r = S.RadialDistributionFunction(u.atoms, verbose=False, mode=mode, arith=arith,
                                             **kw).run()
            assert np.array_equal(r.results.counts, want), (mode, arith)
    p = u.trajectory.coordinates[0]
    got = S.radial_histogram(p, p, 50, (0.0, L / 2), u.dimensions, exclusion=exclusion)
    assert np.array_equal(got, _oracle().radial_histogram(p, p, 50, (0.0, L / 2), u.dimensions,
                                                          exclusion=exclusion))
    from mdhelper_b200 import _lib
    ctx = _lib.Context(0)
    with pytest.raises(ValueError):        # the C ABI refuses the unsound combination
        ctx.rdf_configure(8, 8, True, np.array([0.0, 1.0, 4.0]), 0.0, 2.0, exclusion=(2, 3))
    ctx.close()


def test_same_group_symmetry_and_self_pairs():
    """ag1 is ag2: ordered pairs, N self pairs in bin 0."""
    from mdhelper_b200 import synthetic
    u = synthetic.lj_fluid(1500, 2, seed=11)
    L = float(u.dimensions[0])
    S = _structure()
    a = S.RadialDistributionFunction(u.atoms, n_bins=60, range=(0.0, L / 2),
                                     norm=None, verbose=False).run()
    b = S.RadialDistributionFunction(u.atoms, u.select(np.arange(1500)), n_bins=60,
                                     range=(0.0, L / 2), norm=None, verbose=False).run()
    assert np.array_equal(a.results.counts, b.results.counts)
    want = _oracle().rdf_run(u, u.atoms, n_bins=60, range=(0.0, L / 2), norm=None)
    assert np.array_equal(a.results.counts, want["counts"])
    assert a.results.counts[0] >= 2 * 1500     # the self pairs, both frames
    # swapping the groups of a two-group RDF cannot change the counts
    g1, g2 = u.select(slice(0, 400)), u.select(slice(400, 1500))
    c12 = S.RadialDistributionFunction(g1, g2, n_bins=60, range=(0.0, L / 2),
                                       norm=None, verbose=False).run().results.counts
    c21 = S.RadialDistributionFunction(g2, g1, n_bins=60, range=(0.0, L / 2),
                                       norm=None, verbose=False).run().results.counts
    assert np.array_equal(c12, c21)


def test_frame_selection_batching_and_staged_feeder():
    from mdhelper_b200 import synthetic
    u = synthetic.lj_fluid(700, 9, seed=12)
    S, rp = _structure(), _oracle()
    kw = dict(n_bins=40, range=(0.0, 4.0))
    r = S.RadialDistributionFunction(u.atoms, verbose=False, batch_frames=2, **kw).run(
        start=1, stop=9, step=3)
    want = rp.rdf_run(u, u.atoms, frames=[1, 4, 7], **kw)
    assert np.array_equal(r.results.counts, want["counts"])
    np.testing.assert_allclose(r.results.rdf, want["rdf"], rtol=1e-6)
    # scattered selection -> staged (gather) feeder; explicit frame list
    ix = np.random.default_rng(1).permutation(700)[:300]
    grp = u.select(ix)
    r = S.RadialDistributionFunction(grp, verbose=False, batch_frames=2, **kw).run(
        frames=[0, 2, 3, 8])
    want = rp.rdf_run(u, grp, frames=[0, 2, 3, 8], **kw)
    assert np.array_equal(r.results.counts, want["counts"])


def test_centre_of_mass_groupings():
    from mdhelper_b200 import synthetic
    u = synthetic.polymer_melt(40, 5, 2, seed=13)
    S, rp = _structure(), _oracle()
    L = float(u.dimensions[0])
    r = S.RadialDistributionFunction(u.atoms, groupings="residues", n_bins=20,
                                     range=(0.0, L / 2), verbose=False).run()
    counts = np.zeros(20, dtype=np.int64)
    for f in range(2):
        ts = u.trajectory[f]
        com = S._centers_of_mass(u.atoms, "residues", ts.positions)
        counts += rp.radial_histogram(com, com, 20, (0.0, L / 2), ts.dimensions)
    assert np.array_equal(r.results.counts, counts)


def test_device_centres_of_mass_equal_host():
    """groupings on the device (com.cu) vs the host helper: identical float32 centres,
    hence identical counts -- residues x residues, residues x atoms (identity plan for
    the atoms side), unequal residue sizes, staged (strided frames) input."""
    from mdhelper_b200.universe import SyntheticUniverse
    rng = np.random.default_rng(21)
    n_res = 900
    sizes = rng.integers(1, 7, n_res)
    res = np.repeat(np.arange(n_res), sizes)
    n = res.size
    dims = np.array([14.0, 15.0, 16.0, 90, 90, 90], np.float32)
    pos = (rng.random((5, n, 3)) * dims[:3]).astype(np.float32)
    u = SyntheticUniverse(pos, dims, resindices=res, segindices=res // 3,
                          masses=rng.uniform(1.0, 40.0, n))
    S = _structure()
    half = int(np.searchsorted(res, n_res // 2))
    g1, g2 = u.select(slice(0, half)), u.select(slice(half, n))
    cases = [(u.atoms, None, "residues"), (u.atoms, None, "segments"),
             (g1, g2, ("residues", "atoms")), (g1, g2, ("residues", "residues"))]
    for a, b, grp in cases:
        kw = dict(n_bins=60, range=(0.0, 6.0), groupings=grp, norm=None, verbose=False)
        dev = S.RadialDistributionFunction(a, b, batch_frames=2, **kw).run()
        assert dev._com is not None                         # the device path was taken
        host = S.RadialDistributionFunction(a, b, host_com=True, **kw).run()
        assert host._com is None
        assert np.array_equal(dev.results.counts, host.results.counts)
        assert dev.results.counts.sum() > 0
    # scattered atoms: no device plan, host helper
    sc = u.select(rng.permutation(n)[: n // 2])
    r = S.RadialDistributionFunction(sc, groupings="residues", n_bins=20,
                                     range=(0.0, 5.0), verbose=False).run()
    assert r._com is None


def _com_universe(g, tag):
    from mdhelper_b200.universe import SyntheticUniverse
    return SyntheticUniverse(g[f"{tag}_positions"], g[f"{tag}_dims"],
                             resindices=g[f"{tag}_resindices"],
                             segindices=g[f"{tag}_segindices"], masses=g[f"{tag}_masses"])


@pytest.mark.parametrize("tag", ["equal", "unequal"])
@pytest.mark.parametrize("host_com", [False, True])
def test_groupings_against_the_reference_centres_of_mass(golden, tag, host_com):
    """groupings="residues"/"segments": counts of the reference's own class running on its
    own ``center_of_mass`` (algorithm/molecule.py:15-310; tests/golden/com_ref.npz), with
    the device kernel (com.cu) and with the host helper."""
    g = golden("com_ref")
    u = _com_universe(g, tag)
    n_a, n = int(g[f"{tag}_n_a"]), u.atoms.n_atoms
    a, b = u.select(slice(0, n_a)), u.select(slice(n_a, n))
    S = _structure()
    kw = dict(n_bins=40, range=(0.0, 6.0), verbose=False, host_com=host_com)
    for sel, grp, key in [((a, b), "residues", "res"), ((u.atoms, None), "segments", "seg"),
                          ((a, b), ("residues", "atoms"), "mix")]:
        r = S.RadialDistributionFunction(sel[0], sel[1], groupings=grp, **kw).run()
        assert (r._com is None) == host_com
        assert np.array_equal(r.results.counts, g[f"{tag}_rdf_{key}_counts"]), (tag, key)
        np.testing.assert_allclose(r.results.rdf, g[f"{tag}_rdf_{key}"], rtol=1e-6)


def test_config2_sized_frame_against_oracle():
    """One frame of the bench workload (10k x 10k two-group) vs the CPU oracle."""
    from mdhelper_b200 import synthetic
    u, cat, an = synthetic.electrolyte(20_000, 1, seed=20260002)
    S, rp = _structure(), _oracle()
    kw = dict(n_bins=201, range=(0.0, 14.5))
    r = S.RadialDistributionFunction(cat, an, verbose=False, **kw).run()
    want = rp.rdf_run(u, cat, an, method="bruteforce", **kw)
    assert np.array_equal(r.results.counts, want["counts"])
    assert r._filter_stats["eligible"] == 1 and r._filter_stats["declined_frames"] == 0
    for arith in ("off", "audit"):
        x = S.RadialDistributionFunction(cat, an, verbose=False, arith=arith, **kw).run()
        assert np.array_equal(x.results.counts, want["counts"])
        assert x._filter_stats["audit_violations"] == 0
        if arith == "audit":      # about one pair in a thousand needs the fp64 path
            assert 0 < x._filter_stats["audit_uncertain_pairs"] < 4e5
    np.testing.assert_allclose(r.results.rdf, want["rdf"], rtol=1e-6)
    # pair conservation at full range: every ordered pair is within L*sqrt(3)/2
    L = float(u.dimensions[0])
    full = S.RadialDistributionFunction(cat, an, n_bins=64, range=(0.0, L),
                                        norm=None, verbose=False).run()
    assert full.results.counts.sum() == cat.n_atoms * an.n_atoms


def test_large_cutoff_run_modes_agree():
    """Size-independent property at a config-3-like scale: the all-pairs and the
    cell-list kernels (and both histogram schemes) give identical counts."""
    from mdhelper_b200 import synthetic
    u = synthetic.lj_fluid(60_000, 2, seed=20260003)
    S = _structure()
    kw = dict(n_bins=100, range=(0.0, 2.5), norm=None, verbose=False)
    ref = S.RadialDistributionFunction(u.atoms, mode="allpairs", **kw).run().results.counts
    for hist in HISTS:
        c = S.RadialDistributionFunction(u.atoms, mode="cells", hist=hist,
                                         **kw).run().results.counts
        assert np.array_equal(c, ref)
    auto = S.RadialDistributionFunction(u.atoms, **kw).run().results.counts
    assert np.array_equal(auto, ref)
    for arith in ("off", "audit"):
        x = S.RadialDistributionFunction(u.atoms, mode="cells", arith=arith, **kw).run()
        assert np.array_equal(x.results.counts, ref)
        assert x._filter_stats["audit_violations"] == 0
    # one frame against the oracle's grid search
    want = _oracle().rdf_run(u, u.atoms, n_bins=100, range=(0.0, 2.5), norm=None,
                             frames=[0], method="nsgrid")["counts"]
    one = S.RadialDistributionFunction(u.atoms, **kw).run(stop=1).results.counts
    assert np.array_equal(one, want)


@pytest.mark.parametrize("which", ["cfg3", "cfg5"])
def test_full_size_cutoff_frames_against_oracle(which):
    """One frame at the size of BASELINE configs 3 and 5 (500,000-particle LJ fluid,
    1,000,000-bead melt; cut-off 2.5) against the oracle's grid search, plus the fp64
    kernel and the audited filter on the same frame."""
    from mdhelper_b200 import synthetic
    if which == "cfg3":
        u = synthetic.lj_fluid(500_000, 1, seed=20260003)
    else:
        u = synthetic.polymer_melt(10_000, 100, 1, seed=20260005)
    S = _structure()
    kw = dict(n_bins=100, range=(0.0, 2.5), norm=None)
    want = _oracle().rdf_run(u, u.atoms, method="nsgrid", **kw)["counts"]
    r = S.RadialDistributionFunction(u.atoms, verbose=False, **kw).run()
    assert np.array_equal(r.results.counts, want)
    assert r._filter_stats["eligible"] == 1 and r._filter_stats["declined_frames"] == 0
    for arith in ("off", "audit"):
        x = S.RadialDistributionFunction(u.atoms, verbose=False, arith=arith, **kw).run()
        assert np.array_equal(x.results.counts, want)
        assert x._filter_stats["audit_violations"] == 0


def test_host_batches_in_overlapped_pieces():
    """A host batch of more than ~24 MB is cut into pieces whose copies overlap the
    kernels of the previous piece (HostStager): same counts as small batches, for an
    uneven split (103 frames -> 2 pieces of 52 + 51) and the device-resident path."""
    import torch
    from mdhelper_b200 import _lib, synthetic
    from mdhelper_b200.analysis._binning import squared_thresholds
    u, cat, an = synthetic.electrolyte(20_000, 103, seed=99)
    S = _structure()
    kw = dict(n_bins=64, range=(0.0, 7.0), norm=None, verbose=False)
    big = S.RadialDistributionFunction(cat, an, batch_frames=103, **kw).run()
    small = S.RadialDistributionFunction(cat, an, batch_frames=10, **kw).run()
    assert np.array_equal(big.results.counts, small.results.counts)
    ctx = _lib.Context(0)
    n1, N = cat.n_atoms, 20_000
    ctx.rdf_configure(n1, N - n1, False, squared_thresholds(64, (0.0, 7.0)), 0.0, 7.0)
    dev = torch.from_numpy(u.trajectory.coordinates).cuda()
    boxes = np.ascontiguousarray(u.trajectory.unitcells[:, :3])
    ctx.rdf_accumulate(dev.data_ptr(), 3 * N, dev.data_ptr() + 12 * n1, 3 * N, boxes, 103,
                       device=True)
    assert np.array_equal(ctx.rdf_fetch(), big.results.counts)
    ctx.close()


def test_argument_errors():
    from mdhelper_b200 import _lib
    ctx = _lib.Context(0)
    with pytest.raises(RuntimeError):      # accumulate before configure
        ctx.rdf_accumulate(np.zeros((2, 3), np.float32), 6, None, 0,
                           np.ones((1, 3), np.float32), 1)
    with pytest.raises(ValueError):
        ctx.rdf_configure(4, 5, True, np.array([0.0, 1.0, 4.0]), 0.0, 2.0)
    with pytest.raises(ValueError):
        ctx.rdf_configure(4, 4, True, np.array([0.0, 4.0, 1.0]), 0.0, 2.0)
    ctx.rdf_configure(4, 4, True, np.array([0.0, 1.0, 4.0]), 0.0, 2.0)
    with pytest.raises(ValueError):        # non-positive box edge
        ctx.rdf_accumulate(np.zeros((4, 3), np.float32), 12, None, 0,
                           np.zeros((1, 3), np.float32), 1)
    with pytest.raises(ValueError):        # angles that do not span a cell
        _structure().radial_histogram(np.zeros((2, 3)), np.zeros((2, 3)), 4, (0, 1),
                                      (5, 5, 5, 30, 40, 120))
    ctx.close()


def _triclinic_universe(n=900, F=3, seed=4):
    from mdhelper_b200.universe import SyntheticUniverse
    rng = np.random.default_rng(seed)
    dims = np.array([[10.5, 11.25, 12.0, 75.0, 82.0, 64.0],
                     [10.0, 11.5, 12.5, 90.0, 90.0, 70.0],
                     [11.0, 11.0, 11.0, 60.0, 60.0, 60.0]], np.float32)[:F]
    pos = np.empty((F, n, 3), np.float32)
    for f in range(F):
        h = _oracle().triclinic_vectors(dims[f]).astype(np.float64)
        frac = rng.random((n, 3)) * (1 if f != 1 else 3) - (0 if f != 1 else 1)   # frame 1 unwrapped
        pos[f] = (frac @ h).astype(np.float32)
    return SyntheticUniverse(pos, dims)


def test_triclinic_cells_against_the_oracle():
    """Triclinic frames (wrap into the cell + shortest of 27 images in fp64): counts equal
    the restated triclinic path of capped_distance bit for bit -- same group, two groups,
    exclusions, coordinates outside the cell, one cell per frame."""
    u = _triclinic_universe()
    S = _structure()
    a, b = u.select(slice(0, 350)), u.select(slice(350, 900))
    for sel, excl in [((u.atoms, None), None), ((a, b), None), ((u.atoms, None), (3, 3)),
                      ((a, b), (2, 5))]:
        kw = dict(n_bins=70, range=(0.0, 4.5), exclusion=excl)
        want = _oracle().rdf_run(u, sel[0], sel[1], **kw)
        r = S.RadialDistributionFunction(sel[0], sel[1], verbose=False, **kw).run()
        assert np.array_equal(r.results.counts, want["counts"])
        np.testing.assert_allclose(r.results.rdf, want["rdf"], rtol=1e-6)
    p = u.trajectory.coordinates[0]
    got = S.radial_histogram(p, p, 40, (0.5, 4.0), u.trajectory.unitcells[0])
    assert np.array_equal(got, _oracle().radial_histogram(p, p, 40, (0.5, 4.0),
                                                          u.trajectory.unitcells[0]))


def test_triclinic_kernel_with_right_angles_equals_orthorhombic_kernels():
    """Property: a diagonal cell matrix through the triclinic kernel gives the counts of
    the orthorhombic kernels; mixed trajectories route every frame to its own kernel."""
    from mdhelper_b200 import _lib
    from mdhelper_b200.analysis._binning import squared_thresholds
    from mdhelper_b200.universe import SyntheticUniverse
    rng = np.random.default_rng(10)
    dims = np.array([10.0, 11.0, 12.0, 90, 90, 90], np.float32)
    p = (rng.random((2, 700, 3)) * dims[:3] * 2 - dims[:3] / 2).astype(np.float32)
    S = _structure()
    want = sum(S.radial_histogram(p[f], p[f], 64, (0.0, 5.0), dims) for f in range(2))
    ctx = _lib.Context(0)
    ctx.rdf_configure(700, 700, True, squared_thresholds(64, (0.0, 5.0)), 0.0, 5.0)
    cell = np.tile(np.diag(dims[:3])[None], (2, 1, 1)).astype(np.float32)
    ctx.rdf_accumulate_triclinic(p, 2100, p, 2100, cell, 2)
    assert np.array_equal(ctx.rdf_fetch(), want)
    ctx.close()
    # frames 0 and 2 orthorhombic, frame 1 triclinic
    tri = _triclinic_universe(n=700, F=1)
    pos = np.stack([p[0], tri.trajectory.coordinates[0], p[1]])
    cells = np.stack([dims, tri.trajectory.unitcells[0], dims])
    u = SyntheticUniverse(pos, cells)
    r = S.RadialDistributionFunction(u.atoms, n_bins=64, range=(0.0, 5.0), verbose=False).run()
    want = _oracle().rdf_run(u, u.atoms, n_bins=64, range=(0.0, 5.0))
    assert np.array_equal(r.results.counts, want["counts"])
    np.testing.assert_allclose(r.results.rdf, want["rdf"], rtol=1e-6)


def test_on_disk_style_reader_gives_the_in_memory_results(tmp_path):
    """The analyses on a reader that seeks frames from a file (no whole-trajectory array;
    the staged feeder path) equal the runs on the in-memory trajectory: RDF with strided
    frames and explicit frame lists, residue groupings, per-frame cells."""
    from conftest import file_universe
    from mdhelper_b200.universe import SyntheticUniverse
    rng = np.random.default_rng(31)
    F, n = 9, 1200
    edges = rng.uniform(9.5, 10.5, (F, 3)).astype(np.float32)
    dims = np.concatenate([edges, np.full((F, 3), 90, np.float32)], 1)
    pos = (rng.random((F, n, 3)) * edges[:, None, :]).astype(np.float32)
    mem = SyntheticUniverse(pos, dims)
    disk = file_universe(tmp_path, pos, dims)
    S = _structure()
    kw = dict(n_bins=50, range=(0.0, 4.5), verbose=False, batch_frames=4)
    for run_kw in (dict(), dict(start=1, stop=8, step=3), dict(frames=[7, 2, 5])):
        a = S.RadialDistributionFunction(mem.atoms, **kw).run(**run_kw)
        b = S.RadialDistributionFunction(disk.atoms, **kw).run(**run_kw)
        assert np.array_equal(a.results.counts, b.results.counts)
        np.testing.assert_allclose(a.results.rdf, b.results.rdf, rtol=1e-12)
    g1m, g2m = mem.select(slice(0, 500)), mem.select(slice(500, n))
    g1d, g2d = disk.select(slice(0, 500)), disk.select(slice(500, n))
    a = S.RadialDistributionFunction(g1m, g2m, exclusion=(2, 2), **kw).run()
    b = S.RadialDistributionFunction(g1d, g2d, exclusion=(2, 2), **kw).run()
    assert np.array_equal(a.results.counts, b.results.counts)
    # the wavevector grid comes from the cell of the reader's CURRENT frame: same frame
    mem.trajectory[0]
    disk.trajectory[0]
    sa = S.StructureFactor([g1m, g2m], mode="partial", n_points=6, verbose=False).run()
    sb = S.StructureFactor([g1d, g2d], mode="partial", n_points=6, verbose=False,
                           batch_frames=2).run()
    np.testing.assert_allclose(sa.results.ssf, sb.results.ssf, rtol=1e-12)
    assert disk.trajectory.reads > 0


def test_coordinates_outside_the_cell_follow_the_reference_method_choice():
    """capped_distance's grid search moves the coordinates into the cell in float32 before
    taking differences, its brute force does not: the GPU
    path makes the same choice per frame ("auto"), and either can be forced."""
    rng = np.random.default_rng(123)
    dims = np.array([24.0, 25.5, 27.0, 90, 90, 90], np.float32)
    p = ((rng.random((4000, 3)) * 7 - 3) * dims[:3]).astype(np.float32)      # +-3 cells away
    q = ((rng.random((2500, 3)) * 5 - 2) * dims[:3]).astype(np.float32)
    p[:3] = [[-1e-9, 3.0, 4.0], [24.0, -25.5, 54.0], [-1e-6, 25.4999, 26.9999]]   # edge cases
    S, rp = _structure(), _oracle()
    small, large = (0.0, 3.0), (0.0, 11.5)      # r_max <= 0.3 L: grid search; else brute force
    for rng_, method in ((small, "nsgrid"), (large, "bruteforce")):
        for a, b, excl in ((p, p, None), (p, q, (2, 2))):
            want_auto = rp.radial_histogram(a, b, 60, rng_, dims, exclusion=excl)
            assert np.array_equal(want_auto, rp.radial_histogram(a, b, 60, rng_, dims,
                                                                 exclusion=excl, method=method))
            for mode in ("allpairs", "cells") if rng_ is small else ("allpairs",):
                for arith in ("auto", "off", "audit"):
                    st = {}
                    got = S.radial_histogram(a, b, 60, rng_, dims, exclusion=excl, mode=mode,
                                             arith=arith, stats=st)
                    assert np.array_equal(got, want_auto), (rng_, mode, arith)
                    assert st["audit_violations"] == 0
    # forcing either semantics
    brute = rp.radial_histogram(p, q, 60, small, dims, method="bruteforce")
    grid = rp.radial_histogram(p, q, 60, small, dims, method="nsgrid")
    assert np.array_equal(S.radial_histogram(p, q, 60, small, dims, wrap="never"), brute)
    assert np.array_equal(S.radial_histogram(p, q, 60, small, dims, wrap="always"), grid)
    assert np.array_equal(S.radial_histogram(p, q, 60, small, dims, wrap="never",
                                             mode="cells"), brute)
    assert abs(int(brute.sum()) - int(grid.sum())) < 1e-4 * brute.sum() + 10


def test_filter_bound_fuzz_in_audit_mode():
    """Seeded fuzz of the fp32 filter's error bound: random (non-cubic) boxes, ranges with
    and without a lower cut-off, bin counts, coordinate offsets of up to a few cells and group
    sizes -- all-pairs and cell-list kernels, same group and two groups.  In audit mode
    every pair is ALSO evaluated with the reference's fp64 arithmetic on the device: no
    pair the filter calls certain may disagree, and the counts equal the oracle's."""
    rng = np.random.default_rng(20260618)
    S, rp = _structure(), _oracle()
    n_checked = 0
    for trial in range(24):
        box = rng.uniform(6.0, 40.0, 3).astype(np.float32)
        dims = np.concatenate([box, [90, 90, 90]]).astype(np.float32)
        r_hi = float(np.float32(rng.uniform(0.08, 0.5) * box.min()))
        r_lo = float(np.float32(rng.choice([0.0, 0.0, rng.uniform(0.05, 0.5) * r_hi])))
        n_bins = int(rng.choice([1, 7, 64, 201, 500, 1500]))
        n1, n2 = int(rng.integers(40, 2500)), int(rng.integers(40, 2500))
        spread = rng.choice([1.0, 1.0, 3.0, 9.0])           # coordinates up to +-4 cells away
        shift = (rng.random(3) - 0.5) * (spread - 1.0)
        p1 = ((rng.random((n1, 3)) * spread - shift * 0 - (spread - 1) / 2) * box).astype(np.float32)
        same = bool(rng.integers(0, 2))
        p2 = p1 if same else ((rng.random((n2, 3)) * spread - (spread - 1) / 2) * box
                              ).astype(np.float32)
        excl = None if rng.random() < 0.6 else ((3, 3) if same else (2, 5))
        want = rp.radial_histogram(p1, p2, n_bins, (r_lo, r_hi), dims, exclusion=excl)
        modes = ["allpairs"] + (["cells"] if box.min() / (r_hi * 1.00001) >= 3.0 else [])
        for mode in modes:
            st = {}
            got = S.radial_histogram(p1, p2, n_bins, (r_lo, r_hi), dims, exclusion=excl,
                                     mode=mode, arith="audit", stats=st)
            info = (trial, mode, box.tolist(), r_lo, r_hi, n_bins, n1, n2, same, excl, st)
            assert st["audit_violations"] == 0, info
            assert np.array_equal(got, want), info
            n_checked += st["eligible"]
    assert n_checked >= 20          # the filter really ran for most configurations
