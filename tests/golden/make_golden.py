"""
Generates the golden fixtures in this directory FROM THE REAL REFERENCE CODE.

Run in the build container (needs ``/root/reference``; cannot run on the GPU box):

    python tests/golden/make_golden.py

What runs is the reference's own ``radial_histogram``,
``RadialDistributionFunction`` and ``StructureFactor`` (including its numba
kernels in ``algorithm/accelerated.py``), imported from ``/root/reference/src``
behind the stubs described in ``oracle/ref_harness.py``.  The one piece that is
NOT reference code is ``MDAnalysis.lib.distances.capped_distance`` (third-party,
absent here): the restated C oracle is injected in its place, so the RDF
fixtures pin "reference class code + restated capped_distance + real
numpy.histogram", while the S(q) fixtures pin the reference end to end.

Inputs are stored in the fixtures (small cases) or regenerated from a seed with
a sha256 of the coordinates stored beside the expected output (large cases).
"""

import hashlib
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from mdhelper_b200 import synthetic  # noqa: E402
from mdhelper_b200.universe import SyntheticUniverse  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def kat_radial_histogram(S):
    """The construction of tests/test_analysis_structure.py:21-40, seeded."""
    rng = np.random.default_rng(20260001)
    L = 20
    half_L = L // 2
    dims = np.array((L, L, L, 90, 90, 90), dtype=int)
    origin = half_L * np.ones(3)
    N = 1_000
    norm = L // 2 * rng.random(N)
    expected = np.histogram(norm, bins=half_L, range=(0, half_L + 1))[0]
    neighbors = rng.random((N, 3))
    neighbors *= norm[:, None] / np.linalg.norm(neighbors, axis=1, keepdims=True)
    neighbors += dims[:3] / 2
    got = S.radial_histogram(origin, neighbors, n_bins=half_L,
                             range=(0, half_L + 1), dims=dims)
    # the reference's own assertion (float32 rounding of the inputs can move a
    # point across a bin edge; record both)
    np.savez_compressed(OUT / "kat_radial_histogram.npz", origin=origin,
                        neighbors=neighbors, dims=dims, n_bins=half_L,
                        range=np.array((0, half_L + 1)),
                        expected_from_norms=expected, reference_counts=got)
    print("kat: reference == np.histogram(norms):", np.array_equal(expected, got))


def rdf_cases(S):
    cases = {}
    # (2) LJ fluid, N=1000, 5 frames, same group
    u = synthetic.lj_fluid(1000, 5, seed=20260001)
    pos = u.trajectory.coordinates.copy()
    dims = u.trajectory.unitcells[0].copy()
    r = S.RadialDistributionFunction(u.atoms, n_bins=201, range=(0.0, 5.375),
                                     verbose=False).run()
    cases["lj1000"] = dict(positions=pos, dims=dims, n_bins=201,
                           range=np.array((0.0, 5.375)),
                           counts=r.results.counts, rdf=r.results.rdf,
                           edges=r.results.edges, bins=r.results.bins)
    # parallel=True of the reference must give the same counts
    # (its per-frame worker and reduction are driven in-process: a fork pool
    # cannot pickle the stubbed third-party modules)
    rp = S.RadialDistributionFunction(u.atoms, n_bins=201, range=(0.0, 5.375),
                                      parallel=True, verbose=False)
    rp._setup_frames(rp._trajectory)
    rp._prepare()
    rp._results = [rp._single_frame_parallel(f, i)
                   for i, f in enumerate(range(rp.n_frames))]
    rp._conclude()
    assert np.array_equal(rp.results.counts, r.results.counts)
    assert np.allclose(rp.results.rdf, r.results.rdf, rtol=1e-12)

    # (3) two groups + exclusions, N=600 (ions: 300 + 300), 3 frames
    u, cat, an = synthetic.electrolyte(600, 3, seed=20260002)
    pos = u.trajectory.coordinates.copy()
    dims = u.trajectory.unitcells[0].copy()
    L = float(dims[0])
    r = S.RadialDistributionFunction(cat, an, n_bins=64, range=(0.0, L / 2),
                                     verbose=False).run()
    cases["twogroup"] = dict(positions=pos, dims=dims, n_cat=cat.n_atoms,
                             n_bins=64, range=np.array((0.0, L / 2)),
                             counts=r.results.counts, rdf=r.results.rdf)
    r = S.RadialDistributionFunction(u.atoms, n_bins=50, range=(0.0, 4.0),
                                     exclusion=(1, 1), verbose=False).run()
    cases["excl11"] = dict(positions=pos, dims=dims, n_bins=50,
                           range=np.array((0.0, 4.0)), exclusion=np.array((1, 1)),
                           counts=r.results.counts, rdf=r.results.rdf)
    r = S.RadialDistributionFunction(cat, an, n_bins=50, range=(0.5, 4.0),
                                     exclusion=(4, 10), verbose=False).run()
    cases["excl410"] = dict(positions=pos, dims=dims, n_cat=cat.n_atoms,
                            n_bins=50, range=np.array((0.5, 4.0)),
                            exclusion=np.array((4, 10)),
                            counts=r.results.counts, rdf=r.results.rdf)
    # (4) drop_axis=2 (two-dimensional RDF)
    r = S.RadialDistributionFunction(u.atoms, n_bins=40, range=(0.0, 3.5),
                                     drop_axis=2, verbose=False).run()
    cases["dropz"] = dict(positions=pos, dims=dims, n_bins=40,
                          range=np.array((0.0, 3.5)), drop_axis=2,
                          counts=r.results.counts, rdf=r.results.rdf)
    r = S.RadialDistributionFunction(cat, an, n_bins=40, range=(0.0, 3.5),
                                     drop_axis="x", norm="density",
                                     verbose=False).run()
    cases["dropx_density"] = dict(positions=pos, dims=dims, n_cat=cat.n_atoms,
                                  n_bins=40, range=np.array((0.0, 3.5)),
                                  drop_axis=0, counts=r.results.counts,
                                  rdf=r.results.rdf)

    # (5) adversarial: perfect lattice, separations exactly L/2 (round half) and
    # exactly on bin edges; non-cubic box; per-frame box change
    m, a = 8, np.float32(1.25)
    g = np.arange(m, dtype=np.float32) * a
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    pos = np.stack([lat, lat[::-1].copy(), np.roll(lat, 7, axis=0)]).astype(np.float32)
    dims = np.array([[10.0, 10.0, 10.0, 90, 90, 90],
                     [10.0, 10.0, 10.0, 90, 90, 90],
                     [10.0, 10.0, 10.0, 90, 90, 90]], np.float32)
    u = SyntheticUniverse(pos, dims)
    r = S.RadialDistributionFunction(u.atoms, n_bins=20, range=(0.0, 5.0),
                                     verbose=False).run()
    cases["lattice_edges"] = dict(positions=pos, dims=dims, n_bins=20,
                                  range=np.array((0.0, 5.0)),
                                  counts=r.results.counts, rdf=r.results.rdf)
    rng = np.random.default_rng(20260005)
    dims = np.array([[9.0, 11.5, 14.25, 90, 90, 90],
                     [9.5, 11.0, 14.0, 90, 90, 90]], np.float32)
    pos = (rng.random((2, 700, 3)) * dims[:, None, :3]).astype(np.float32)
    u = SyntheticUniverse(pos, dims)
    r = S.RadialDistributionFunction(u.atoms, n_bins=33, range=(0.0, 4.4),
                                     verbose=False).run()
    cases["noncubic_npt"] = dict(positions=pos, dims=dims, n_bins=33,
                                 range=np.array((0.0, 4.4)),
                                 counts=r.results.counts, rdf=r.results.rdf)
    # unwrapped coordinates (outside [0, L)): brute-force path of the reference
    pos = ((rng.random((1, 500, 3)) * 3 - 1) * dims[0, None, :3]).astype(np.float32)
    u = SyntheticUniverse(pos, dims[:1])
    r = S.RadialDistributionFunction(u.atoms, n_bins=33, range=(0.0, 4.4),
                                     verbose=False).run()
    cases["unwrapped"] = dict(positions=pos, dims=dims[:1], n_bins=33,
                              range=np.array((0.0, 4.4)),
                              counts=r.results.counts, rdf=r.results.rdf)
    for name, d in cases.items():
        np.savez_compressed(OUT / f"rdf_{name}.npz", **d)
        print("rdf", name, int(d["counts"].sum()))


def sq_cases(S):
    # (6a) N=1000, 2 frames, default lattice grid, all modes and both forms
    u, cat, an = synthetic.electrolyte(1000, 2, seed=20260006)
    pos = u.trajectory.coordinates.copy()
    dims = u.trajectory.unitcells[0].copy()
    L = float(dims[0])
    out = dict(positions=pos, dims=dims, n_cat=cat.n_atoms, n_points=12,
               q_max=2 * np.pi * 8 / L)
    for mode, groups in ((None, [cat, an]), ("pair", [cat, an]),
                         ("partial", [cat, an])):
        for form in ("exp", "trig"):
            s = S.StructureFactor(groups, mode=mode, form=form, n_points=12,
                                  q_max=out["q_max"], verbose=False).run()
            key = f"{mode}_{form}"
            out[f"ssf_{key}"] = s.results.ssf
            out[f"wavenumbers_{key}"] = s.results.wavenumbers
    s = S.StructureFactor([u.atoms], n_points=12, q_max=out["q_max"],
                          sort=False, unique=False, parallel=True,
                          verbose=False).run()
    out["ssf_raw"] = s.results.ssf
    out["wavevectors_raw"] = s._wavevectors
    # raw rho(q) of the last frame straight from the numba kernel (seam #2)
    acc = ref_harness.accelerated()
    out["rho_last"] = acc.delta_fourier_transform_sum_2d_2d(
        s._wavevectors, pos[-1].astype(np.float64))
    # off-lattice: n_surfaces, and user wavevectors
    s = S.StructureFactor([u.atoms], n_points=6, n_surfaces=3,
                          n_surface_points=8, verbose=False).run()
    out["ssf_surfaces"] = s.results.ssf
    out["wavenumbers_surfaces"] = s.results.wavenumbers
    rng = np.random.default_rng(20260007)
    wv = rng.normal(size=(37, 3)) * 2.0
    s = S.StructureFactor([cat, an], mode="partial", wavevectors=wv, sort=False,
                          unique=False, verbose=False).run()
    out["wavevectors_user"] = wv
    out["ssf_user"] = s.results.ssf
    np.savez_compressed(OUT / "sq_small.npz", **out)
    print("sq small", out["ssf_None_exp"].shape, out["ssf_partial_exp"].shape)

    # (6b) non-cubic box
    rng = np.random.default_rng(20260008)
    dims = np.array([9.0, 11.5, 14.25, 90, 90, 90], np.float32)
    pos = (rng.random((2, 300, 3)) * dims[:3]).astype(np.float32)
    u = SyntheticUniverse(pos, dims)
    s = S.StructureFactor([u.atoms], n_points=10, q_max=4.0, verbose=False).run()
    np.savez_compressed(OUT / "sq_noncubic.npz", positions=pos, dims=dims,
                        n_points=10, q_max=4.0, ssf=s.results.ssf,
                        wavenumbers=s.results.wavenumbers)

    # (6c) config-4 sized frame: N=50,000, n_max=16 (N_q=2446), one frame;
    # coordinates regenerated from the seed (checksum stored)
    u = synthetic.lj_fluid(50_000, 1, seed=20260004)
    L = float(u.trajectory.unitcells[0, 0])
    q_max = 2 * np.pi * 16 / L
    s = S.StructureFactor([u.atoms], n_points=32, q_max=q_max, parallel=True,
                          sort=False, unique=False, verbose=False).run()
    su = S.StructureFactor([u.atoms], n_points=32, q_max=q_max, parallel=True,
                           verbose=False).run()
    np.savez_compressed(OUT / "sq_cfg4_frame.npz", n=50_000, seed=20260004,
                        positions_sha256=sha(u.trajectory.coordinates),
                        L=np.float32(L), q_max=q_max, ssf_raw=s.results.ssf,
                        ssf_unique=su.results.ssf,
                        wavenumbers_unique=su.results.wavenumbers)
    print("sq cfg4", s.results.ssf.shape, su.results.ssf.shape)


def rdf_post(S):
    """The reference's g(r) post-processing functions (structure.py:106-442) on the
    g(r) of the lj1000 fixture.  (Its Hankel transform only accepts a one-element
    wavenumber array: ``jv(0, q * r)`` does not broadcast otherwise.)"""
    g = dict(np.load(OUT / "rdf_lj1000.npz"))
    bins, rdf = g["bins"], g["rdf"]
    qa, sa = S.calculate_structure_factor(bins, rdf, True, 0.8, n_q=64)
    qb, sb = S.calculate_structure_factor(bins, rdf, False, 0.8, 0.3, 0.7, n_q=32,
                                          formalism="AL")
    np.savez_compressed(
        OUT / "rdf_post.npz", bins=bins, rdf=rdf, rho=0.8,
        coordination_numbers=S.calculate_coordination_numbers(bins, rdf, 0.8,
                                                              n_coord_nums=3),
        q_fz=qa, ssf_fz=sa, q_al=qb, ssf_al=sb,
        hankel_07=S.zeroth_order_hankel_transform(bins, rdf - 1, np.array([0.7])),
        rft=S.radial_fourier_transform(bins, rdf - 1, np.array([0.0, 0.5, 1.0])))
    print("rdf post-processing saved")


def isf_cases(S):
    """IntermediateScatteringFunction (structure.py:1552-2127), the real class: the
    coherent and incoherent parts, all modes, a strided frame selection, both forms,
    user wavevectors (off the lattice)."""
    u, cat, an = synthetic.electrolyte(600, 14, seed=20260009)
    pos = u.trajectory.coordinates.copy()
    dims = u.trajectory.unitcells[0].copy()
    L = float(dims[0])
    out = dict(positions=pos, dims=dims, n_cat=cat.n_atoms, n_points=8,
               q_max=2 * np.pi * 5 / L, n_lags=5, dt=0.25)
    common = dict(n_points=8, q_max=out["q_max"], n_lags=5, incoherent=True, dt=0.25,
                  verbose=False)
    for mode, groups in ((None, [cat, an]), ("pair", [cat, an]), ("partial", [cat, an])):
        for form in ("exp", "trig"):
            r = S.IntermediateScatteringFunction(groups, mode=mode, form=form,
                                                 **common).run()
            key = f"{mode}_{form}"
            out[f"cisf_{key}"] = r.results.cisf
            out[f"iisf_{key}"] = r.results.iisf
            out[f"wavenumbers_{key}"] = r.results.wavenumbers
            out[f"times_{key}"] = r.results.times
    # strided frames, default n_lags (= number of frames), raw ordering
    r = S.IntermediateScatteringFunction([u.atoms], n_points=8, q_max=out["q_max"],
                                         incoherent=True, sort=False, unique=False,
                                         dt=0.25, verbose=False).run(start=2, stop=14,
                                                                     step=3)
    out["cisf_strided"] = r.results.cisf
    out["iisf_strided"] = r.results.iisf
    out["times_strided"] = r.results.times
    rng = np.random.default_rng(20260010)
    wv = rng.normal(size=(23, 3)) * 1.5
    r = S.IntermediateScatteringFunction([cat, an], mode="partial", wavevectors=wv,
                                         n_lags=4, incoherent=True, sort=False,
                                         unique=False, dt=0.25, verbose=False).run()
    out["wavevectors_user"] = wv
    out["cisf_user"] = r.results.cisf
    out["iisf_user"] = r.results.iisf
    np.savez_compressed(OUT / "isf_small.npz", **out)
    print("isf small", out["cisf_None_exp"].shape, out["cisf_partial_exp"].shape,
          "exp vs trig", np.abs(out["cisf_partial_exp"] - out["cisf_partial_trig"]).max())


def scsf_cases(S):
    """SingleChainStructureFactor (analysis/polymer.py:805-1129), the real class, on a
    small bead-spring melt: wrapped coordinates as stored, and unwrap=True."""
    P = ref_harness.polymer()
    u = synthetic.polymer_melt(12, 20, 6, seed=20260012)
    out = dict(positions=u.trajectory.coordinates.copy(),
               dims=u.trajectory.unitcells[0].copy(), n_chains=12, n_monomers=20,
               n_points=6)
    for unwrap in (False, True):
        r = P.SingleChainStructureFactor(u.atoms, n_points=6, n_chains=12, n_monomers=20,
                                         unwrap=unwrap, verbose=False).run()
        out[f"scsf_unwrap{int(unwrap)}"] = r.results.scsf
        out["wavenumbers"] = r.results.wavenumbers
    r = P.SingleChainStructureFactor(u.atoms, n_points=6, n_chains=12, n_monomers=20,
                                     verbose=False).run(start=1, stop=6, step=2)
    out["scsf_strided"] = r.results.scsf
    np.savez_compressed(OUT / "scsf_small.npz", **out)
    print("scsf small", out["scsf_unwrap0"].shape,
          "wrapped vs unwrapped", np.abs(out["scsf_unwrap0"] - out["scsf_unwrap1"]).max())


def com_cases(S):
    """Centres of mass (groupings="residues"/"segments") from the reference's REAL
    ``center_of_mass`` (algorithm/molecule.py:15-310, imported by structure.py:25) and the
    real analysis classes on top of it.  Equal-sized residues go through its einsum
    (:303-304) -- pure reference code; unequal sizes go through
    ``g.atoms.center_of_mass()`` (:240-241), third-party MDAnalysis, restated in
    mdhelper_b200.universe.AtomGroup.center_of_mass."""
    import importlib
    mol = importlib.import_module("mdhelper.algorithm.molecule")
    assert S.center_of_mass is mol.center_of_mass
    rng = np.random.default_rng(20260013)
    out = {}
    for tag, sizes in (("equal", np.full(150, 4)), ("unequal", rng.integers(1, 7, 160))):
        n = int(sizes.sum())
        L = np.float32(12.5)
        pos = (rng.random((4, n, 3)) * float(L)).astype(np.float32)
        res = np.repeat(np.arange(len(sizes)), sizes)
        seg = res // 5
        masses = rng.choice([1.008, 12.011, 14.007, 15.999, 32.06], n)
        u = SyntheticUniverse(pos, np.array([L, L, L, 90, 90, 90], np.float32),
                              resindices=res, segindices=seg, masses=masses)
        out[f"{tag}_positions"] = pos
        out[f"{tag}_dims"] = u.trajectory.unitcells[0].copy()
        out[f"{tag}_resindices"], out[f"{tag}_segindices"] = res, seg
        out[f"{tag}_masses"] = masses
        n_a = int(np.searchsorted(res, len(sizes) // 3))       # a residue boundary
        a, b = u.select(slice(0, n_a)), u.select(slice(n_a, n))
        out[f"{tag}_n_a"] = n_a
        # the centres themselves, frame by frame
        for grouping in ("residues", "segments"):
            com = []
            for f in range(4):
                u.trajectory[f]
                com.append(np.asarray(mol.center_of_mass(u.atoms, grouping), np.float64))
            out[f"{tag}_com_{grouping}"] = np.stack(com)
        kw = dict(n_bins=40, range=(0.0, 6.0), verbose=False)
        r = S.RadialDistributionFunction(a, b, groupings="residues", **kw).run()
        out[f"{tag}_rdf_res_counts"], out[f"{tag}_rdf_res"] = r.results.counts, r.results.rdf
        r = S.RadialDistributionFunction(u.atoms, groupings="segments", **kw).run()
        out[f"{tag}_rdf_seg_counts"], out[f"{tag}_rdf_seg"] = r.results.counts, r.results.rdf
        r = S.RadialDistributionFunction(a, b, groupings=("residues", "atoms"), **kw).run()
        out[f"{tag}_rdf_mix_counts"], out[f"{tag}_rdf_mix"] = r.results.counts, r.results.rdf
        s = S.StructureFactor([a, b], groupings="residues", mode="partial", n_points=5,
                              verbose=False).run()
        out[f"{tag}_ssf_res"], out[f"{tag}_ssf_wavenumbers"] = s.results.ssf, s.results.wavenumbers
        print("com", tag, out[f"{tag}_com_residues"].shape, out[f"{tag}_rdf_res_counts"].sum())
    np.savez_compressed(OUT / "com_ref.npz", **out)


def f64_cases(S):
    """Positions the reference keeps in float64 on the Fourier path, from its REAL classes:
    IntermediateScatteringFunction with groupings="residues" (centres of mass and their
    displacements, structure.py:1927-2033) and SingleChainStructureFactor with
    grouping="residues" (equal-sized monomers through center_of_mass' array form,
    polymer.py:1080-1093), wrapped and unwrapped over several box lengths."""
    P = ref_harness.polymer()
    rng = np.random.default_rng(20260651)
    out = {}
    # ISF: 120 residues of 3 atoms, two groups, drifting coordinates
    n_res, per, F, L = 120, 3, 8, np.float32(13.0)
    n = n_res * per
    pos = (rng.random((1, n, 3)) * float(L)
           + rng.normal(0, 0.25, (F, n, 3)).cumsum(axis=0)).astype(np.float32)
    res = np.arange(n) // per
    masses = rng.choice([1.008, 12.011, 15.999], n)
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    u = SyntheticUniverse(pos, dims, resindices=res, segindices=res // 10, masses=masses)
    n_a = (n_res // 3) * per
    a, b = u.select(slice(0, n_a)), u.select(slice(n_a, n))
    r = S.IntermediateScatteringFunction([a, b], groupings="residues", mode="partial",
                                         n_points=5, n_lags=4, incoherent=True, dt=1.0,
                                         verbose=False).run()
    out.update(isf_positions=pos, isf_dims=dims, isf_resindices=res, isf_masses=masses,
               isf_n_a=n_a, isf_cisf=r.results.cisf, isf_iisf=r.results.iisf,
               isf_wavenumbers=r.results.wavenumbers)
    # SCSF: 10 chains of 12 monomers of 3 atoms, drifting several boxes in 7 frames
    n_chains, n_mono, F, L = 10, 12, 7, np.float32(9.0)
    start = rng.random((n_chains, 1, 3)) * float(L)
    chain = start + rng.normal(0, 0.5, (n_chains, n_mono, 3)).cumsum(axis=1)
    drift = np.arange(F)[:, None, None, None] * np.array([3.7, -2.9, 2.3])
    true = chain[None] + drift + rng.normal(0, 0.05, (F, n_chains, n_mono, 3))
    true = np.repeat(true.reshape(F, -1, 3), per, axis=1)
    true = true + rng.normal(0, 0.1, true.shape)
    wrapped = np.mod(true, float(L)).astype(np.float32)
    wrapped[wrapped >= L] = 0.0
    n = wrapped.shape[1]
    masses = rng.uniform(1, 20, n)
    dims = np.array([L, L, L, 90, 90, 90], np.float32)
    u = SyntheticUniverse(wrapped, dims, resindices=np.arange(n) // per, masses=masses)
    out.update(scsf_positions=wrapped, scsf_dims=dims, scsf_masses=masses, scsf_per=per,
               scsf_n_chains=n_chains, scsf_n_monomers=n_mono)
    for unwrap in (False, True):
        r = P.SingleChainStructureFactor(u.atoms, grouping="residues", n_points=5,
                                         n_chains=n_chains, n_monomers=n_mono, unwrap=unwrap,
                                         verbose=False).run()
        out[f"scsf_res_unwrap{int(unwrap)}"] = r.results.scsf
        out["scsf_wavenumbers"] = r.results.wavenumbers
    np.savez_compressed(OUT / "f64_ref.npz", **out)
    print("f64", out["isf_cisf"].shape, out["scsf_res_unwrap1"].shape,
          np.abs(out["scsf_res_unwrap0"] - out["scsf_res_unwrap1"]).max())


if __name__ == "__main__":
    S = ref_harness.load()
    which = sys.argv[1:] or ["kat", "rdf", "post", "sq", "isf", "scsf", "com", "f64"]
    if "com" in which:
        com_cases(S)
    if "f64" in which:
        f64_cases(S)
    if "kat" in which:
        kat_radial_histogram(S)
    if "rdf" in which:
        rdf_cases(S)
    if "post" in which:
        rdf_post(S)
    if "sq" in which:
        sq_cases(S)
    if "isf" in which:
        isf_cases(S)
    if "scsf" in which:
        scsf_cases(S)
