"""
CPU restatement of the reference hot path -- TEST INFRASTRUCTURE ONLY
(see ``oracle/__init__.py``).  numpy + the C functions of ``mdh_oracle.c``.

Every function cites the reference lines it follows (paths relative to
``/root/reference/``).  Binning is the REAL ``numpy.histogram``; distances are
the restated third-party ``MDAnalysis.lib.distances.capped_distance``
("parity unpinned" against MDAnalysis itself -- see ``mdh_oracle.c``).
"""

from itertools import combinations_with_replacement

import numpy as np

from . import lib

_CHUNK_PAIRS = 40_000_000


def _as_coords(x):
    x = np.asarray(x)
    if x.ndim == 1:            # a single (3,) coordinate is one particle
        x = x[None, :]
    return np.ascontiguousarray(x, dtype=np.float32)


def triclinic_vectors(dimensions):
    """Restated third-party ``MDAnalysis.lib.mdamath.triclinic_vectors`` [recall]: the
    lower-triangular float32 cell matrix of ``(lx, ly, lz, alpha, beta, gamma)``; float64
    trigonometry, exact values for right angles, zeros for an invalid cell."""
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim
    if not (np.all(dim > 0.0) and alpha < 180.0 and beta < 180.0 and gamma < 180.0):
        return np.zeros((3, 3), dtype=np.float32)
    if alpha == beta == gamma == 90.0:
        return np.diag(dim[:3]).astype(np.float32)
    m = np.zeros((3, 3), dtype=np.float64)
    m[0, 0] = lx
    cos_alpha = 0.0 if alpha == 90.0 else np.cos(np.deg2rad(alpha))
    cos_beta = 0.0 if beta == 90.0 else np.cos(np.deg2rad(beta))
    cos_gamma = 0.0 if gamma == 90.0 else np.cos(np.deg2rad(gamma))
    sin_gamma = 1.0 if gamma == 90.0 else np.sin(np.deg2rad(gamma))
    m[1, 0], m[1, 1] = ly * cos_gamma, ly * sin_gamma
    m[2, 0] = lz * cos_beta
    m[2, 1] = lz * (cos_alpha - cos_beta * cos_gamma) / sin_gamma
    m[2, 2] = np.sqrt(lz * lz - m[2, 0] ** 2 - m[2, 1] ** 2)
    if not m[2, 2] > 0.0:
        return np.zeros((3, 3), dtype=np.float32)
    return m.astype(np.float32)


def _capped_distance_triclinic(ref, conf, max_cutoff, lo, box, return_distances):
    """Triclinic cells: brute force over all pairs (mdh_oracle.c, wrap + 27 images)."""
    h = np.ascontiguousarray(triclinic_vectors(box), dtype=np.float32)
    if not np.all(np.diag(h) > 0):
        raise ValueError("oracle: invalid unit cell")
    L = lib()
    n1, n2 = len(ref), len(conf)
    cap = max(1024, int(1.5 * n1 * n2 * 4.19 * max_cutoff ** 3
                        / float(np.prod(np.diag(h), dtype=np.float64))))
    cap = min(cap, n1 * n2)
    while True:
        pairs = np.empty((cap, 2), dtype=np.int64)
        dist = np.empty(cap, dtype=np.float64)
        m = L.mdho_capped_distance_triclinic(
            ref.ctypes.data, n1, conf.ctypes.data, n2, h.ctypes.data, float(max_cutoff), lo,
            pairs.ctypes.data, dist.ctypes.data, cap)
        if m < 0:
            raise MemoryError("oracle allocation failed")
        if m <= cap:
            return (pairs[:m], dist[:m]) if return_distances else pairs[:m]
        cap = int(m)


def capped_distance(reference, configuration, max_cutoff, min_cutoff=None,
                    box=None, method=None, return_distances=True):
    """
    Restated ``MDAnalysis.lib.distances.capped_distance`` (third-party; call
    site ``src/mdhelper/analysis/structure.py:93-96``; arithmetic: SURVEY.md
    Appendix A).  Triclinic boxes: all pairs, wrap + shortest of 27 images.

    Returns ``pairs`` (int64 ``[M, 2]``) and ``distances`` (float64 ``[M]``).
    """
    ref = _as_coords(reference)
    conf = _as_coords(configuration)
    if box is None:
        raise NotImplementedError("oracle: a periodic box is required")
    box = np.ascontiguousarray(box, dtype=np.float32)
    if box.shape != (6,):
        raise NotImplementedError("oracle: box must be (lx, ly, lz, alpha, beta, gamma)")
    lo = -np.inf if min_cutoff is None else float(min_cutoff)
    if not np.all(box[3:] == 90):
        return _capped_distance_triclinic(ref, conf, float(max_cutoff), lo, box,
                                          return_distances)
    n1, n2 = len(ref), len(conf)
    if method is None:                       # Appendix A item 2
        if n1 < 10 or n2 < 10:
            method = "bruteforce"
        elif n1 * n2 >= 1e8:
            method = "nsgrid"
        elif max_cutoff > 0.3 * box[:3].min():
            method = "bruteforce"
        else:
            method = "nsgrid"
    L = lib()
    if method == "nsgrid":
        # the grid search works on copies moved into the primary cell in float32
        # (FastNS -> apply_PBC -> _ortho_pbc); the pair arithmetic is the same afterwards
        ref, conf = ref.copy(), conf.copy()
        L.mdho_ortho_pbc(ref.ctypes.data, n1, box.ctypes.data)
        L.mdho_ortho_pbc(conf.ctypes.data, n2, box.ctypes.data)
        cap = max(1024, int(1.3 * n1 * n2 * 4.19 * max_cutoff ** 3
                            / float(np.prod(box[:3], dtype=np.float64))))
        while True:
            pairs = np.empty((cap, 2), dtype=np.int64)
            dist = np.empty(cap, dtype=np.float64)
            m = L.mdho_capped_distance_cells(
                ref.ctypes.data, n1, conf.ctypes.data, n2, box.ctypes.data,
                float(max_cutoff), lo, pairs.ctypes.data, dist.ctypes.data, cap)
            if m == -1:                      # fewer than 3 cells per axis: the restated
                method = "bruteforce"        # 27-cell search does not apply; same pairs and
                break                        # (wrapped) coordinates through all pairs
            if m < 0:
                raise MemoryError("oracle cell list allocation failed")
            if m <= cap:
                pairs, dist = pairs[:m], dist[:m]
                return (pairs, dist) if return_distances else pairs
            cap = int(m)
    # brute force, in row chunks to bound memory
    rows = max(1, _CHUNK_PAIRS // max(n2, 1))
    out_p, out_d = [], []
    for i0 in range(0, n1, rows):
        i1 = min(n1, i0 + rows)
        cap = (i1 - i0) * n2
        pairs = np.empty((cap, 2), dtype=np.int64)
        dist = np.empty(cap, dtype=np.float64)
        m = L.mdho_capped_distance_bruteforce(
            ref.ctypes.data, i0, i1, conf.ctypes.data, n2, box.ctypes.data,
            float(max_cutoff), lo, pairs.ctypes.data, dist.ctypes.data, cap)
        out_p.append(pairs[:m].copy() if m < cap else pairs)
        out_d.append(dist[:m].copy() if m < cap else dist)
    pairs = np.concatenate(out_p) if out_p else np.empty((0, 2), np.int64)
    dist = np.concatenate(out_d) if out_d else np.empty(0, np.float64)
    return (pairs, dist) if return_distances else pairs


def radial_histogram(pos1, pos2, n_bins, range, dims, *, exclusion=None,
                     method=None):
    """Restates ``radial_histogram``, ``analysis/structure.py:32-104``."""
    pairs, dist = capped_distance(
        pos1, pos2, range[1], range[0] - np.finfo(np.float64).eps, box=dims,
        method=method)
    if exclusion is not None:
        dist = dist[pairs[:, 0] // exclusion[0] != pairs[:, 1] // exclusion[1]]
    return np.histogram(dist, bins=n_bins, range=range)[0]


def rdf_run(universe, ag1, ag2=None, n_bins=201, range=(0.0, 15.0), *,
            drop_axis=None, norm="rdf", exclusion=None, frames=None,
            method=None):
    """
    Restates the serial RDF frame loop: ``_prepare`` (``structure.py:734-748``),
    ``_single_frame`` (``:750-791``, ``groupings="atoms"``, no ``n_batches``) and
    ``_conclude`` (``:837-862``).  Returns a dict with the ``results.*`` arrays
    plus ``volume`` (the accumulated area or volume).
    """
    ag2 = ag1 if ag2 is None else ag2
    traj = universe.trajectory
    frames = np.arange(len(traj)) if frames is None else np.asarray(frames)
    edges = np.linspace(*range, n_bins + 1)
    counts = np.zeros(n_bins, dtype=int)
    vol = 0.0
    for f in frames:
        ts = traj[int(f)]
        dims = ts.dimensions.copy()
        pos1, pos2 = ag1.positions, ag2.positions
        if drop_axis is None:
            vol += ts.volume
        else:
            pos1[:, drop_axis] = pos2[:, drop_axis] = 0
            dims[drop_axis] = dims[:3].max()
            vol += np.delete(dims[:3], drop_axis).prod()
        counts += radial_histogram(pos1, pos2, n_bins, range, dims,
                                   exclusion=exclusion, method=method)
    n_frames = len(frames)
    nrm = n_frames
    if norm is not None:
        if drop_axis is None:
            nrm = nrm * 4 * np.pi * np.diff(edges ** 3) / 3
        else:
            nrm = nrm * np.pi * np.diff(edges ** 2)
        if norm == "rdf":
            n2 = ag2.n_atoms - (exclusion[1] if exclusion else 0)
            nrm = nrm * (ag1.n_atoms * n2 * n_frames / vol)
    return {"edges": edges, "bins": (edges[:-1] + edges[1:]) / 2,
            "counts": counts, "rdf": counts / nrm, "volume": vol}


def delta_fourier_transform_sum(qs, rs, n_threads=1):
    """
    Restates ``delta_fourier_transform_sum_2d_2d`` (``algorithm/accelerated.py:81-122``)
    and, for ``n_threads > 1``, its prange twin (``:124-165``).
    """
    qs = np.ascontiguousarray(qs, dtype=np.float64)
    rs = np.ascontiguousarray(rs, dtype=np.float64)
    out = np.empty(len(qs), dtype=np.complex128)
    lib().mdho_delta_fourier_transform_sum(
        qs.ctypes.data, len(qs), rs.ctypes.data, len(rs), out.ctypes.data,
        int(n_threads))
    return out


def lattice_wavevectors(dimensions, n_points=32, q_max=None):
    """
    Restates the default wavevector grid of ``StructureFactor.__init__``
    (``analysis/structure.py:1376-1416`` without ``n_surfaces``): first-octant
    reciprocal lattice, ``np.meshgrid`` 'xy' ordering, optional ``q_max`` filter.
    ``dimensions`` is the float32 box edge triple of the universe.
    """
    dimensions = np.asarray(dimensions)
    if np.allclose(dimensions, dimensions[0]):
        g = 2 * np.pi * np.arange(n_points) / dimensions[0]
        grids = (g, g, g)
    else:
        grids = [2 * np.pi * np.arange(n_points) / L for L in dimensions]
    wv = np.stack(np.meshgrid(*grids), -1).reshape(-1, 3)
    wn = np.linalg.norm(wv, axis=1)
    if q_max is not None:
        keep = wn <= q_max
        wv, wn = wv[keep], wn[keep]
    return wv, wn


def ssf_run(universe, groups, *, mode=None, wavevectors=None, n_points=32,
            q_max=None, sort=True, unique=True, frames=None, n_threads=1):
    """
    Restates the ``form="exp"`` structure-factor loop: ``_prepare``
    (``structure.py:1456-1479``), ``_single_frame`` (``:1481-1508``) and
    ``_conclude`` (``:1529-1550``), ``groupings="atoms"``.
    """
    traj = universe.trajectory
    frames = np.arange(len(traj)) if frames is None else np.asarray(frames)
    if wavevectors is None:
        wavevectors, wavenumbers = lattice_wavevectors(
            universe.dimensions[:3].copy(), n_points, q_max)
    else:
        wavevectors = np.asarray(wavevectors, dtype=np.float64)
        wavenumbers = np.linalg.norm(wavevectors, axis=1)
        if q_max is not None:
            keep = wavenumbers <= q_max
            wavevectors, wavenumbers = wavevectors[keep], wavenumbers[keep]
    n_groups = len(groups)
    pairs = (tuple(combinations_with_replacement(np.arange(n_groups).tolist(), 2))
             if mode == "partial"
             else ((0, n_groups - 1),) if mode == "pair" else ((None, None),))
    Ns = [g.n_atoms for g in groups]
    N = sum(Ns)
    slices, idx = [], 0
    for n in Ns:
        slices.append(slice(idx, idx + n))
        idx += n
    positions = np.empty((N, 3))
    ssf = np.zeros((len(pairs), len(wavenumbers)))
    for f in frames:
        traj[int(f)]
        for g, s in zip(groups, slices):
            positions[s] = g.positions
        if mode is None:
            rho = delta_fourier_transform_sum(wavevectors, positions, n_threads)
            ssf += (rho * rho.conj()).real
        else:
            for i, (j, k) in enumerate(pairs):
                rj = delta_fourier_transform_sum(
                    wavevectors, positions[slices[j]], n_threads)
                if j == k:
                    ssf[i] += (rj * rj.conj()).real
                else:
                    rk = delta_fourier_transform_sum(
                        wavevectors, positions[slices[k]], n_threads)
                    ssf[i] += 2 * (rj * rk.conj()).real
    ssf /= len(frames) * N
    wn_out = np.unique(wavenumbers.round(11)) if unique else wavenumbers
    if unique:
        ssf = np.hstack([ssf[:, np.isclose(q, wavenumbers)]
                         .mean(axis=1, keepdims=True) for q in wn_out])
    if sort:
        order = np.argsort(wn_out)
        wn_out, ssf = wn_out[order], ssf[:, order]
    return {"pairs": pairs, "wavenumbers": wn_out, "ssf": ssf,
            "wavevectors": wavevectors}


def isf_run(universe, groups, *, mode=None, wavevectors=None, n_points=32,
            q_max=None, sort=True, unique=True, n_lags=None, incoherent=False,
            dt=None, start=None, stop=None, step=None, n_threads=1):
    """
    Restates ``IntermediateScatteringFunction`` with ``form="exp"``,
    ``groupings="atoms"``: ``_prepare`` (``structure.py:1899-1957``), the sliding
    window of ``_single_frame`` (``:1959-2033``) and ``_conclude`` (``:2087-2127``).
    """
    traj = universe.trajectory
    sl = slice(start, stop, step).indices(len(traj))
    frames = np.arange(*sl)
    n_frames = len(frames)
    df = sl[2]
    dt = traj.dt if dt is None else dt
    if wavevectors is None:
        wavevectors, wavenumbers = lattice_wavevectors(
            universe.dimensions[:3].copy(), n_points, q_max)
    else:
        wavevectors = np.asarray(wavevectors, dtype=np.float64)
        wavenumbers = np.linalg.norm(wavevectors, axis=1)
        if q_max is not None:
            keep = wavenumbers <= q_max
            wavevectors, wavenumbers = wavevectors[keep], wavenumbers[keep]
    n_groups = len(groups)
    n_lags = n_lags or n_frames
    pairs = (tuple(combinations_with_replacement(np.arange(n_groups).tolist(), 2))
             if mode == "partial"
             else ((0, n_groups - 1),) if mode == "pair" else ((None, None),))
    Ns = [g.n_atoms for g in groups]
    N = sum(Ns)
    slices, idx = [], 0
    for n in Ns:
        slices.append(slice(idx, idx + n))
        idx += n
    n_q = len(wavenumbers)
    positions = np.zeros((n_lags, N, 3))
    exp_sum = np.empty((n_lags, 1 if mode is None else n_groups, n_q), dtype=complex)
    cisf = np.zeros((n_lags, 1 if mode is None else len(pairs), n_q))
    iisf = np.zeros((n_lags, 1 if mode is None else n_groups, n_q)) if incoherent else None
    dfts = lambda r: delta_fourier_transform_sum(wavevectors, r, n_threads)  # noqa: E731
    for fi, f in enumerate(frames):
        traj[int(f)]
        rcfi = fi % n_lags
        for g, s in zip(groups, slices):
            positions[rcfi, s] = g.positions
        if mode is None:
            exp_sum[rcfi] = dfts(positions[rcfi])
            for lag in range(min(n_lags, fi + 1)):
                rifi = (fi - lag) % n_lags
                cisf[lag] += (exp_sum[rifi] * exp_sum[rcfi].conj()).real
                if incoherent:
                    iisf[lag] += dfts(positions[rcfi] - positions[rifi]).real
        else:
            for i in range(n_groups):
                exp_sum[rcfi, i] = dfts(positions[rcfi, slices[i]])
            for lag in range(min(n_lags, fi + 1)):
                rifi = (fi - lag) % n_lags
                for i, (j, k) in enumerate(pairs):
                    if j == k:
                        cisf[lag, i] += (exp_sum[rifi, j] * exp_sum[rcfi, j].conj()).real
                        if incoherent:
                            iisf[lag, j] += dfts(positions[rcfi, slices[j]]
                                                 - positions[rifi, slices[j]]).real
                    else:
                        cisf[lag, i] += ((exp_sum[rifi, j] * exp_sum[rcfi, k].conj()).real
                                         + (exp_sum[rifi, k] * exp_sum[rcfi, j].conj()).real)
    normalization = N * np.arange(n_frames, n_frames - n_lags, -1)[:, None, None]
    cisf /= normalization
    if incoherent:
        iisf /= normalization
    wn_out = np.unique(wavenumbers.round(11)) if unique else wavenumbers
    if unique:
        cisf = np.stack([cisf[:, :, np.isclose(q, wavenumbers)].mean(axis=2)
                         for q in wn_out], axis=-1)
        if incoherent:
            iisf = np.stack([iisf[:, :, np.isclose(q, wavenumbers)].mean(axis=2)
                             for q in wn_out], axis=-1)
    if sort:
        order = np.argsort(wn_out)
        wn_out, cisf = wn_out[order], cisf[:, :, order]
        if incoherent:
            iisf = iisf[:, :, order]
    return {"pairs": pairs, "wavenumbers": wn_out, "cisf": cisf, "iisf": iisf,
            "times": df * dt * np.arange(n_lags), "wavevectors": wavevectors}


def scsf_run(universe, group, *, n_points=32, n_chains, n_monomers, unwrap=False,
             start=None, stop=None, step=None):
    """
    Restates ``SingleChainStructureFactor`` (``analysis/polymer.py:805-1129``) for
    ``grouping="atoms"`` with explicit ``n_chains`` / ``n_monomers``: wavevector grid
    (``:977-984``), per-frame unwrapping (``algorithm/topology.py:294-383``), the
    per-chain trigonometric sums (``:1096-1099``) and ``_conclude`` (``:1101-1129``).
    """
    traj = universe.trajectory
    frames = np.arange(*slice(start, stop, step).indices(len(traj)))
    dims = universe.dimensions[:3].copy()
    wavevectors = np.stack(
        np.meshgrid(*[2 * np.pi * np.arange(n_points) / L for L in dims]), -1
    ).reshape(-1, 3)
    wavenumbers = np.linalg.norm(wavevectors, axis=1)
    wn_out = np.unique(wavenumbers.round(11))
    scsf = np.zeros(len(wavevectors))
    old = images = None
    for f in frames:
        traj[int(f)]
        positions = group.positions.astype(np.float32)
        if unwrap:
            if old is None:
                old = positions.copy()
                images = np.zeros(positions.shape, dtype=int)
            dpos = positions - old
            mask = np.abs(dpos) >= dims / 2
            images[mask] -= np.sign(dpos[mask]).astype(int)
            old[:] = positions[:]
            positions += images * dims          # in place: stays float32, as in the reference
        for chain in positions.reshape((n_chains, n_monomers, 3)):
            arg = np.einsum("ij,kj->ki", wavevectors, chain)
            scsf += np.sin(arg).sum(axis=0) ** 2 + np.cos(arg).sum(axis=0) ** 2
    scsf /= n_chains * n_monomers * len(frames)
    out = np.fromiter((scsf[np.isclose(q, wavenumbers)].mean() for q in wn_out),
                      dtype=float, count=len(wn_out))
    return {"wavenumbers": wn_out, "scsf": out, "wavevectors": wavevectors}
