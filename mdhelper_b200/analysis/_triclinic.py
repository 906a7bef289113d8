"""
Triclinic cells: cell matrix and volume from ``(lx, ly, lz, alpha, beta, gamma)``
=================================================================================

The reference hands ``ts.dimensions`` to ``MDAnalysis.lib.distances.capped_distance``
(``/root/reference/src/mdhelper/analysis/structure.py:93-96``) and reads ``ts.volume``
(``:760``); for cells with angles other than 90 degrees the third-party MDAnalysis code
first turns the six numbers into a lower-triangular matrix
(``MDAnalysis.lib.mdamath.triclinic_vectors``) and computes the volume with
``MDAnalysis.lib.mdamath.box_volume``.  Both are restated here from their published
definitions (MDAnalysis is neither vendored by the reference nor installable in this
image, so this restatement is NOT pinned against it).
"""

from __future__ import annotations

import numpy as np


def is_orthorhombic(dims) -> np.ndarray:
    """Per frame: all three angles are exactly 90 (``dims``: ``[..., 6]`` or ``[..., 3]``)."""
    dims = np.atleast_2d(np.asarray(dims))
    if dims.shape[1] < 6:
        return np.ones(len(dims), dtype=bool)
    return np.all(dims[:, 3:6] == 90, axis=1)


def triclinic_vectors(dimensions) -> np.ndarray:
    """
    ``[3, 3]`` float32 matrix ``[[a_x, 0, 0], [b_x, b_y, 0], [c_x, c_y, c_z]]`` of a cell
    given as ``(lx, ly, lz, alpha, beta, gamma)``; zeros for an invalid cell.  The
    trigonometry runs in float64, right angles use exact 0 / 1, the result is rounded to
    float32 once.
    """
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
    if gamma == 90.0:
        cos_gamma, sin_gamma = 0.0, 1.0
    else:
        g = np.deg2rad(gamma)
        cos_gamma, sin_gamma = np.cos(g), np.sin(g)
    m[1, 0] = ly * cos_gamma
    m[1, 1] = ly * sin_gamma
    m[2, 0] = lz * cos_beta
    m[2, 1] = lz * (cos_alpha - cos_beta * cos_gamma) / sin_gamma
    m[2, 2] = np.sqrt(lz * lz - m[2, 0] ** 2 - m[2, 1] ** 2)
    # the discriminant is positive exactly for angle triplets that span a cell
    if not m[2, 2] > 0.0:
        return np.zeros((3, 3), dtype=np.float32)
    return m.astype(np.float32)


def box_volume(dimensions) -> float:
    """Volume of the cell ``(lx, ly, lz, alpha, beta, gamma)`` in float64 (0 if invalid)."""
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim
    if alpha == beta == gamma == 90.0 and lx > 0 and ly > 0 and lz > 0:
        return float(lx * ly * lz)
    if np.all(dim > 0.0) and alpha < 180.0 and beta < 180.0 and gamma < 180.0:
        ca, cb, cg = (np.cos(np.deg2rad(x)) for x in (alpha, beta, gamma))
        with np.errstate(invalid="ignore"):
            v = lx * ly * lz * np.sqrt(1.0 - ca * ca - cb * cb - cg * cg + 2.0 * ca * cb * cg)
        return 0.0 if np.isnan(v) else float(v)
    return 0.0
