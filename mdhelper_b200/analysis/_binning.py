"""
Exact histogram thresholds on squared distances.

The reference bins ``d = sqrt(d2)`` with ``numpy.histogram(d, bins=n_bins,
range=range)`` after ``capped_distance`` kept ``min_cutoff < d <= max_cutoff``
(``/root/reference/src/mdhelper/analysis/structure.py:93-104``).  numpy's
uniform-bin path ends in a fix-up against the ``np.linspace`` edge array
(``numpy/lib/_histograms_impl.py``: ``decrement``/``increment``), so the final
rule is: bin ``k`` iff ``edges[k] <= d < edges[k+1]``, the last bin closed.

IEEE ``sqrt`` is monotone non-decreasing, so ``{x : sqrt(x) >= e}`` is an up-set
of the doubles and the same rule can be evaluated on ``d2`` against
``T[k] = min{x : sqrt(x) >= edges[k]}`` -- no square root (and no division) in
the GPU inner loop, and bit-identical counts.
"""

import numpy as np


def _min_x_sqrt_ge(e: np.ndarray) -> np.ndarray:
    """Smallest non-negative double ``x`` with ``sqrt(x) >= e`` (elementwise)."""
    e = np.asarray(e, dtype=np.float64)
    x = e * e
    with np.errstate(invalid="ignore"):
        for _ in range(16):
            xm = np.nextafter(x, -np.inf)
            m = (x > 0) & (np.sqrt(xm) >= e)
            if not m.any():
                break
            x = np.where(m, xm, x)
        for _ in range(16):
            m = np.sqrt(x) < e
            if not m.any():
                break
            x = np.where(m, np.nextafter(x, np.inf), x)
        ok = (np.sqrt(x) >= e) & ((x == 0)
                                  | (np.sqrt(np.nextafter(x, -np.inf)) < e))
    if not ok.all():
        raise ValueError("could not bracket the squared histogram edges")
    return x


def squared_thresholds(n_bins: int, range_: tuple) -> np.ndarray:
    """
    ``T[0..n_bins]`` such that a pair lands in bin ``k`` iff
    ``T[k] <= d2 < T[k+1]``; ``d2 >= T[n_bins]`` or ``d2 < T[0]`` is not counted.
    """
    lo, hi = float(range_[0]), float(range_[1])
    if not (n_bins >= 1 and hi > lo and lo >= 0):
        raise ValueError("invalid histogram range or number of bins")
    edges = np.linspace(lo, hi, n_bins + 1)
    T = np.empty(n_bins + 1)
    T[:n_bins] = _min_x_sqrt_ge(edges[:-1])
    # capped_distance: d > range[0] - eps  (structure.py:94-95)
    lo_cut = lo - np.finfo(np.float64).eps
    if lo_cut >= 0:
        T[0] = max(T[0], _min_x_sqrt_ge(np.nextafter(lo_cut, np.inf))[()])
    # capped_distance: d <= range[1]; histogram: d <= edges[-1] (last bin closed)
    top = min(hi, edges[-1])
    T[n_bins] = _min_x_sqrt_ge(np.nextafter(top, np.inf))[()]
    if not np.all(np.diff(T) > 0):
        raise ValueError("histogram bins are too narrow to resolve in float64")
    return T
