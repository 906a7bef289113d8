"""
Host-side transforms of a finished g(r)
=======================================

O(n_bins) / O(n_bins x n_q) numpy + scipy work on the *result* of the GPU pass; none of
it is on the hot path, it is here so that ``RadialDistributionFunction`` keeps the
reference's whole results API (``results.coordination_numbers``, ``results.pmf``,
``results.wavenumbers`` / ``results.ssf``).  Each function follows the operation order of
its reference counterpart in ``/root/reference/src/mdhelper/analysis/structure.py`` so
that the same g(r) gives the same numbers (checked against the real functions in
``tests/test_host.py`` whenever ``/root/reference`` is present, and against
``tests/golden/rdf_post.npz``).
"""

import warnings

import numpy as np
from scipy.integrate import simpson
from scipy.signal import argrelextrema
from scipy.special import jv

# N_A k_B in kJ / (mol K): both constants are exact SI defining values
_MOLAR_GAS_CONSTANT_KJ = 6.02214076e23 * 1.380649e-23 / 1000.0


def zeroth_order_hankel_transform(r: np.ndarray, f: np.ndarray,
                                  q: np.ndarray) -> np.ndarray:
    r"""
    Hankel transform :math:`F_0(q)=\int_0^\infty f(r)J_0(qr)r\,dr` (times
    :math:`2\pi`) of sampled data, Simpson's rule (reference: ``structure.py:106-146``).
    """
    q = np.asarray(q)
    if q.ndim == 0:
        return 2 * np.pi * simpson(f * r * (jv(0, q * r) if q != 0 else 1.0), r)
    # one row per wavenumber.  (The reference writes ``jv(0, q * r)``, which only
    # broadcasts for a scalar ``q``; an array of wavenumbers -- what
    # ``calculate_structure_factor`` passes for 2-D systems -- raises there.)
    ht = 2 * np.pi * simpson(f * r * jv(0, np.outer(q, r)), x=r)
    if 0 in q:
        ht[q == 0] = 2 * np.pi * simpson(f * r, r)
    return ht


def radial_fourier_transform(r: np.ndarray, f: np.ndarray,
                             q: np.ndarray) -> np.ndarray:
    r"""
    Radial Fourier transform :math:`\hat f(q)=\frac{4\pi}{q}\int_0^\infty f(r)\,r
    \sin(qr)\,dr` of sampled data, Simpson's rule (reference:
    ``structure.py:148-188``).
    """
    rft = 4 * np.pi * np.divide(simpson(f * r * np.sin(np.outer(q, r)), x=r), q)
    if 0 in q:
        rft[q == 0] = 4 * np.pi * simpson(f * r ** 2, x=r)
    return rft


def calculate_coordination_numbers(bins: np.ndarray, rdf: np.ndarray, rho: float, *,
                                   n_coord_nums: int = 2, n_dims: int = 3,
                                   threshold: float = 0.1) -> np.ndarray:
    """
    Coordination numbers: integrals of :math:`g(r)` between its successive local
    minima that are at least ``threshold`` high (reference: ``structure.py:190-285``).
    Entries that cannot be determined are NaN.
    """
    if n_dims not in {2, 3}:
        raise ValueError("Invalid number of dimensions.")

    def shell(r, lo, hi):
        if n_dims == 3:
            return 4 * np.pi * rho * simpson(r ** 2 * rdf[lo:hi], r)
        return 2 * np.pi * rho * simpson(r * rdf[lo:hi], r)

    out = np.full(n_coord_nums, np.nan)
    i_min, = argrelextrema(rdf, np.less)
    i_min = i_min[rdf[i_min] >= threshold]
    if len(i_min) == 0:
        warnings.warn("No local minima found.")
        return out
    out[0] = shell(bins[:i_min[0] + 1], None, i_min[0] + 1)
    for i in range(min(n_coord_nums, len(i_min)) - 1):
        out[i + 1] = shell(bins[i_min[i]:i_min[i + 1] + 1], i_min[i], i_min[i + 1] + 1)
    return out


def calculate_structure_factor(r: np.ndarray, g: np.ndarray, equal: bool, rho: float,
                               x_i: float = 1, x_j: float = None,
                               q: np.ndarray = None, *, q_lower: float = None,
                               q_upper: float = None, n_q: int = 1_000,
                               n_dims: int = 3, formalism: str = "FZ") -> tuple:
    """
    (Partial) static structure factor from a radial distribution function by a
    radial Fourier (3-D) or Hankel (2-D) transform of :math:`g(r)-1` (reference:
    ``structure.py:287-442``; formalisms ``"FZ"``, ``"AL"``, ``"general"``).
    """
    if q is None:
        if q_lower is None:
            q_lower = 2 * np.pi / r[-1]
        if q_upper is None:
            q_upper = 2 * np.pi / r[0]
        q = np.linspace(q_lower, q_upper,
                        int((q_upper - q_lower) / q_lower) if n_q is None else n_q)
    if n_dims == 3:
        transform = radial_fourier_transform
    elif n_dims == 2:
        transform = zeroth_order_hankel_transform
    else:
        raise ValueError("Invalid number of dimensions.")
    rho_sft = rho * transform(r, g - 1, q)
    if equal or formalism == "FZ":
        return q, 1 + rho_sft
    if formalism == "AL":
        return q, (x_i == x_j) + np.sqrt(x_i * x_j) * rho_sft
    if formalism == "general":
        return q, 1 + x_i * x_j * rho_sft
    raise ValueError("Invalid formalism.")


def thermal_energy(temperature: float, reduced: bool) -> float:
    """:math:`k_\\mathrm{B}T` in kJ/mol, or ``temperature`` itself in reduced units
    (reference: ``structure.py:942-957``)."""
    if reduced:
        return float(temperature)
    return _MOLAR_GAS_CONSTANT_KJ * float(temperature)
