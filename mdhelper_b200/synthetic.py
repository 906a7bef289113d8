"""
Synthetic trajectories of the BASELINE.json configurations
==========================================================

Seeded Lennard-Jones-like / electrolyte / melt trajectories in reduced units
(SURVEY.md section 8(d)): a jittered simple-cubic lattice that performs a small
random walk from frame to frame, wrapped into ``[0, L)`` and stored as float32.
They only have to exercise the kernels with liquid-like pair statistics; no
force field is integrated.

The coordinate array is pinned host memory when a GPU is present, so the frame
feeder can issue asynchronous copies straight out of it.
"""

from __future__ import annotations

import numpy as np

from .universe import SyntheticUniverse, pinned_empty


def box_edge(n: int, rho: float) -> np.float32:
    """Cubic box edge for ``n`` particles at number density ``rho`` (float32)."""
    return np.float32((n / rho) ** (1.0 / 3.0))


def _lattice(n: int, L: float) -> np.ndarray:
    m = int(np.ceil(n ** (1.0 / 3.0) - 1e-9))
    g = (np.arange(m) + 0.5) * (L / m)
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    return xyz[:n], np.indices((m, m, m)).reshape(3, -1).T[:n]


def _wrap_into(out: np.ndarray, pos: np.ndarray, L: np.float32) -> None:
    np.mod(pos, L, out=out)
    out[out >= L] = 0.0           # fmod of a tiny negative can round to L
    out[out < 0] = 0.0


def fluid_positions(n: int, n_frames: int, *, rho: float = 0.8, seed: int = 0,
                    jitter: float = 0.15, step: float = 0.05,
                    order: np.ndarray = None, pinned: bool = True):
    """
    ``(positions float32 [F, n, 3], L float32)`` of a jittered-lattice fluid.
    ``order`` optionally permutes the particles (e.g. cations first).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    L = box_edge(n, rho)
    base, _ = _lattice(n, float(L))
    pos = (base + jitter * rng.standard_normal((n, 3))).astype(np.float32)
    if order is not None:
        pos = pos[order]
    out, keep = _alloc((n_frames, n, 3), pinned)
    for f in range(n_frames):
        pos += np.float32(step) * rng.standard_normal((n, 3), dtype=np.float32)
        _wrap_into(out[f], pos, L)
    return out, L, keep


def _alloc(shape, pinned: bool):
    if pinned:
        return pinned_empty(shape, np.float32)
    a = np.empty(shape, dtype=np.float32)
    return a, a


def lj_fluid(n: int, n_frames: int, *, rho: float = 0.8, seed: int = 0,
             pinned: bool = True):
    """Universe of an ``n``-particle LJ-like fluid (configs 1 and 3)."""
    pos, L, keep = fluid_positions(n, n_frames, rho=rho, seed=seed, pinned=pinned)
    u = SyntheticUniverse(pos, np.array([L, L, L, 90, 90, 90], np.float32))
    u._keepalive = keep
    return u


def electrolyte(n_ions: int, n_frames: int, *, rho: float = 0.8, seed: int = 0,
                pinned: bool = True):
    """
    Universe of ``n_ions`` ions, half cations and half anions with rock-salt
    species assignment on the jittered lattice (config 2).  Cations occupy
    indices ``[0, n/2)``, anions ``[n/2, n)``.  Returns ``(universe, cations,
    anions)``.
    """
    L = box_edge(n_ions, rho)
    _, ijk = _lattice(n_ions, float(L))
    parity = ijk.sum(axis=1) % 2
    order = np.concatenate((np.nonzero(parity == 0)[0], np.nonzero(parity == 1)[0]))
    n_cat = int((parity == 0).sum())
    pos, L, keep = fluid_positions(n_ions, n_frames, rho=rho, seed=seed, order=order,
                                   pinned=pinned)
    u = SyntheticUniverse(pos, np.array([L, L, L, 90, 90, 90], np.float32))
    u._keepalive = keep
    return u, u.select(slice(0, n_cat)), u.select(slice(n_cat, n_ions))


def polymer_melt(n_chains: int, chain_length: int, n_frames: int, *,
                 rho: float = 0.85, bond: float = 0.97, seed: int = 0,
                 pinned: bool = True):
    """
    Universe of a coarse-grained melt: ``n_chains`` random-walk chains of
    ``chain_length`` beads (bond length ``bond``), wrapped (config 5).  One
    residue / segment per chain.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n = n_chains * chain_length
    L = box_edge(n, rho)
    start = rng.random((n_chains, 1, 3)) * float(L)
    steps = rng.standard_normal((n_chains, chain_length, 3))
    steps *= bond / np.linalg.norm(steps, axis=2, keepdims=True)
    steps[:, 0] = 0
    pos = (start + np.cumsum(steps, axis=1)).reshape(n, 3).astype(np.float32)
    out, keep = _alloc((n_frames, n, 3), pinned)
    for f in range(n_frames):
        pos += np.float32(0.05) * rng.standard_normal((n, 3), dtype=np.float32)
        _wrap_into(out[f], pos, L)
    chain = np.repeat(np.arange(n_chains), chain_length)
    u = SyntheticUniverse(out, np.array([L, L, L, 90, 90, 90], np.float32),
                          resindices=chain, segindices=chain)
    u._keepalive = keep
    return u
