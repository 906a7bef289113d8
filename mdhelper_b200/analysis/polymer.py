"""
Polymeric analysis on B200
==========================

GPU drop-in for the single-chain structure factor of
``mdhelper.analysis.polymer`` (reference:
``/root/reference/src/mdhelper/analysis/polymer.py:805-1129``), the third "next" row of
the scope table: it is the structure-factor kernel with one accumulator per chain,
:math:`\\sum_\\mathrm{chains}|\\rho_\\mathrm{chain}(\\mathbf{q})|^2`.
"""

from __future__ import annotations

import numpy as np

from .base import GpuAnalysisBase, all_reduce_sum
from .structure import (_centers_of_mass, _grouped_mean, _isclose_members,
                        _lattice_indices,
                        _record)


def unwrap(positions: np.ndarray, positions_old: np.ndarray, dimensions: np.ndarray,
           *, thresholds=None, images: np.ndarray = None) -> None:
    """
    In-place unwrapping of particle positions across periodic boundaries from the
    jump since the previous frame (reference: ``algorithm/topology.py:294-383``,
    ``in_place=True`` branch): ``images`` counts the crossings, ``positions_old``
    becomes the wrapped positions of this frame.
    """
    if thresholds is None:
        thresholds = np.min(dimensions) / 2
    dpos = positions - positions_old
    mask = np.abs(dpos) >= thresholds
    images[mask] -= np.sign(dpos[mask]).astype(int)
    positions_old[:] = positions[:]
    positions += images * dimensions


class SingleChainStructureFactor(GpuAnalysisBase):
    r"""
    Single-chain structure factor :math:`S_\mathrm{sc}(q)` of a homopolymer,

    .. math::

       S_\mathrm{sc}(\mathbf{q})=\frac{1}{MN_\mathrm{p}}\sum_{m=1}^M\left\langle
       \left|\sum_{i=1}^{N_\mathrm{p}}\exp(i\mathbf{q}\cdot\mathbf{r}_i)
       \right|^2\right\rangle

    computed on the GPU for all first-octant reciprocal-lattice wavevectors of an
    ``n_points``\ :sup:`3` grid.  Same constructor and results as the reference class
    (``polymer.py:805-1129``).

    Parameters
    ----------
    group : atom group
        The polymer; chains must be consecutive runs of equal length.
    grouping : `str`, default: :code:`"atoms"`
        ``"atoms"`` or ``"residues"`` (monomer centres of mass, computed on the host).
    n_points : `int`, default: :code:`32`
        Wavevector grid points per axis.
    n_chains, n_monomers : `int`, keyword-only, optional
        Number of chains and monomers per chain; default: from the segments of the
        group.
    dimensions : array-like, keyword-only, optional
        Box edges (default: the universe's).
    unwrap : `bool`, keyword-only, default: :code:`False`
        Unwrap the positions on the host before the sums.  For lattice wavevectors of a
        constant box :math:`\exp(i\mathbf{q}\cdot\mathbf{r})` does not depend on the
        periodic image, so this only matters when the box fluctuates.
    parallel, verbose
        As in the reference (``parallel`` is accepted and ignored).
    kernel : `str`, keyword-only, optional
        GPU kernel strategy: ``"lattice_dmma"`` (FP64 matrix unit; what the default
        ``"auto"`` picks for all but tiny grids) or ``"lattice_fp64"`` (scalar DFMA).

    Attributes
    ----------
    results.wavenumbers, results.scsf
        Unique wavenumbers and the single-chain structure factor for them.
    """

    def __init__(self, group, grouping: str = "atoms", n_points: int = 32, *,
                 n_chains: int = None, n_monomers: int = None, dimensions=None,
                 unwrap: bool = False, parallel: bool = False, verbose: bool = True,
                 kernel: str = None, **kwargs) -> None:
        self._group = group
        self._kernel = kernel
        self.universe = group.universe
        super().__init__(self.universe.trajectory, verbose, **kwargs)
        self._parallel = parallel

        if dimensions is not None:
            if len(dimensions) != 3:
                raise ValueError("'dimensions' must have length 3.")
            self._dimensions = np.asarray(dimensions)
        elif self.universe.dimensions is not None:
            self._dimensions = self.universe.dimensions[:3].copy()
        else:
            raise ValueError("No system dimensions found or provided.")

        if grouping not in (groupings := {"atoms", "residues"}):
            emsg = (f"Invalid grouping '{grouping}'. Valid values: "
                    f"{', '.join(groupings)}.")
            raise ValueError(emsg)
        self._grouping = grouping

        if n_chains is None or n_monomers is None:
            self._internal = True
            self._n_chains = int(group.n_segments)
            n_units = group.n_atoms if grouping == "atoms" else group.n_residues
            self._n_monomers = n_units // self._n_chains
        else:
            self._internal = False
            if not isinstance(n_chains, (int, np.integer)):
                emsg = ("The number of chains must be specified when "
                        "the universe does not contain segment "
                        "information.")
                raise ValueError(emsg)
            if not isinstance(n_monomers, (int, np.integer)):
                emsg = ("The number of monomers per chain must be "
                        "specified when the universe does not contain "
                        "segment information.")
                raise ValueError(emsg)
            self._n_chains, self._n_monomers = int(n_chains), int(n_monomers)

        # reference: polymer.py:977-984 (meshgrid 'xy' ordering, float64 grid from the
        # box edges as stored)
        self._n_points = n_points
        self._wavevectors = np.stack(
            np.meshgrid(*[2 * np.pi * np.arange(n_points) / L
                          for L in self._dimensions]), -1
        ).reshape(-1, 3)
        self._wavenumbers = np.linalg.norm(self._wavevectors, axis=1)
        self._lattice_n, self._lattice_b = _lattice_indices(
            self._wavevectors, self._dimensions)
        if self._lattice_n is None:
            raise ValueError("The wavevector grid is not a reciprocal lattice.")
        self._unwrap = unwrap

    def _prepare(self) -> None:
        self.results.wavenumbers = np.unique(self._wavenumbers.round(11))
        self.results.units = {"results.wavenumbers": "angstrom^-1"}
        self.results.scsf = np.zeros(len(self._wavevectors))

    def _begin(self, frames: np.ndarray):
        ctx = self._context()
        n = self._n_chains * self._n_monomers
        ctx.sq_configure(n, [0, n], self._wavevectors, [(-1, -1)],
                         lattice_n=self._lattice_n, lattice_b=self._lattice_b,
                         mode=self._kernel or "auto")
        ctx.sq_configure_chains(self._n_chains, self._n_monomers)
        g = self._group
        if self._grouping == "atoms" and not self._unwrap:
            if g.n_atoms != n:
                raise ValueError("n_chains * n_monomers does not match the group.")
            return [g.ix], None, 12 * n

        state = {}

        def positions_fn(ts):
            pos = ts.positions[g.ix]
            if self._grouping == "residues":
                if self._internal:
                    pos = _centers_of_mass(g, "residues", pos)
                else:          # polymer.py:1083-1092: equal-sized monomers
                    m = np.asarray(g.masses, dtype=np.float64).reshape(
                        self._n_chains, self._n_monomers, -1)
                    p = pos.reshape(self._n_chains, self._n_monomers, -1, 3)
                    pos = (np.einsum("...a,...ad->...d", m, p)
                           / m.sum(axis=-1, keepdims=True)).reshape(-1, 3)
            # atoms: the reference unwraps the reader's float32 array in place
            # (polymer.py:1079, 1091-1093 -- the unwrapped coordinates are rounded to
            # float32); centres of mass are float64 and stay so
            if self._grouping == "residues":
                pos = np.asarray(pos, dtype=np.float64)
            else:
                pos = np.array(pos, dtype=np.float32)
            if self._unwrap:
                if "old" not in state:          # the first analysed frame is the anchor
                    state["old"] = pos.copy()
                    state["images"] = np.zeros(pos.shape, dtype=int)
                else:
                    unwrap(pos, state["old"], self._dimensions,
                           thresholds=self._dimensions / 2, images=state["images"])
            return [pos]
        return [np.arange(n)], positions_fn, 12 * n

    def _staging_dtype(self, positions_fn):
        # centres of mass (unwrapped or not) stay float64 up to the trigonometric sums
        # (polymer.py:1076-1099); atom coordinates are float32 in the reference too
        return np.float64 if self._grouping == "residues" else np.float32

    def _consume(self, batch, device: bool = False) -> None:
        self._ctx.sq_accumulate(batch.ptrs[0], batch.strides[0], batch.n_frames,
                                device=device, keepalive=batch.keepalive, f64=batch.f64)
        _record(batch)

    def _finish(self) -> None:
        self._local = self._ctx.sq_fetch()[0]

    def run(self, start: int = None, stop: int = None, step: int = None,
            frames=None, verbose: bool = None, **kwargs):
        if self._unwrap:
            # unwrapping follows the particles from frame to frame: no frame sharding
            self._setup_frames(self._trajectory, start=start, stop=stop, step=step,
                               frames=frames)
            self._prepare()
            from .base import world
            rank, _ = world()
            local = self._frame_list if rank == 0 else self._frame_list[:0]
            self.n_local_frames = len(local)
            self._process(local)
            self._conclude()
            return self
        return super().run(start=start, stop=stop, step=step, frames=frames,
                           verbose=verbose, **kwargs)

    def _conclude(self) -> None:
        # reference: polymer.py:1101-1129
        scsf = all_reduce_sum(self._local, self._device)
        scsf = scsf / (self._n_chains * self._n_monomers * self.n_frames)
        # same index sets as the reference's np.isclose(q, wavenumbers) scan
        if getattr(self, "_members", None) is None:
            self._members = _isclose_members(self.results.wavenumbers, self._wavenumbers)
        self.results.scsf = _grouped_mean(scsf, self._members)
