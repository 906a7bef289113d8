"""
In-memory trajectory carrier
============================

The reference reads trajectories through an ``MDAnalysis.Universe``; that stays
on the host (BASELINE.json north_star) and is not reimplemented here.  What the
GPU analysis classes need from a universe is small, and this module provides a
duck-typed in-memory stand-in for it so that the classes (and the reference's
own classes, in the parity harness) can be driven without MDAnalysis:

* ``universe.trajectory`` -- ``len()``, ``n_frames``, ``trajectory[i]`` (seek,
  returns the timestep), ``trajectory[a:b:c]`` (iterable of timesteps), ``ts``,
  ``check_slice_indices`` (the calls MDAnalysis' ``AnalysisBase._setup_frames``
  makes; reference frame loop: /root/reference/src/mdhelper/analysis/base.py:137-172).
* ``ts.frame``, ``ts.time``, ``ts.positions`` (float32 ``[N, 3]``),
  ``ts.dimensions`` (float32 ``[6]``: lx, ly, lz, alpha, beta, gamma),
  ``ts.volume``.
* ``universe.atoms`` / ``universe.select(indices)`` -- :class:`AtomGroup` with
  ``positions`` (a fresh float32 copy, as in MDAnalysis), ``n_atoms``,
  ``n_residues``, ``n_segments``, ``universe``, ``ix``.

The whole trajectory is one C-contiguous float32 array ``[F, N, 3]``; when torch
sees a GPU it is allocated in pinned host memory so the frame feeder can issue
asynchronous host->device copies straight out of it.
"""

from __future__ import annotations

import numpy as np


class Timestep:
    """One frame: a view into the trajectory arrays (no copy)."""

    __slots__ = ("frame", "time", "positions", "_unitcell")

    def __init__(self, frame, time, positions, unitcell):
        self.frame = frame
        self.time = time
        self.positions = positions
        self._unitcell = unitcell

    @property
    def dimensions(self):
        if self._unitcell is None or (self._unitcell[:3] == 0).all():
            return None
        return self._unitcell

    @property
    def volume(self):
        dims = self.dimensions
        if dims is None:
            return 0.0
        lx, ly, lz, alpha, beta, gamma = np.asarray(dims, dtype=np.float64)
        if alpha == beta == gamma == 90.0:
            return lx * ly * lz
        ca, cb, cg = (np.cos(np.deg2rad(x)) for x in (alpha, beta, gamma))
        return lx * ly * lz * np.sqrt(1 - ca * ca - cb * cb - cg * cg
                                      + 2 * ca * cb * cg)

    @property
    def n_atoms(self):
        return self.positions.shape[0]


class _SlicedTrajectory:
    """``trajectory[slice]`` / ``trajectory[indices]``.  Like MDAnalysis'
    ``FrameIteratorSliced`` / ``FrameIteratorIndices`` it carries ``step`` (slices)
    or ``frames`` (index arrays), which the reference's
    ``IntermediateScatteringFunction._prepare`` reads (``structure.py:1907-1917``)."""

    def __init__(self, trajectory, frames, step=None):
        self._trajectory = trajectory
        self._frames = frames
        if step is not None:
            self.step = step
        else:
            self.frames = np.asarray(frames)

    def __len__(self):
        return len(self._frames)

    def __iter__(self):
        for f in self._frames:
            yield self._trajectory[int(f)]


class MemoryTrajectory:
    """
    Float32 coordinates ``[F, N, 3]`` plus per-frame (or constant) unit cells.
    """

    def __init__(self, positions, dimensions, dt=1.0):
        positions = np.asarray(positions)
        if positions.dtype != np.float32 or positions.ndim != 3 \
                or positions.shape[2] != 3:
            raise ValueError("'positions' must be a float32 array with "
                             "shape (n_frames, n_atoms, 3).")
        if not positions.flags.c_contiguous:
            positions = np.ascontiguousarray(positions)
        self.coordinates = positions
        self.n_frames, self.n_atoms = positions.shape[:2]
        if dimensions is None:
            self.unitcells = None
        else:
            dimensions = np.asarray(dimensions, dtype=np.float32)
            if dimensions.ndim == 1:
                dimensions = np.tile(dimensions, (self.n_frames, 1))
            if dimensions.shape[1] == 3:
                dimensions = np.hstack(
                    (dimensions, np.full((self.n_frames, 3), 90, np.float32))
                )
            if dimensions.shape != (self.n_frames, 6):
                raise ValueError("'dimensions' must have shape (6,), "
                                 "(n_frames, 3) or (n_frames, 6).")
            self.unitcells = np.ascontiguousarray(dimensions, np.float32)
        self.dt = dt
        self._frame = 0

    def __len__(self):
        return self.n_frames

    def _ts(self, frame):
        return Timestep(
            frame, frame * self.dt, self.coordinates[frame],
            None if self.unitcells is None else self.unitcells[frame]
        )

    @property
    def ts(self):
        return self._ts(self._frame)

    @property
    def frame(self):
        return self._frame

    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            item = int(item)
            if item < 0:
                item += self.n_frames
            if not 0 <= item < self.n_frames:
                raise IndexError(f"Index {item} exceeds length of trajectory "
                                 f"({self.n_frames}).")
            self._frame = item
            return self._ts(item)
        if isinstance(item, slice):
            start, stop, step = item.indices(self.n_frames)
            return _SlicedTrajectory(self, np.arange(start, stop, step), step=step)
        item = np.asarray(item)
        if item.dtype == bool:
            item = np.nonzero(item)[0]
        return _SlicedTrajectory(self, item)

    def __iter__(self):
        for f in range(self.n_frames):
            yield self[f]

    def check_slice_indices(self, start, stop, step):
        """Same contract as ``ProtoReader.check_slice_indices``."""
        for name, v in (("start", start), ("stop", stop), ("step", step)):
            if v is not None and not isinstance(v, (int, np.integer)):
                raise TypeError(f"{name} is not an integer")
        if step == 0:
            raise ValueError("Step size is zero")
        return slice(start, stop, step).indices(self.n_frames)


class RingTrajectory(MemoryTrajectory):
    """
    A trajectory of ``n_frames`` frames that cycles through ``ring`` distinct frames held in
    memory: frame ``f`` is ``positions[f % ring]``.  For benchmarks of long runs of large
    systems (BASELINE configs 3 and 5: 1,000 x 6 MB and 500 x 12 MB of coordinates) whose
    host memory is bounded this way (SURVEY.md section 8(d) allows it: "a ring of 32 distinct
    frames may be cycled"); every frame is still read from host memory and copied to the
    device when it is analysed.  ``ring_coordinates`` / ``ring_period`` tell the frame
    feeder how to address it without copies.
    """

    def __init__(self, positions, dimensions, n_frames, dt=1.0):
        super().__init__(positions, dimensions, dt=dt)
        self.ring_coordinates = self.coordinates
        self.ring_period = self.coordinates.shape[0]
        del self.coordinates                  # not one [F, N, 3] array
        self.n_frames = int(n_frames)
        if self.unitcells is not None:
            reps = -(-self.n_frames // self.ring_period)
            self.unitcells = np.ascontiguousarray(
                np.tile(self.unitcells, (reps, 1))[:self.n_frames])

    def _ts(self, frame):
        return Timestep(
            frame, frame * self.dt, self.ring_coordinates[frame % self.ring_period],
            None if self.unitcells is None else self.unitcells[frame]
        )


class AtomGroup:
    """Index set into a :class:`SyntheticUniverse` (duck-types ``mda.AtomGroup``)."""

    def __init__(self, universe, ix):
        self.universe = universe
        self.ix = np.asarray(ix, dtype=np.intp)
        # contiguous ranges let the feeder copy slices without a gather
        self._contiguous = (
            self.ix.size > 0
            and self.ix[-1] - self.ix[0] + 1 == self.ix.size
            and bool(np.all(np.diff(self.ix) == 1))
        )

    @property
    def positions(self):
        return self.universe.trajectory.ts.positions[self.ix]

    @property
    def n_atoms(self):
        return self.ix.size

    @property
    def resindices(self):
        return self.universe._resindices[self.ix]

    @property
    def segindices(self):
        return self.universe._segindices[self.ix]

    @property
    def masses(self):
        return self.universe._masses[self.ix]

    @property
    def n_residues(self):
        return np.unique(self.resindices).size

    @property
    def n_segments(self):
        return np.unique(self.segindices).size

    @property
    def atoms(self):
        return self

    @property
    def dimensions(self):
        return self.universe.dimensions

    def _entities(self, key):
        # one group per residue / segment that has atoms in this group, in index order
        # (what MDAnalysis' ``AtomGroup.residues`` / ``.segments`` iterate over); the
        # reference's ``center_of_mass`` only uses ``entity.atoms`` of them
        # (algorithm/molecule.py:232-236, 270-273)
        return [AtomGroup(self.universe, self.ix[key == k]) for k in np.unique(key)]

    @property
    def residues(self):
        return self._entities(self.resindices)

    @property
    def segments(self):
        return self._entities(self.segindices)

    def center_of_mass(self):
        """Restated third-party ``MDAnalysis.core.groups.AtomGroup.center_of_mass`` [recall]:
        float64 mass-weighted mean of the float32 coordinates.  The reference reaches it for
        residues / segments of unequal size (algorithm/molecule.py:240-241)."""
        w = self.masses.astype(np.float64, copy=False)
        return (self.positions.astype(np.float64) * w[:, None]).sum(axis=0) / w.sum()

    def __len__(self):
        return self.ix.size

    def __eq__(self, other):
        return (isinstance(other, AtomGroup)
                and self.universe is other.universe
                and np.array_equal(self.ix, other.ix))

    def __hash__(self):
        return hash((id(self.universe), self.ix.tobytes()))

    def __getitem__(self, item):
        return AtomGroup(self.universe, np.atleast_1d(self.ix[item]))


class SyntheticUniverse:
    """
    Minimal in-memory universe.

    Parameters
    ----------
    positions : `numpy.ndarray`
        float32 coordinates, shape :math:`(N_\\mathrm{frames},\\,N,\\,3)`.
    dimensions : array-like
        ``(6,)`` for a constant cell or ``(n_frames, 6)`` / ``(n_frames, 3)``.
    resindices, segindices, masses : array-like, optional
        Per-atom topology attributes (default: one residue/segment per atom,
        unit masses).
    """

    def __init__(self, positions, dimensions, *, resindices=None,
                 segindices=None, masses=None, dt=1.0, n_frames=None):
        # n_frames: a longer trajectory that cycles through the frames given (RingTrajectory)
        self.trajectory = (MemoryTrajectory(positions, dimensions, dt=dt)
                           if n_frames is None
                           else RingTrajectory(positions, dimensions, n_frames, dt=dt))
        n = self.trajectory.n_atoms
        self._resindices = (np.arange(n) if resindices is None
                            else np.asarray(resindices, dtype=np.intp))
        self._segindices = (np.arange(n) if segindices is None
                            else np.asarray(segindices, dtype=np.intp))
        self._masses = (np.ones(n) if masses is None
                        else np.asarray(masses, dtype=np.float64))
        self.atoms = AtomGroup(self, np.arange(n))

    @property
    def dimensions(self):
        return self.trajectory.ts.dimensions

    def select(self, ix):
        """Atom group from an index array, slice or boolean mask."""
        if isinstance(ix, slice):
            ix = np.arange(*ix.indices(self.atoms.n_atoms))
        ix = np.asarray(ix)
        if ix.dtype == bool:
            ix = np.nonzero(ix)[0]
        return AtomGroup(self, ix)


def pinned_empty(shape, dtype=np.float32):
    """
    Host array for trajectory data: pinned (page-locked) through torch when a
    GPU is present so that the feeder's cudaMemcpyAsync calls are truly
    asynchronous; plain numpy otherwise.  Returns ``(array, keepalive)``.
    """
    try:
        import torch
        if torch.cuda.is_available():
            tdtype = {np.dtype(np.float32): torch.float32,
                      np.dtype(np.float64): torch.float64,
                      np.dtype(np.int64): torch.int64}[np.dtype(dtype)]
            t = torch.empty(tuple(shape), dtype=tdtype, pin_memory=True)
            return t.numpy(), t
    except ImportError:
        pass
    a = np.empty(shape, dtype=dtype)
    return a, a
