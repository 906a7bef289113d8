import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (B200); run with `-m gpu` on the GPU box"
    )


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(GOLDEN / f"{name}.npz", allow_pickle=False))
    return load


def universe_from(g, dims=None):
    from mdhelper_b200.universe import SyntheticUniverse
    return SyntheticUniverse(g["positions"], g["dims"] if dims is None else dims)


class FileReader:
    """A reader in the style of MDAnalysis' on-disk readers (what ``Universe.trajectory``
    is for a DCD / NetCDF file): frames are read one at a time from a memory-mapped file,
    there is no whole-trajectory array, ``trajectory[i]`` seeks and returns the timestep.
    Exercises the frame feeder's staged path (structure.py:796 seeks ``self._trajectory[frame]``
    the same way)."""

    def __init__(self, path, n_frames, n_atoms, dims, dt=1.0):
        self._mm = np.memmap(path, dtype=np.float32, mode="r", shape=(n_frames, n_atoms, 3))
        self.n_frames, self.n_atoms, self.dt = n_frames, n_atoms, dt
        self._dims = np.asarray(dims, np.float32)
        self._frame = 0
        self.reads = 0

    def __len__(self):
        return self.n_frames

    def _ts(self, frame):
        from mdhelper_b200.universe import Timestep
        self.reads += 1
        d = self._dims if self._dims.ndim == 1 else self._dims[frame]
        return Timestep(frame, frame * self.dt, np.array(self._mm[frame]), d.copy())

    @property
    def ts(self):
        return self._ts(self._frame)

    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            item = int(item)
            if not 0 <= item < self.n_frames:
                raise IndexError(item)
            self._frame = item
            return self._ts(item)
        raise TypeError("FileReader only seeks single frames")

    def check_slice_indices(self, start, stop, step):
        return slice(start, stop, step).indices(self.n_frames)


def file_universe(tmp_path, positions, dims):
    """A SyntheticUniverse whose trajectory is a :class:`FileReader` over a file holding
    ``positions``."""
    from mdhelper_b200.universe import SyntheticUniverse
    path = tmp_path / "traj.f32"
    np.asarray(positions, np.float32).tofile(path)
    u = SyntheticUniverse(positions[:1], np.atleast_2d(dims)[0])
    u.trajectory = FileReader(path, positions.shape[0], positions.shape[1], dims)
    return u
