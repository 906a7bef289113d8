"""
Drive the REAL reference classes in this container -- TEST INFRASTRUCTURE ONLY.

``/root/reference`` is pure Python but imports MDAnalysis, pint, matplotlib,
pandas and netCDF4, none of which is installed here.
``load()`` imports ``mdhelper.analysis.structure`` from ``/root/reference/src``
(nothing is copied) behind a meta-path finder that stubs those five roots, with:

* ``MDAnalysis.AtomGroup``  -> ``mdhelper_b200.universe.AtomGroup`` (so the
  reference's ``isinstance`` checks accept the in-memory universe),
* ``MDAnalysis.analysis.base.AnalysisBase`` -> a minimal restatement of the
  third-party frame loop (``_setup_frames`` -> ``_prepare`` -> per frame
  ``_frame_index``/``_ts``/``_single_frame`` -> ``_conclude``),
* ``MDAnalysis.lib.distances.capped_distance`` -> the restated C oracle.

Everything else that runs -- ``radial_histogram``, ``RadialDistributionFunction``
``_prepare/_single_frame/_conclude``, ``StructureFactor`` and the numba kernels
of ``algorithm/accelerated.py`` -- is the reference's own code.  This cannot
travel to the GPU box; it is used to generate ``tests/golden/*.npz`` (see
``tests/golden/make_golden.py``) and by the CPU tests that validate
``oracle/reference_port.py`` when ``/root/reference`` exists.
"""

import importlib
import importlib.abc
import importlib.machinery
import pathlib
import sys
import types

import numpy as np

REFERENCE_SRC = pathlib.Path("/root/reference/src")
_STUB_ROOTS = ("MDAnalysis", "pint", "matplotlib", "pandas", "netCDF4")


def available() -> bool:
    return (REFERENCE_SRC / "mdhelper" / "analysis" / "structure.py").exists()


class _Anything:
    """Absorbs any use a stubbed third-party object is put to at import time."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def _same(self, *a, **k):
        return self

    __pow__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _same
    __getitem__ = _same

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _AnalysisBase:
    """Restated third-party MDAnalysis ``AnalysisBase`` frame loop [recall]."""

    def __init__(self, trajectory, verbose=False, **kwargs):
        self._trajectory = trajectory
        self._verbose = verbose
        self.results = _load_state["Hash"]()

    def _setup_frames(self, trajectory, start=None, stop=None, step=None,
                      frames=None):
        self._trajectory = trajectory
        if frames is not None:
            if not all(o is None for o in (start, stop, step)):
                raise ValueError("start/stop/step cannot be combined with frames")
            sl = frames
            self.start = self.stop = self.step = None
        else:
            start, stop, step = trajectory.check_slice_indices(start, stop, step)
            sl = slice(start, stop, step)
            self.start, self.stop, self.step = start, stop, step
        self._sliced_trajectory = trajectory[sl]
        self.n_frames = len(self._sliced_trajectory)
        self.frames = np.zeros(self.n_frames, dtype=int)
        self.times = np.zeros(self.n_frames)

    def _prepare(self):
        pass

    def _conclude(self):
        pass

    def run(self, start=None, stop=None, step=None, frames=None, verbose=None,
            **kwargs):
        self._setup_frames(self._trajectory, start=start, stop=stop, step=step,
                           frames=frames)
        self._prepare()
        for i, ts in enumerate(self._sliced_trajectory):
            self._frame_index = i
            self._ts = ts
            self.frames[i] = ts.frame
            self.times[i] = ts.time
            self._single_frame()
        self._conclude()
        return self


_load_state = {}


def load():
    """Returns the reference's ``mdhelper.analysis.structure`` module."""
    if "structure" in _load_state:
        return _load_state["structure"]
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    from mdhelper_b200.universe import AtomGroup
    from . import reference_port

    finder = _StubFinder()
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, str(REFERENCE_SRC))
    try:
        pint = importlib.import_module("pint")
        pint.Quantity = type("Quantity", (), {})          # isinstance() target
        pint.UnitRegistry = type("UnitRegistry", (_Anything,), {})
        mda = importlib.import_module("MDAnalysis")
        mda.AtomGroup = AtomGroup
        mda_base = importlib.import_module("MDAnalysis.analysis.base")
        mda_base.AnalysisBase = _AnalysisBase
        mda_dist = importlib.import_module("MDAnalysis.lib.distances")
        mda_dist.capped_distance = reference_port.capped_distance
        mda_lib = importlib.import_module("MDAnalysis.lib")
        mda_lib.distances = mda_dist
        base = importlib.import_module("mdhelper.analysis.base")
        _load_state["Hash"] = base.Hash
        structure = importlib.import_module("mdhelper.analysis.structure")
    finally:
        sys.path.remove(str(REFERENCE_SRC))
    _load_state["structure"] = structure
    _load_state["accelerated"] = importlib.import_module(
        "mdhelper.algorithm.accelerated")
    return structure


def accelerated():
    """Returns the reference's ``mdhelper.algorithm.accelerated`` module."""
    load()
    return _load_state["accelerated"]


def polymer():
    """Returns the reference's ``mdhelper.analysis.polymer`` module (same stubs)."""
    load()
    if "polymer" not in _load_state:
        sys.path.insert(0, str(REFERENCE_SRC))
        try:
            _load_state["polymer"] = importlib.import_module("mdhelper.analysis.polymer")
        finally:
            sys.path.remove(str(REFERENCE_SRC))
    return _load_state["polymer"]
