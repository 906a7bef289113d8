"""
mdhelper_b200
=============
B200-native (sm_100a) implementation of the per-frame structural-analysis hot
path of bbye98/mdhelper: the minimum-image pair-distance histogram behind
``RadialDistributionFunction`` and the direct-sum static structure factor.

* :mod:`mdhelper_b200.analysis.structure` -- drop-in analysis classes
  (constructor, ``.run(start, stop, step)`` and ``results.*`` as in the reference):
  ``RadialDistributionFunction``, ``StructureFactor``,
  ``IntermediateScatteringFunction``; :mod:`mdhelper_b200.analysis.polymer` --
  ``SingleChainStructureFactor``; ``mdhelper_b200.analysis.CombinedAnalysis`` -- several
  analyses over one upload of every frame.
* :mod:`mdhelper_b200.universe` -- in-memory trajectory carrier (duck-types the
  slice of ``MDAnalysis.Universe`` the classes touch).
* :mod:`mdhelper_b200._lib` -- ctypes binding of ``libmdh_b200.so``
  (C ABI in ``include/mdh_b200.h``; CUDA in ``mdhelper_b200/csrc``).

There is no CPU fallback: the analysis classes raise if the CUDA library or a
CUDA device is missing.
"""

VERSION = "0.1.0"

from . import universe  # noqa: E402,F401
from . import analysis  # noqa: E402,F401

__all__ = ["analysis", "universe", "VERSION"]
