"""
Simulation trajectory analysis (GPU hot path only)
==================================================
Mirrors ``mdhelper.analysis`` for the two classes on the accelerated path.
"""

from . import base, structure
from .base import CombinedAnalysis

__all__ = ["base", "structure", "CombinedAnalysis"]
