"""
Simulation trajectory analysis (GPU hot path only)
==================================================
Mirrors ``mdhelper.analysis`` for the classes on the accelerated path (``structure``:
RDF, S(q), F(q, t); ``polymer``: single-chain structure factor).
"""

from . import base, polymer, structure
from .base import CombinedAnalysis

__all__ = ["base", "polymer", "structure", "CombinedAnalysis"]
