"""
Simulation trajectory analysis (GPU hot path only)
==================================================
Mirrors ``mdhelper.analysis`` for the two classes on the accelerated path.
"""

from . import base, structure

__all__ = ["base", "structure"]
