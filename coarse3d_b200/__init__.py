"""coarse3d_b200 -- B200-native (sm_100a) implementation of COARSE3D's per-scan
hot path: range projection, prototype contrastive loss (fwd/bwd) + EMA
prototype update, KNN label post-processing.  See DESIGN.md.

Importing the package loads libcoarse3d_b200.so; if it has not been built the
import fails (there is no fallback).
"""
from . import _lib  # noqa: F401  (loads the shared library or raises)
from . import ops  # noqa: F401
from .install import install  # noqa: F401

__version__ = "0.1.0"
