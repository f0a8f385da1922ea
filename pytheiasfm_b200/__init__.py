"""theia-b200: B200-native bundle adjustment and RANSAC verification behind the pyTheia API.

Product path = libtheia_b200.so (hand-written sm_100a CUDA behind the C-ABI of
include/theia_b200.h). No CPU fallback exists: see capi.load_library().
"""
from . import capi  # noqa: F401

__version__ = "0.1.0"
