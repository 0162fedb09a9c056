"""justpic.jl_b200 -- B200-native (sm_100a) implementation of JustPIC.jl's
particle-in-cell hot path behind the reference's API names.

The compute lives in ``libjustpic_sm100a.so`` (hand-written CUDA, C ABI in
``include/justpic_c.h``); this package is the thin host mirror of the Julia
API.  Importing works without a GPU (so the build / symbol checks can run);
any compute call without the built library or a CUDA device raises.
"""
from . import _cabi
from .api import *  # noqa: F401,F403
from .api import move_stats, inject_stats, last_move_path, last_move_reasons, advect_affine_level  # noqa: F401

__version__ = "0.1.0"
