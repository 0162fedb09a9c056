"""ctypes binding of libjustpic_sm100a.so (include/justpic_c.h).

The library is the product: if it is missing or fails to load this module
raises -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

HERE = Path(__file__).resolve().parent
# JUSTPIC_LIB: developer override used by tools/ to A/B two builds of the SAME library
LIB_PATH = Path(os.environ["JUSTPIC_LIB"]) if os.environ.get("JUSTPIC_LIB") else HERE / "libjustpic_sm100a.so"

JP_MAX_ARGS = 16
JP_MAX_SLOTS = 64          # occupancy-word kernels
JP_REGION_ALL, JP_REGION_SHELL, JP_REGION_INTERIOR = 0, 1, 2
JP_MAX_SLOTS_WIDE = 1024   # largest max_xcell accepted (chunked launches + literal move / inject above 64)
JP_MAX_PHASES = 32
JP_OPT_P2G_MODE = 1
JP_OPT_MOVE_MODE = 2
JP_OPT_ADVECT_AFFINE = 3
JP_OPT_MOVE_POLICY = 4
JP_OPT_ADVECT_CLASSIFY = 5
JP_OPT_LAST_CLASSIFY = 6
JP_OPT_MOVE_INTERP = 7
JP_OPT_LAST_INTERP = 8
JP_OPT_PROFILE = 9
JP_OPT_GRAPH_STEP_OFFSET = 10
JP_F64, JP_F32, JP_BOOL = 0, 1, 2
JP_LAYOUT_TO_HOST, JP_LAYOUT_TO_DEVICE = 0, 1
JP_MOVE_POLICY_REFERENCE, JP_MOVE_POLICY_COMPACT, JP_MOVE_POLICY_DENSE = 0, 1, 2
JP_MOVE_AUTO, JP_MOVE_DIRECT = 0, 1
JP_P2G_EXACT, JP_P2G_TWOPASS, JP_P2G_TWOPASS_FASTW = 0, 1, 2

c_double_p = C.POINTER(C.c_double)


class GridDesc(C.Structure):
    """jp_grid_desc (HOST pointers)."""
    _fields_ = [
        ("ndim", C.c_int32),
        ("n", C.c_int32 * 3),
        ("S", C.c_int32),
        ("uniform", C.c_int32),
        ("xv", c_double_p * 3),
        ("xc", c_double_p * 3),
        ("xvel", (c_double_p * 3) * 3),
        ("nvel", (C.c_int32 * 3) * 3),
    ]


class ParticlesC(C.Structure):
    """jp_particles (DEVICE pointers)."""
    _fields_ = [("coords", C.c_void_p * 3), ("index", C.c_void_p)]


# every symbol include/justpic_c.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "jp_ctx_create": (C.c_int, [C.POINTER(GridDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "jp_ctx_destroy": (None, [C.c_void_p]),
    "jp_last_error": (C.c_char_p, []),
    "jp_version": (C.c_int, []),
    "jp_set_option": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "jp_get_option": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "jp_init_particles": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_int32, C.c_uint64, C.c_void_p]),
    "jp_advect": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_int32, C.c_double,
                            C.POINTER(C.c_void_p), C.c_double, C.c_void_p]),
    "jp_advect_region": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_int32, C.c_double,
                                   C.POINTER(C.c_void_p), C.c_double, C.c_int32, C.c_void_p]),
    "jp_advect_interp": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_int32, C.c_double,
                                   C.POINTER(C.c_void_p), C.c_double, C.c_int32, C.c_void_p]),
    "jp_move": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "jp_move_interp_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "jp_invalidate_handoffs": (C.c_int, [C.c_void_p]),
    "jp_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "jp_move_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "jp_last_move_path": (C.c_int, [C.c_void_p]),
    "jp_inject": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                            C.c_uint64, C.c_uint32, C.c_void_p]),
    "jp_inject_phase": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, C.c_void_p]),
    "jp_inject_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "jp_clean": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "jp_force_injection": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_double), C.c_int32, C.c_void_p]),
    "jp_grid2particle": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_void_p]),
    "jp_grid2particle_flip": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_void_p]),
    "jp_subgrid_diffusion": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int32,
                                       C.c_void_p]),
    "jp_centroid2particle": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_void_p]),
    "jp_particle2grid": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_void_p]),
    "jp_particle2centroid": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_void_p]),
    "jp_phase_ratios_center": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p]),
    "jp_phase_ratios_vertex": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p]),
    "jp_phase_ratios_face": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_void_p]),
    "jp_phase_ratios_midpoint": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_int32, C.c_void_p]),
    "jp_update_phase_ratios": (C.c_int, [C.c_void_p, C.POINTER(ParticlesC), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "jp_cellarray_permute": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_void_p]),
    "jp_halo_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                   C.c_void_p]),
    "jp_halo_exchange_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "jp_comm_unique_id": (C.c_int, [C.c_void_p]),
    "jp_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "jp_comm_destroy": (C.c_int, [C.c_void_p]),
    "jp_allreduce_max": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "jp_halo_plane_bytes": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32]),
    "jp_halo_pack": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "jp_halo_unpack": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
}

_lib = None


class JustPICError(RuntimeError):
    """Non-zero status from libjustpic_sm100a.so (message from jp_last_error)."""


def load():
    """Load the shared library and declare every prototype.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m justpic.jl_b200._build` "
            "(nvcc, sm_100a).  There is no CPU fallback for the JustPIC hot path."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, who: str = "") -> None:
    if rc != 0:
        msg = load().jp_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{who}: {msg}")   # the reference throws ArgumentError host-side
        raise JustPICError(f"{who} failed (status {rc}): {msg}")
