"""ctypes binding of libnatrix_b200.so (the C ABI declared in include/natrix_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call fails (for
example because no CUDA device is present), a ``NatrixError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# NATRIX_B200_LIB: another build of the same library (A/B timing of kernel variants, scripts/ab_build.sh)
LIB_PATH = Path(os.environ.get("NATRIX_B200_LIB") or _PKG / "libnatrix_b200.so")

# enum natrix_field / natrix_option (include/natrix_b200.h)
VELOCITY, PRESSURE, DIVERGENCE, VORTICITY, OBSTACLES, NBMASK, DIV4 = range(7)
(OPT_PIPELINE, OPT_JACOBI_DEPTH, OPT_TIMING, OPT_WARM_START, OPT_PACKED, OPT_JACOBI_KERNEL, OPT_SMEM_DEPTH, OPT_SOLVER,
 OPT_SOR_OMEGA_MILLI, OPT_MG_SMOOTH) = range(10)

FIELD_COMPONENTS = {VELOCITY: 2, PRESSURE: 1, DIVERGENCE: 1, VORTICITY: 1, OBSTACLES: 2, NBMASK: 1, DIV4: 1}

# every symbol include/natrix_b200.h declares: name -> (restype, argtypes)
_vp, _f, _i, _d, _sz = C.c_void_p, C.c_float, C.c_int, C.c_double, C.c_size_t
_pvp, _psz, _pi, _pd, _pf = C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)
SIGNATURES = {
    "natrix_create": (_i, [_i, _i, _i, _pvp]),
    "natrix_create_slab": (_i, [_i, _i, _i, _i, _i, _i, _pvp]),
    "natrix_destroy": (_i, [_vp]),
    "natrix_set_params": (_i, [_vp, _f, _i, _f, _f, _d, _i]),
    "natrix_set_option": (_i, [_vp, _i, _i]),
    "natrix_get_option": (_i, [_vp, _i, _pi]),
    "natrix_add_velocity": (_i, [_vp, _f, _f, _f, _f, _f]),
    "natrix_add_circle_obstacle": (_i, [_vp, _f, _f, _f, _i]),
    "natrix_add_triangle_obstacle": (_i, [_vp, _f, _f, _f, _f, _f, _f, _i]),
    "natrix_step": (_i, [_vp, _f]),
    "natrix_comm_unique_id": (_i, [_vp]),
    "natrix_comm_init": (_i, [_vp, _vp, _i, _i]),
    "natrix_comm_stats": (_i, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "natrix_step_phase": (_i, [_vp, _i, _f, _i]),
    "natrix_halo_rows_needed": (_i, [_vp, _i, _f]),
    "natrix_halo_region": (_i, [_vp, _i, _i, _i, _pvp, _pvp, _psz]),
    "natrix_field_ptr": (_i, [_vp, _i, _pvp, _psz]),
    "natrix_copy_out": (_i, [_vp, _i, _vp, _sz]),
    "natrix_copy_in": (_i, [_vp, _i, _vp, _sz]),
    "natrix_field_stats": (_i, [_vp, _i, _pd]),
    "natrix_dye_create": (_i, [_vp, _i, _i, _pvp]),
    "natrix_dye_create_slab": (_i, [_vp, _i, _i, _i, _i, _i, _pvp]),
    "natrix_dye_halo_rows_needed": (_i, [_vp, _i, _f, _f]),
    "natrix_dye_halo_region": (_i, [_vp, _i, _i, _pvp, _pvp, _psz]),
    "natrix_dye_destroy": (_i, [_vp]),
    "natrix_dye_add": (_i, [_vp, _f, _f, _f, _f]),
    "natrix_dye_step": (_i, [_vp, _f, _f, _f]),
    "natrix_dye_field_ptr": (_i, [_vp, _pvp, _psz]),
    "natrix_dye_copy_out": (_i, [_vp, _vp, _sz]),
    "natrix_dye_copy_in": (_i, [_vp, _vp, _sz]),
    "natrix_dye_stats": (_i, [_vp, _pd]),
    "natrix_dye_export_rgba8": (_i, [_vp, _vp, _sz, _i]),
    "natrix_render_frame": (_i, [_vp, _vp, _sz, _i, _f]),
    "natrix_sync": (_i, [_vp]),
    "natrix_stream": (_i, [_vp, _pvp]),
    "natrix_comm_stream": (_i, [_vp, _pvp]),
    "natrix_get_timings": (_i, [_vp, _pf, _i]),
    "natrix_launch_count": (_i, [_vp, C.POINTER(C.c_ulonglong)]),
    "natrix_debug_plan_tiles": (_i, [_i, _i, _i, _i, _pi, _i, _i, _pi, _i]),
    "natrix_debug_plan_stats": (_i, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "natrix_last_error": (C.c_char_p, []),
    "natrix_version": (C.c_char_p, []),
}


class NatrixError(RuntimeError):
    """A libnatrix_b200 call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libnatrix_b200 error {code}: {message}")
        self.code = code


def build(force: bool = False) -> Path:
    """Compile libnatrix_b200.so in-tree with nvcc for sm_100a (natrix_b200/csrc/Makefile)."""
    srcdir = _PKG / "csrc"
    if force:
        subprocess.check_call(["make", "-s", "-C", str(srcdir), "clean"])
    subprocess.check_call(["make", "-s", "-C", str(srcdir)])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NatrixError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                  "g.build()'` (nvcc, sm_100a); there is no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise NatrixError(rc, lib().natrix_last_error().decode("utf-8", "replace"))
    return rc


def loaded_library_path() -> str:
    return str(LIB_PATH)
