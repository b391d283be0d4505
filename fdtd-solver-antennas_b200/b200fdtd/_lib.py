"""ctypes binding of libb200fdtd.so (include/b200fdtd.h).

The CUDA library is the product: there is no CPU fallback.  If the shared object is missing
this module raises at load time with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(_PKG, "..", "csrc"))
SO_PATH = os.environ.get("B200FDTD_LIB") or os.path.join(CSRC, "libb200fdtd.so")   # override: kernel experiments only
INCLUDE = os.path.normpath(os.path.join(_PKG, "..", "..", "include"))

c_f = C.POINTER(C.c_float)
c_d = C.POINTER(C.c_double)
c_i64 = C.POINTER(C.c_int64)
c_i32 = C.POINTER(C.c_int32)
vp = C.c_void_p


class PmlBox(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("z0", C.c_int32),
                ("bx", C.c_int32), ("by", C.c_int32), ("bz", C.c_int32),
                ("flux_v", vp), ("flux_i", vp), ("vv", vp), ("vvfo", vp), ("vvfn", vp),
                ("ii", vp), ("iifo", vp), ("iifn", vp),
                ("xvecs_v", vp), ("meta_v", vp), ("nvec_v", C.c_int32),
                ("xvecs_i", vp), ("meta_i", vp), ("nvec_i", C.c_int32)]


class Nf2ffSrcFace(C.Structure):
    _fields_ = [("normal", C.c_int32), ("side", C.c_int32), ("na", C.c_int32), ("nb", C.c_int32), ("coord", C.c_double),
                ("acc", vp), ("comp_stride", C.c_int64), ("xa", c_d), ("xb", c_d), ("wa", c_d), ("wb", c_d)]


class Nf2ffFace(C.Structure):
    _fields_ = [("normal", C.c_int32), ("plane", C.c_int32), ("a0", C.c_int32), ("a1", C.c_int32),
                ("b0", C.c_int32), ("b1", C.c_int32), ("acc", vp)]


# name -> (restype, argtypes); every symbol declared in include/b200fdtd.h
SYMBOLS = {
    "b200fdtd_last_error": (C.c_char_p, []),
    "b200fdtd_version": (C.c_int, []),
    "b200fdtd_launch_count": (C.c_int64, []),
    "b200fdtd_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "b200fdtd_destroy": (C.c_int, [vp]),
    "b200fdtd_bind_fields": (C.c_int, [vp, vp, vp]),
    "b200fdtd_bind_coeffs": (C.c_int, [vp, vp, vp, vp, vp]),
    "b200fdtd_set_tuning": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
    "b200fdtd_set_row_compression": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, c_i64, c_i64]),
    "b200fdtd_expand_rows": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "b200fdtd_set_excitation": (C.c_int, [vp, C.c_int64, c_i64, c_f, c_i32, c_f, C.c_int32]),
    "b200fdtd_set_mur": (C.c_int, [vp, C.c_int64, c_i64, c_i64, c_f]),
    "b200fdtd_set_pml": (C.c_int, [vp, C.c_int, C.POINTER(PmlBox)]),
    "b200fdtd_pml_compression_info": (C.c_int, [vp, c_i64, c_i64]),
    "b200fdtd_set_probes": (C.c_int, [vp, C.c_int, c_i32, c_i64, c_i64, c_f, C.c_int, C.c_int, vp, C.c_int, c_d, vp, C.c_double]),
    "b200fdtd_set_nf2ff": (C.c_int, [vp, C.c_int, C.POINTER(Nf2ffFace), C.c_int, c_d, C.c_int, C.c_double,
                                     c_f, c_f, c_f, c_f, c_f, c_f]),
    "b200fdtd_set_nf2ff_td": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.c_int]),
    "b200fdtd_nf2ff_td_dft": (C.c_int, [vp, C.c_int, C.c_int, c_d, C.c_int, vp]),
    "b200fdtd_get_timestep": (C.c_int, [vp, c_i64]),
    "b200fdtd_set_timestep": (C.c_int, [vp, C.c_int64]),
    "b200fdtd_run": (C.c_int, [vp, C.c_int64, C.c_int]),
    "b200fdtd_half_step": (C.c_int, [vp, C.c_int]),
    "b200fdtd_half_step_part": (C.c_int, [vp, C.c_int, C.c_int]),
    "b200fdtd_update_only": (C.c_int, [vp, C.c_int]),
    "b200fdtd_bind_alt_fields": (C.c_int, [vp, vp, vp]),
    "b200fdtd_fused_step_part": (C.c_int, [vp, C.c_int]),
    "b200fdtd_current_copy": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200fdtd_reset_current_copy": (C.c_int, [vp]),
    "b200fdtd_plan_info": (C.c_int, [vp, c_i64, c_i64, c_i64]),
    "b200fdtd_set_he_tuning": (C.c_int, [vp, C.c_int, C.c_int]),
    "b200fdtd_he_info": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "b200fdtd_energy": (C.c_int, [vp, c_d]),
    "b200fdtd_sync": (C.c_int, [vp]),
    "b200fdtd_num_samples": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "b200fdtd_nf2ff_sources": (C.c_int, [C.c_int, vp, C.c_int, C.POINTER(Nf2ffSrcFace), C.c_double, c_d, vp, vp, vp, c_d]),
    "b200fdtd_farfield": (C.c_int, [C.c_int, vp, C.c_int64, vp, vp, vp, C.c_double, C.c_int, c_d, c_d, vp]),
}

_LIB = None


class B200FDTDError(RuntimeError):
    pass


def build_command():
    return ("nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "
            f"-I{INCLUDE} -shared -Xcompiler -fPIC -o {SO_PATH} {os.path.join(CSRC, 'b200fdtd.cu')}")


def lib():
    """Load libb200fdtd.so (ctypes releases the GIL for the duration of every call)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise B200FDTDError(
                "libb200fdtd.so is not built; the CUDA engine is the only implementation (no CPU fallback). "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or:\n  " + build_command())
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        msg = lib().b200fdtd_last_error()
        raise B200FDTDError(msg.decode() if msg else "b200fdtd call failed")
