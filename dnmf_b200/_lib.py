"""ctypes binding of include/dnmf_b200.h.  Fails loudly: there is no CPU fallback."""
import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DNMF_B200_LIB") or os.path.join(_HERE, "_C", "libdnmf_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dnmf_b200.h")

_lib = None


class DnmfError(RuntimeError):
    pass


def declared_symbols(header: str = HEADER_PATH):
    """Names of every function declared in the public header."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dnmf_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = True):
    """dlopen the CUDA library (building it with nvcc when it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        if not build_if_missing:
            raise DnmfError("CUDA library %s is missing; run `python -m dnmf_b200.build`" % LIB_PATH)
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    sig = {
        "dnmf_abi_version": (c_int, []),
        "dnmf_last_error": (c_char_p, []),
        "dnmf_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int]),
        "dnmf_destroy": (None, [P]),
        "dnmf_set_footprints": (c_int, [P, P, P, c_float, P]),
        "dnmf_get_ranges": (c_int, [P, P]),
        "dnmf_get_table": (c_int, [P, c_int, P]),
        "dnmf_set_tiling": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int]),
        "dnmf_get_tiling": (c_int, [P, P]),
        "dnmf_set_affine": (c_int, [P, c_int]),
        "dnmf_upload_frames": (c_int, [P, P, c_int, c_int, c_int, P]),
        "dnmf_video_devptr": (c_int, [P, POINTER(c_void_p)]),
        "dnmf_attach_frames": (c_int, [P, P, c_int, P]),
        "dnmf_bin_tiles": (c_int, [P, P, P, c_int, P, P, P, P, c_int64, POINTER(c_int64), P]),
        "dnmf_loss_grad": (c_int, [P, P, P, c_int, c_int, P, P, P, P, P]),
        "dnmf_adam_step": (c_int, [P, P, P, P, P, c_double, c_double, c_double, c_double, c_int64, c_int, P,
                                   c_int, c_int, P, P]),
        "dnmf_motion_step": (c_int, [P, P, P, c_int, c_int, P, P, P, P, c_double, c_double, c_double, c_double,
                                     c_int64, c_int, P, P]),
        "dnmf_motion_epoch": (c_int, [P, P, P, c_int, c_int, P, P, P, P, c_double, c_double, c_double, c_double,
                                      c_int64, c_int, P, P]),
        "dnmf_epoch_mode": (c_int, [P, c_int, POINTER(c_int)]),
        "dnmf_motion_step_host": (c_int, [P, P, P, c_int, c_int, P, P, P, P, c_double, c_double, c_double,
                                          c_double, c_int64, c_int, POINTER(c_double), P]),
        "dnmf_forward": (c_int, [P, P, c_int, P, P, P, P, P, P]),
        "dnmf_mu_stats": (c_int, [P, P, P, c_int, P, P]),
        "dnmf_get_mu_stats": (c_int, [P, c_int, P, P]),
        "dnmf_mu_path": (c_int, [P, c_int, POINTER(c_int)]),
        "dnmf_mu_begin": (c_int, [P, P, P]),
        "dnmf_mu_sweep": (c_int, [P, c_double, c_int, P, P, P]),
        "dnmf_mu_boundary": (c_int, [P, P, P, P]),
        "dnmf_mu_end": (c_int, [P, P, P]),
        "dnmf_mu_sweeps": (c_int, [P, P, c_double, c_int, c_int, P]),
        "dnmf_iwarp": (c_int, [P, P, P, c_int, P, P, P]),
        "dnmf_get_counters": (c_int, [P, P]),
        "dnmf_check_status": (c_int, [P, P]),
        "dnmf_render_cells": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P]),
        "dnmf_update_spatial": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_double, c_int, c_int64, c_int, c_int,
                                        P, P, P]),
        "dnmf_update_temporal_dense": (c_int, [P, P, P, c_double, c_int, c_int64, c_int, c_int, P, P, P]),
        "dnmf_forward_maxz": (c_int, [P, P, c_int, P, P, P]),
        "dnmf_frames_maxz": (c_int, [P, c_int64, c_int, P, P]),
        "dnmf_ext_enable": (c_int, [P]),
        "dnmf_ext_loss_grad": (c_int, [P, P, P, c_int, c_int, P, P, c_float, P, P, P, P, P, P]),
        "dnmf_ext_step_begin": (c_int, [P, P, P, c_int, c_int, P, P, P, P]),
        "dnmf_ext_step_end": (c_int, [P, P, P, P, c_double, c_double, c_double, c_double, c_int64, c_int, P, c_int,
                                      c_double, c_double, c_double, c_float, P, P]),
        "dnmf_ext_set_params": (c_int, [P, c_float, c_int, P]),
        "dnmf_ext_get_params": (c_int, [P, P, P, P, P]),
        "dnmf_measure_fp32_peak": (c_int, [c_int, c_int, POINTER(c_double)]),
        "dnmf_build_info": (c_int, []),
        "dnmf_debug_trip_assert": (c_int, [c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._signatures = sig
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dnmf_last_error()
        raise DnmfError("%s failed: %s" % (what or "dnmf call", msg.decode() if msg else "unknown error"))
