"""ctypes binding of libpyrhe_b200.so (the C ABI declared in include/pyrhe_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every
entry point turns a negative return code into a Python exception carrying
`rhe_last_error()`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpyrhe_b200.so")

PATH_SIMT = 0
PATH_TCGEN05 = 1


class RheConfig(C.Structure):
    _fields_ = [(name, C.c_int32) for name in (
        "device", "n_indv", "n_kept", "pitch_bytes", "n_cols_set", "n_sets", "n_ops", "n_vec",
        "n_bins", "max_block_snps", "impute_binary", "kernel_path")]


class RheError(RuntimeError):
    pass


# name -> (restype, argtypes); must list every symbol include/pyrhe_b200.h declares
SIGNATURES = {
    "rhe_version": (C.c_int, []),
    "rhe_last_error": (C.c_char_p, []),
    "rhe_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(RheConfig)]),
    "rhe_ctx_destroy": (C.c_int, [C.c_void_p]),
    "rhe_set_rhs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rhe_set_uniforms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "rhe_upload_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "rhe_block_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "rhe_decode_block": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "rhe_block_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32),
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rhe_loo_gram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "rhe_loo_gram_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    "rhe_launch_count": (C.c_int64, [C.c_void_p]),
    "rhe_synth_genotypes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_uint64, C.c_float,
                                      C.c_void_p]),
    "rhe_timing_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "rhe_timing_collect": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
}

_lib = None


def load():
    """Load the library (once).  Raises RheError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RheError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(pyrhe_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise RheError(f"pyrhe_b200 error {rc}: {load().rhe_last_error().decode()}")


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, or NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def tcgen05_supported(plan, limbs: int = None) -> bool:
    """Mirror of the limits rhe_tc_create enforces (TMEM columns per CTA, staged metadata sizes)."""
    L = int(os.environ.get("PYRHE_B200_LIMBS", "3")) if limbs is None else limbs
    r1p = -(-(plan.n_sets * plan.Rs) // 4) * 4
    nba = -(-(L * r1p) // 16) * 16
    bp = -(-plan.B // 2) * 2
    ncb = -(-(plan.n_groups * L * bp) // 16) * 16
    if plan.K * ncb <= 512:
        kg = plan.K                                   # all bins in one pass-B launch
    else:
        kg = 512 // (2 * ncb) if 2 * ncb <= 512 else (512 // ncb if ncb <= 512 else 0)   # bin groups
    return (nba <= 256 and 1 <= kg <= 255 and plan.n_groups * kg * plan.B <= 512 and plan.n_groups * plan.B <= 64)
