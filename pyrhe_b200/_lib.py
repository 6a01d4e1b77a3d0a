"""ctypes binding of libpyrhe_b200.so (the C ABI declared in include/pyrhe_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every
entry point turns a negative return code into a Python exception carrying
`rhe_last_error()`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PYRHE_B200_LIB: another build of the same library (profiling / ablation builds made with PYRHE_B200_EXTRA_NVCC)
LIB_PATH = os.environ.get("PYRHE_B200_LIB") or os.path.join(_HERE, "csrc", "libpyrhe_b200.so")

PATH_SIMT = 0
PATH_TCGEN05 = 1
ROWS_PLINK = 0
ROWS_TILED = 1


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
    "rhe_tc_supported": (C.c_int, [C.POINTER(RheConfig)]),
    "rhe_block_plan_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p,
                                        C.POINTER(C.c_void_p)]),
    "rhe_block_plan_destroy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rhe_block_fast_bytes": (C.c_int64, [C.c_void_p, C.c_void_p]),
    "rhe_block_transpose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rhe_block_tiled_bytes": (C.c_int64, [C.c_void_p, C.c_void_p]),
    "rhe_block_retile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rhe_block_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "rhe_sum_partials": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
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


def plan_config(plan, **fields) -> RheConfig:
    """rhe_config of a PathPlan (assemble.py); `fields` fill in the sizes the plan does not know."""
    base = dict(device=0, n_indv=4, n_kept=4, pitch_bytes=128, n_cols_set=plan.Rs, n_sets=plan.n_sets, n_ops=plan.n_ops,
                n_vec=plan.B, n_bins=plan.K, max_block_snps=1, impute_binary=1, kernel_path=PATH_TCGEN05)
    base.update(fields)
    return RheConfig(**base)


def tcgen05_supported(plan) -> bool:
    """Whether the layout of `plan` fits the tcgen05 kernels: asked of the library itself (rhe_tc_supported), which
    applies exactly the checks rhe_ctx_create enforces (TMEM columns, shared-memory budget, M-tile selection)."""
    cfg = plan_config(plan)
    return bool(load().rhe_tc_supported(C.byref(cfg)))


def tcgen05_unsupported_reason(plan) -> str:
    cfg = plan_config(plan)
    lib = load()
    return "" if lib.rhe_tc_supported(C.byref(cfg)) else lib.rhe_last_error().decode()
