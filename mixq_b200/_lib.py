"""ctypes loader for libmixq_sm100.so (the C ABI declared in include/mixq.h).

There is no fallback: if the shared library is missing the import raises, and every call that
returns non-zero raises `MixqError` with the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libmixq_sm100.so"


class MixqError(RuntimeError):
    pass


class LinearArgs(C.Structure):
    """Mirror of `mixq_linear_args` (include/mixq.h)."""

    _fields_ = [
        ("x", C.c_void_p),
        ("norm_weight", C.c_void_p),
        ("norm_out", C.c_void_p),
        ("eps", C.c_float),
        ("M", C.c_int),
        ("N", C.c_int),
        ("K", C.c_int),
        ("q_weight", C.c_void_p),
        ("scale_col", C.c_void_p),
        ("bias", C.c_void_p),
        ("bit", C.c_int),
        ("q_weight_up", C.c_void_p),
        ("scale_col_up", C.c_void_p),
        ("weight_cache_up", C.c_void_p),
        ("ind", C.c_void_p),
        ("n_ind", C.c_int),
        ("weight_cache", C.c_void_p),
        ("ld_wc", C.c_int),
        ("q_x", C.c_void_p),
        ("x_scale", C.c_void_p),
        ("act_outliers", C.c_void_p),
        ("ld_ao", C.c_int),
        ("sigma", C.c_float),
        ("col_over", C.c_void_p),
        ("over_flag", C.c_void_p),
        ("residual", C.c_void_p),
        ("ld_res", C.c_int),
        ("y", C.c_void_p),
        ("act", C.c_int),
        ("skip_prologue", C.c_int),
        ("grid_sync", C.c_void_p),
        ("tile_n", C.c_int),
        ("y_peer", C.c_void_p * 8),
        ("peer_cols", C.c_int),
        ("peer_bcast", C.c_int),
        ("splitk_ws", C.c_void_p),
        ("splitk_ws_bytes", C.c_longlong),
    ]


class AllReduceArgs(C.Structure):
    """Mirror of `mixq_allreduce_args` (include/mixq.h)."""

    _fields_ = [
        ("partial0", C.c_void_p * 8),
        ("partial1", C.c_void_p * 8),
        ("flags", C.c_void_p * 8),
        ("result0", C.c_void_p * 8),
        ("result1", C.c_void_p * 8),
        ("epoch", C.c_void_p),
        ("done", C.c_void_p),
        ("residual", C.c_void_p),
        ("out", C.c_void_p),
        ("n", C.c_longlong),
        ("world", C.c_int),
        ("rank", C.c_int),
        ("buf", C.c_int),
    ]


class McAllReduceArgs(C.Structure):
    """Mirror of `mixq_mc_allreduce_args` (include/mixq.h)."""

    _fields_ = [
        ("mc", C.c_void_p),
        ("local", C.c_void_p),
        ("partial_off", C.c_ulonglong * 2),
        ("result_off", C.c_ulonglong * 2),
        ("flags_off", C.c_ulonglong),
        ("epoch", C.c_void_p),
        ("done", C.c_void_p),
        ("residual", C.c_void_p),
        ("n", C.c_longlong),
        ("world", C.c_int),
        ("rank", C.c_int),
        ("buf", C.c_int),
    ]


class ExchangeFinishArgs(C.Structure):
    """Mirror of `mixq_exchange_finish_args` (include/mixq.h)."""

    _fields_ = [
        ("recv", C.c_void_p),
        ("result", C.c_void_p * 8),
        ("mc_result", C.c_void_p),
        ("flags", C.c_void_p * 8),
        ("mc_flags", C.c_void_p),
        ("epoch", C.c_void_p),
        ("done", C.c_void_p),
        ("residual", C.c_void_p),
        ("M", C.c_int),
        ("N", C.c_int),
        ("world", C.c_int),
        ("rank", C.c_int),
        ("one_shot", C.c_int),
    ]


class ExchangePollArgs(C.Structure):
    """Mirror of `mixq_exchange_poll_args` (include/mixq.h)."""

    _fields_ = [
        ("recv", C.c_void_p),
        ("result", C.c_void_p * 8),
        ("mc_result", C.c_void_p),
        ("reset", C.c_void_p),
        ("residual", C.c_void_p),
        ("M", C.c_int),
        ("N", C.c_int),
        ("world", C.c_int),
        ("rank", C.c_int),
        ("one_shot", C.c_int),
    ]


class LinearPlan(C.Structure):
    """Mirror of `mixq_linear_plan` (include/mixq.h)."""

    _fields_ = [(n, C.c_int) for n in ("two_cta", "tile_w", "k_atoms", "stage_bytes", "nstages", "tiles", "units",
                                         "tiles_per_unit", "acc_slots", "passes", "pass_cols", "pass_buffers", "tmem_cols")]


_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "mixq_find_row_scale": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "mixq_find_row_scale_scan": [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp],
    "mixq_extract_outliers_and_set_to_zeros": [_vp, _i, _vp, _vp, _i, _i, _i, _vp],
    "mixq_int8_fused_dequantize": [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "mixq_int4_fused_dequantize": [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "mixq_gemm_i8": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "mixq_dequantize_int8": [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp],
    "mixq_unpack_int4_to_fp16": [_vp, _vp, _i, _vp, _i, _i, _i, _vp],
    "mixq_rmsnorm": [_vp, _vp, _vp, _f, _i, _i, _vp],
    "mixq_rmsnorm_extract_outliers": [_vp, _vp, _vp, _f, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp],
    "mixq_gather_weight_columns": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "mixq_compact_outlier_columns": [_vp, _i, _vp, _i, _vp, _vp],
    "mixq_linear_fused": [C.POINTER(LinearArgs), _vp],
    "mixq_rope_attention_decode": [_vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _f, _vp],
    "mixq_rope_attention_decode_quant": [_vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _f, _vp, _i, _vp, _i, _vp, _vp, _i, _vp],
    "mixq_quik_quantize": [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp],
    "mixq_quik_addend": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp],
    "mixq_reload_debug_env": [],
    "mixq_debug_pingpong": [_vp, _vp, _vp, _i, _i, _vp, _vp],
    "mixq_mul_inplace": [_vp, _vp, _ll, _vp],
    "mixq_peer_alloc": [C.c_ulonglong, C.POINTER(C.c_void_p)],
    "mixq_peer_free": [_vp],
    "mixq_ipc_get_handle": [_vp, C.c_char_p],
    "mixq_ipc_open_handle": [C.c_char_p, C.POINTER(C.c_void_p)],
    "mixq_ipc_close_handle": [_vp],
    "mixq_allreduce_residual": [C.POINTER(AllReduceArgs), _vp],
    "mixq_allreduce_multicast": [C.POINTER(McAllReduceArgs), _vp],
    "mixq_exchange_finish": [C.POINTER(ExchangeFinishArgs), _vp],
    "mixq_exchange_finish_poll": [C.POINTER(ExchangePollArgs), _vp],
    "mixq_exchange_finish_poll_quant": [C.POINTER(ExchangePollArgs), _vp, _f, _vp, _i, _vp, _i, _vp, _vp, _i, _vp],
    "mixq_set_peer_timeout_ms": [_ll],
    "mixq_set_tile_n": [_i],
    "mixq_set_pdl": [_i],
    "mixq_set_grid_barrier_mode": [_i],
    "mixq_plan_linear": [_i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(LinearPlan)],
    "mixq_plan_split_k": [_i, _i, _i, _i, _i, _i, _ll],
    "mixq_set_trace_buffer": [_vp],
    "mixq_version": [],
    "mixq_launch_count": [],
    "mixq_last_error": [],
}
_RESTYPES = {"mixq_launch_count": C.c_ulonglong, "mixq_last_error": C.c_char_p, "mixq_reload_debug_env": None}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises MixqError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("MIXQ_LIB", LIB_PATH))
    if not path.exists():
        raise MixqError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mixq_b200/csrc` — there is no CPU or PyTorch fallback for the MixLinear path"
        )
    lib = C.CDLL(str(path))
    lenient = os.environ.get("MIXQ_LIB_LENIENT") == "1"     # A/B-testing an older build of the library (tools only)
    for name, argtypes in SIGNATURES.items():
        if lenient and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)  # AttributeError here == header/library drift
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mixq_last_error()
        raise MixqError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().mixq_launch_count())
