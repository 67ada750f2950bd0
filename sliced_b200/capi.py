"""
ctypes binding of the C ABI declared in include/sliced_b200.h (libsliced_b200.so).

This module is the thinnest possible layer: one Python function per exported symbol, status codes turned into
exceptions.  It contains no arithmetic and no fallback: if the CUDA library is missing or no B200 is visible, calls
fail loudly (SlicedError).
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsliced_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "sliced_b200.h")

SL_OK, SL_ERR_INVALID_ARG, SL_ERR_CUDA, SL_ERR_UNSUPPORTED, SL_ERR_NCCL, SL_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5
F32, F64, I32 = 0, 1, 2
ADD, SUB, MUL, DIV = 0, 1, 2, 3
(UN_SQUARE, UN_POW, UN_RELU, UN_TANH, UN_SIGMOID, UN_EXP, UN_LN, UN_NEG_LN, UN_CLIP, UN_NEG, UN_MUL_SCALAR,
 UN_NEG_DIV_SCALAR, UN_ADD_SCALAR) = range(13)
GEMM_DEFAULT, GEMM_3XTF32, GEMM_TF32, GEMM_SIMT, GEMM_3XF16 = -1, 0, 1, 2, 3


class SlicedError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sliced_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load(build_if_missing: bool = False) -> C.CDLL:
    """Load libsliced_b200.so.  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _b
            _b.build()
        else:
            raise SlicedError(SL_ERR_NO_DEVICE, f"{LIB_PATH} is missing: run `python -m sliced_b200.build` (nvcc, sm_100a)")
    _lib = C.CDLL(LIB_PATH)
    _declare(_lib)
    return _lib


def declared_symbols() -> list[str]:
    """Every function name declared in include/sliced_b200.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sl_[a-z0-9_]+)\s*\(", txt)))


_vp, _sz, _i, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_double


def _declare(lib):
    P = C.POINTER
    sig = {
        "sl_abi_version": ([], _i),
        "sl_device_count": ([], _i),
        "sl_ctx_create": ([_i, P(_vp)], _i),
        "sl_ctx_create_on_stream": ([_i, _vp, P(_vp)], _i),
        "sl_ctx_destroy": ([_vp], _i),
        "sl_last_error_string": ([_vp], C.c_char_p),
        "sl_ctx_stream": ([_vp], _vp),
        "sl_ctx_device": ([_vp], _i),
        "sl_ctx_set_gemm_mode": ([_vp, _i], _i),
        "sl_ctx_launch_count": ([_vp], C.c_uint64),
        "sl_ctx_profile_begin": ([_vp], _i),
        "sl_ctx_profile_end": ([_vp, P(C.c_uint64), P(_d), P(_d)], _i),
        "sl_ctx_profile_report": ([_vp, C.c_char_p, _sz], _i),
        "sl_malloc": ([_vp, _sz, P(_vp)], _i),
        "sl_free": ([_vp, _vp], _i),
        "sl_host_alloc": ([_vp, _sz, P(_vp)], _i),
        "sl_host_free": ([_vp, _vp], _i),
        "sl_write": ([_vp, _vp, _vp, _sz], _i),
        "sl_read": ([_vp, _vp, _vp, _sz], _i),
        "sl_read_async": ([_vp, _vp, _vp, _sz], _i),
        "sl_write_prefetch": ([_vp, _vp, _vp, _sz], _i),
        "sl_prefetch_wait": ([_vp], _i),
        "sl_prefetch_release": ([_vp], _i),
        "sl_copy": ([_vp, _vp, _vp, _sz], _i),
        "sl_clear": ([_vp, _vp, _sz], _i),
        "sl_fill": ([_vp, _i, _vp, _d, _sz], _i),
        "sl_sync": ([_vp], _i),
        "sl_graph_begin": ([_vp], _i),
        "sl_graph_end": ([_vp, P(_vp)], _i),
        "sl_graph_launch": ([_vp, _vp], _i),
        "sl_graph_destroy": ([_vp, _vp], _i),
        "sl_binary_ew": ([_vp, _i, _i, _vp, _vp, _vp, _sz], _i),
        "sl_binary_ew_grad": ([_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz], _i),
        "sl_add_ew_grad": ([_vp, _i, _vp, _vp, _vp, _sz], _i),
        "sl_unary": ([_vp, _i, _i, _d, _d, _vp, _vp, _sz], _i),
        "sl_unary_grad": ([_vp, _i, _i, _d, _d, _vp, _vp, _vp, _sz], _i),
        "sl_row_op": ([_vp, _i, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_add_row": ([_vp, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_add_row_mut": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_add_row_grad": ([_vp, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_add_row_mut_grad": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_row_op_grad": ([_vp, _i, _i, _sz, _sz, _vp, _vp, _vp, _vp, _vp], _i),
        "sl_col_op": ([_vp, _i, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_col_op_grad": ([_vp, _i, _i, _sz, _sz, _vp, _vp, _vp, _vp, _vp], _i),
        "sl_sgd_step": ([_vp, _i, _vp, _vp, _d, _sz], _i),
        "sl_fused_chain": ([_vp, _i, _vp, _vp, _vp, _sz], _i),   # (typed prototype: sliced_b200/chain.py)
        "sl_chained_fwd": ([_vp, _i, _vp, _vp, _vp, _sz], _i),
        "sl_chained_bwd": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _sz], _i),
        "sl_gemm": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _i], _i),
        "sl_gemm_nt": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _i], _i),
        "sl_gemm_tn": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _i], _i),
        "sl_gemm_ex": ([_vp, _i, _i, _i, _sz, _sz, _sz, _vp, _vp, _vp, _i, _i], _i),
        "sl_gemm_grad": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _i, _i], _i),
        "sl_linear_bwd_params": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _i], _i),
        "sl_linear_bwd_params_exchange": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _i, _i], _i),
        "sl_gemm_scope_begin": ([_vp], _i),
        "sl_gemm_scope_end": ([_vp], _i),
        "sl_linear_fwd": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _i], _i),
        "sl_linear_bwd_input_relu": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _i], _i),
        "sl_linear_fwd_bits": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _i], _i),
        "sl_linear_bwd_input_relu_bits": ([_vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _i], _i),
        "sl_sum": ([_vp, _i, _vp, _sz, _vp], _i),
        "sl_mean": ([_vp, _i, _vp, _sz, _vp], _i),
        "sl_max": ([_vp, _i, _vp, _sz, _vp], _i),
        "sl_sum_rows": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_sum_cols": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_mean_rows": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_mean_cols": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_max_rows": ([_vp, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_max_cols": ([_vp, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_sum_rows_grad": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_sum_cols_grad": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_mean_rows_grad": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_mean_cols_grad": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_max_rows_grad": ([_vp, _i, _sz, _sz, _vp, _vp, _vp, _vp], _i),
        "sl_max_cols_grad": ([_vp, _i, _sz, _sz, _vp, _vp, _vp, _vp], _i),
        "sl_max_cols_grad_idx": ([_vp, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_transpose": ([_vp, _i, _sz, _sz, _vp, _vp, _i], _i),
        "sl_softmax": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_softmax_grad": ([_vp, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_softmax_cce": ([_vp, _i, _sz, _sz, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp], _i),
        "sl_mlp_small_fits": ([_vp, _i, _vp, _sz], _i),
        "sl_mlp_small_step": ([_vp, _i, _i, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _d, _vp], _i),
        "sl_diagflat": ([_vp, _i, _sz, _vp, _vp], _i),
        "sl_diagflat_grad": ([_vp, _i, _sz, _vp, _vp], _i),
        "sl_onehot": ([_vp, _i, _sz, _sz, _vp, _vp], _i),
        "sl_onehot_grad": ([_vp, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_count_correct": ([_vp, _i, _sz, _sz, _vp, _vp, _vp], _i),
        "sl_comm_unique_id": ([_vp], _i),
        "sl_comm_init_rank": ([_vp, _i, _i, _vp], _i),
        "sl_allreduce_sum": ([_vp, _i, _vp, _sz], _i),
        "sl_allreduce_sum_async": ([_vp, _i, _vp, _sz], _i),
        "sl_comm_wait": ([_vp], _i),
        "sl_comm_wait_n": ([_vp, _i], _i),
        "sl_comm_issued": ([_vp], _i),
        "sl_comm_nranks": ([_vp], _i),
        "sl_comm_destroy": ([_vp], _i),
    }
    for name, (args, res) in sig.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = res
    lib._sl_signatures = sig


def check(ctx_handle, rc):
    if rc != SL_OK:
        msg = load().sl_last_error_string(ctx_handle)
        raise SlicedError(rc, msg.decode() if msg else "?")


def device_count() -> int:
    return load().sl_device_count()
