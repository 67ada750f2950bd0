"""
sliced_b200 — B200 (sm_100a) implementation of the forward + backward op set of elftausend/sliced behind a C ABI.

    capi   ctypes binding of include/sliced_b200.h (libsliced_b200.so)
    raw    Context / DeviceArray: one method per L2 op trait of the reference
    build  in-tree nvcc build

There is no CPU fallback anywhere in this package: without the CUDA library or without a B200 every op raises.
"""
from .capi import (ADD, DIV, F32, F64, GEMM_3XF16, GEMM_3XTF32, GEMM_DEFAULT, GEMM_SIMT, GEMM_TF32, I32, MUL, SUB, SlicedError, UN_ADD_SCALAR,
                   UN_CLIP, UN_EXP, UN_LN, UN_MUL_SCALAR, UN_NEG, UN_NEG_DIV_SCALAR, UN_NEG_LN, UN_POW, UN_RELU, UN_SIGMOID,
                   UN_SQUARE, UN_TANH, device_count)
from .raw import Context, DeviceArray

__all__ = [n for n in dir() if not n.startswith("_")]
