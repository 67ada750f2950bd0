"""tile-raster sweep for the 2-CTA gemm on the MLP shapes: SLICED_GEMM_GROUP x SLICED_GEMM_GROUP_N -> MMA-kernel time"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sliced_b200 as S
from sliced_b200 import capi
from sliced_b200.raw import DeviceArray

torch.cuda.set_device(0)
ctx = S.Context(0, stream=torch.cuda.current_stream().cuda_stream)
L = ctx.lib


def buf(n):
    t = torch.empty(n, device="cuda").uniform_(-1, 1)
    return t, DeviceArray(ctx, n, np.float32, ptr=t.data_ptr(), owner=t)


for name, ta, tb, m, n, k in [("fwd NN", 0, 0, 65536, 4096, 4096), ("dA NT", 0, 1, 65536, 4096, 4096), ("dW TN", 1, 0, 4096, 4096, 65536)]:
    ta_, a = buf(m * k); tb_, b = buf(k * n); tc_, c = buf(m * n)
    for along_n in ("0", "1"):
        for grp in ("2", "4", "8", "16"):
            os.environ["SLICED_GEMM_GROUP"] = grp
            os.environ["SLICED_GEMM_GROUP_N"] = along_n
            fn = lambda: L.sl_gemm_ex(ctx.h, S.F32, ta, tb, m, n, k, a.ptr, b.ptr, c.ptr, 0, S.GEMM_3XTF32)
            assert fn() == 0
            capi.check(ctx.h, L.sl_ctx_profile_begin(ctx.h))
            for _ in range(4):
                fn()
            nl, ms, fl = C.c_uint64(0), C.c_double(0), C.c_double(0)
            capi.check(ctx.h, L.sl_ctx_profile_end(ctx.h, C.byref(nl), C.byref(ms), C.byref(fl)))
            print(f"{name} group={grp:2s} along_n={along_n}: mma {ms.value / 4:7.3f} ms  {fl.value / (ms.value * 1e-3) / 1e12:6.1f} TF/s", flush=True)
    del ta_, tb_, tc_, a, b, c
    torch.cuda.empty_cache()
