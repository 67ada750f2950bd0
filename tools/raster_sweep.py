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
REPS = int(os.environ.get('REPS', '12'))


def buf(n):
    t = torch.empty(n, device="cuda").uniform_(-1, 1)
    return t, DeviceArray(ctx, n, np.float32, ptr=t.data_ptr(), owner=t)


ROUNDS = int(os.environ.get("ROUNDS", "5"))
CONFIGS = [(g, an) for an in ("0", "1") for g in os.environ.get("GROUPS", "1,2,4,8").split(",")]
for name, ta, tb, m, n, k in [("fwd NN", 0, 0, 65536, 4096, 4096), ("dA NT", 0, 1, 65536, 4096, 4096), ("dW TN", 1, 0, 4096, 4096, 65536)]:
    ta_, a = buf(m * k); tb_, b = buf(k * n); tc_, c = buf(m * n)
    acc = {cfg: 0.0 for cfg in CONFIGS}
    fn = lambda: L.sl_gemm_ex(ctx.h, S.F32, ta, tb, m, n, k, a.ptr, b.ptr, c.ptr, 0, S.GEMM_3XF16)
    for rnd in range(ROUNDS):   # configurations interleaved round-robin so that clock / thermal drift hits all of them alike
        for cfg in CONFIGS:
            os.environ["SLICED_GEMM_GROUP"], os.environ["SLICED_GEMM_GROUP_N"] = cfg
            assert fn() == 0
            capi.check(ctx.h, L.sl_ctx_profile_begin(ctx.h))
            for _ in range(REPS):
                fn()
            nl, ms, fl = C.c_uint64(0), C.c_double(0), C.c_double(0)
            capi.check(ctx.h, L.sl_ctx_profile_end(ctx.h, C.byref(nl), C.byref(ms), C.byref(fl)))
            acc[cfg] += ms.value / REPS
    flops = 2.0 * m * n * k
    for cfg in CONFIGS:
        t = acc[cfg] / ROUNDS
        print(f"{name} group={cfg[0]:2s} along_n={cfg[1]}: mma {t:7.3f} ms  {flops / (t * 1e-3) / 1e12:6.1f} TF/s", flush=True)
    del ta_, tb_, tc_, a, b, c
    torch.cuda.empty_cache()
