"""MMA-kernel-only timing (sl_ctx_profile events) of the nn.rs MLP gemm shapes, per tile config / chunk length."""
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


shapes = [("fwd  NN", 0, 0, 65536, 4096, 4096), ("dA   NT", 0, 1, 65536, 4096, 4096), ("dW   TN", 1, 0, 4096, 4096, 65536),
          ("sq   NT", 0, 1, 8192, 8192, 8192)]
B = int(os.environ.get("BATCH", "65536"))
MODE = {"3xf16": S.GEMM_3XF16, "3xtf32": S.GEMM_3XTF32, "tf32": S.GEMM_TF32}[os.environ.get("MODE", "3xf16")]
for name, ta, tb, m, n, k in shapes:
    if name.startswith(("fwd", "dA")):
        m = B
    if name.startswith("dW"):
        k = B
    ta_, a = buf(m * k)
    tb_, b = buf(k * n)
    tc_, c = buf(m * n)
    for cfg in os.environ.get("CFGS", "1,4").split(","):
        for kc in os.environ.get("KCS", "4").split(","):
            os.environ["SLICED_GEMM_CFG"] = cfg
            os.environ["SLICED_GEMM_KC"] = kc
            fn = lambda: L.sl_gemm_ex(ctx.h, S.F32, ta, tb, m, n, k, a.ptr, b.ptr, c.ptr, 0, MODE)
            for _ in range(2):
                assert fn() == 0
            capi.check(ctx.h, L.sl_ctx_profile_begin(ctx.h))
            reps = 5
            for _ in range(reps):
                fn()
            nl, ms, fl = C.c_uint64(0), C.c_double(0), C.c_double(0)
            capi.check(ctx.h, L.sl_ctx_profile_end(ctx.h, C.byref(nl), C.byref(ms), C.byref(fl)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            tot = e0.elapsed_time(e1) / reps
            eff = fl.value / (ms.value * 1e-3) / 1e12
            print(f"{name} {m}x{n}x{k} cfg={cfg} kc={kc}: mma {ms.value / reps:7.3f} ms  {eff:6.1f} TF/s eff ({3 * eff:6.1f} issued)   "
                  f"call incl. prep {tot:7.3f} ms  prep {tot - ms.value / reps:6.3f} ms", flush=True)
    del ta_, tb_, tc_, a, b, c
    torch.cuda.empty_cache()
