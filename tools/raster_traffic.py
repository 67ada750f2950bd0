"""DRAM traffic of the 2-CTA gemm per tile-raster setting (run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum -k regex:2cta`): every (shape, group, along_n) runs the gemm twice; read the second launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sliced_b200 as S
from sliced_b200.raw import DeviceArray

torch.cuda.set_device(0)
ctx = S.Context(0, stream=torch.cuda.current_stream().cuda_stream)
L = ctx.lib


def buf(n):
    t = torch.empty(n, device="cuda").uniform_(-1, 1)
    return t, DeviceArray(ctx, n, np.float32, ptr=t.data_ptr(), owner=t)


CONFIGS = [tuple(c.split(":")) for c in os.environ.get("CONFIGS", "2:0,8:0,16:0,4:1,8:1,16:1").split(",")]
for name, ta, tb, m, n, k in [("fwd NN", 0, 0, 65536, 4096, 4096), ("dA NT", 0, 1, 65536, 4096, 4096), ("dW TN", 1, 0, 4096, 4096, 65536)]:
    ta_, a = buf(m * k); tb_, b = buf(k * n); tc_, c = buf(m * n)
    for grp, an in CONFIGS:
        os.environ["SLICED_GEMM_GROUP"], os.environ["SLICED_GEMM_GROUP_N"] = grp, an
        for _ in range(2):
            assert L.sl_gemm_ex(ctx.h, S.F32, ta, tb, m, n, k, a.ptr, b.ptr, c.ptr, 0, S.GEMM_3XF16) == 0
        torch.cuda.synchronize()
        print(f"{name} group={grp} along_n={an}", flush=True)
    del ta_, tb_, tc_, a, b, c
    torch.cuda.empty_cache()
