"""Per-kernel CUDA-event times inside one examples/sine_net.rs training step (1-64-64-1, 1000 samples): where a latency-bound step
spends its time.  python tools/sine_net_breakdown.py [out.txt]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from sliced_b200 import capi
from sliced_b200.host import CUDA, Mlp

dims = [1, 64, 64, 1]
xs = (np.arange(1000) / 1000.0).astype(np.float32)
ys = np.sin(2.0 * xs * np.float32(np.pi)).astype(np.float32)
rng = np.random.default_rng(0)
dev = CUDA(0, cached=True)
if len(sys.argv) > 2 and sys.argv[2] == "fusion":
    dev.set_fusion(True)
mlp = Mlp(dev, dims, 1)
for l in range(3):
    mlp.weights(l).write(rng.uniform(-0.5, 0.5, dims[l] * dims[l + 1]).astype(np.float32))
dx, dy = dev.buffer(xs).no_grad(), dev.buffer(ys).no_grad()
for _ in range(5):
    mlp.step(dx, dy, None, 1000, 1e-4)
lib, ctx = capi.load(), dev.ctx_handle
buf = C.create_string_buffer(1 << 16)
capi.check(ctx, lib.sl_ctx_profile_report(ctx, None, 0))
capi.check(ctx, lib.sl_ctx_profile_begin(ctx))
steps = 20
for _ in range(steps):
    mlp.step(dx, dy, None, 1000, 1e-4, want_metrics=False)
capi.check(ctx, lib.sl_ctx_profile_report(ctx, buf, len(buf)))
capi.check(ctx, lib.sl_ctx_profile_end(ctx, None, None, None))
rows = [r.rsplit(",", 2) for r in buf.value.decode().strip().splitlines()]
tot = sum(float(r[2]) for r in rows) / steps
out = [f"# sine_net 1-64-64-1, 1000 samples: per-kernel CUDA-event time inside the step ({steps} steps averaged); sum {tot * 1e3:.1f} us/step",
       f"# {'kernel':64s} launches/step   us/step   us/launch"]
for name, n, ms in rows:
    out.append(f"{name[:66]:66s} {int(n) / steps:8.1f} {float(ms) / steps * 1e3:10.2f} {float(ms) / int(n) * 1e3:10.2f}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt + "\n")
del mlp, dx, dy
dev.close()
