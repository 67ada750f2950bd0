"""
BASELINE.json configs[0]: examples/sine_net.rs at the shipped sizes (x, y: 1000 x 1; Linear 1-64-64-1; lr 1e-4; 1000 iterations).
Launch-latency-bound on a GPU: reports steps/s for the eager tape, for CUDA-graph replay (custos `Lazy` + `run()`), and for the
CPU oracle (the reference's own path) on the host, plus the final-loss parity between them.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle as O
from sliced_b200.host import CUDA, Mlp

dims = [1, 64, 64, 1]
xs = (np.arange(1000) / 1000.0).astype(np.float32)
ys = np.sin(2.0 * xs * np.float32(np.pi)).astype(np.float32)
rng = np.random.default_rng(0)
W0 = [rng.uniform(-0.5, 0.5, dims[i] * dims[i + 1]).astype(np.float32) for i in range(3)]
ITERS = 1000

out = {}
for mode in ("eager", "graph"):
    dev = CUDA(0, cached=True)
    mlp = Mlp(dev, dims, 1)
    for l in range(3):
        mlp.weights(l).write(W0[l])
    dx, dy = dev.buffer(xs).no_grad(), dev.buffer(ys).no_grad()
    fn = mlp.step_replay if mode == "graph" else mlp.step
    fn(dx, dy, None, 1000, 1e-4)
    fn(dx, dy, None, 1000, 1e-4)
    dev.sync()
    l0 = dev.launches
    t0 = time.perf_counter()
    for _ in range(ITERS - 2):
        loss = fn(dx, dy, None, 1000, 1e-4, want_metrics=False)
    loss = fn(dx, dy, None, 1000, 1e-4)[0] if False else None
    dev.sync()
    dt = time.perf_counter() - t0
    final = mlp.step(dx, dy, None, 1000, 1e-4)[0] / 1000
    out[mode] = ((ITERS - 2) / dt, (dev.launches - l0) / (ITERS - 2), final)
    del mlp, dx, dy
    dev.close()

W = [w.copy() for w in W0]
B = [np.zeros(dims[i + 1], np.float32) for i in range(3)]
t0 = time.perf_counter()
for _ in range(ITERS):
    l = O.mlp_step(1, dims, xs, ys, None, W, B, 1e-4)[0]
dt = time.perf_counter() - t0
final_cpu = O.mlp_step(1, dims, xs, ys, None, W, B, 1e-4)[0] / 1000
print(f"sine_net 1-64-64-1, 1000 samples, {ITERS} iterations")
for mode, (sps, lps, fl) in out.items():
    print(f"  gpu {mode:6s}: {sps:9.0f} steps/s  {1e6 / sps:7.1f} us/step  {lps:6.1f} launches/step  loss after {ITERS + 1} steps {fl:.6f}")
print(f"  cpu oracle: {ITERS / dt:9.0f} steps/s  {1e6 * dt / ITERS:7.1f} us/step  (1 thread)               loss after {ITERS + 1} steps {final_cpu:.6f}")
