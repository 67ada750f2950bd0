"""examples/sine_net.rs:119-166 replayed with numpy in a chosen precision (test infrastructure).

Used to pin WHY the loss after 1000 steps cannot be gated at SURVEY 8(d)'s rel 1e-4: the trajectory is chaotic (relu boundaries
flip), so two correct fp32 implementations that sum in a different order (the oracle's sequential loops vs numpy's BLAS) already
end 5.6e-2 apart at step 1000 while agreeing to 1e-7 over the first ten steps; the fp64 replay is the yardstick for all of them."""
import numpy as np

DIMS = [1, 64, 64, 1]


def problem():
    xs = (np.arange(1000) / 1000.0).astype(np.float32)
    ys = np.sin(2.0 * xs * np.float32(np.pi)).astype(np.float32)
    rng = np.random.default_rng(0)
    W = [rng.uniform(-0.5, 0.5, DIMS[i] * DIMS[i + 1]).astype(np.float32) for i in range(3)]
    B = [np.zeros(DIMS[i + 1], np.float32) for i in range(3)]
    return xs, ys, W, B


def replay(dtype, steps, W0, xs, ys, lr=1e-4):
    """loss sums of `steps` SGD steps in `dtype` (matrix products through numpy / BLAS)"""
    W = [w.astype(dtype).reshape(DIMS[i], DIMS[i + 1]) for i, w in enumerate(W0)]
    B = [np.zeros(DIMS[i + 1], dtype) for i in range(3)]
    x, y = xs.astype(dtype).reshape(-1, 1), ys.astype(dtype).reshape(-1, 1)
    lr = dtype(lr)
    losses = []
    for _ in range(steps):
        a, z = [x], []
        for l in range(3):
            zz = a[-1] @ W[l] + B[l]
            z.append(zz)
            a.append(((zz >= 0).astype(dtype) * zz) if l < 2 else zz)
        d = a[-1] - y
        losses.append(float(np.sum(d * d, dtype=dtype)))
        g = d * dtype(2)
        for l in (2, 1, 0):
            gW, gB = a[l].T @ g, g.sum(0, dtype=dtype)
            if l > 0:
                g = (g @ W[l].T) * (z[l - 1] >= 0).astype(dtype)
            W[l] = W[l] - gW * lr
            B[l] = B[l] - gB * lr
    return np.array(losses)
