"""
Device-level fusion of element-wise chains (custos `Lazy` + `optimize()` analogue; north_star: "fuse chained ops from the tape where the
graph allows"): with `device.set_fusion(True)` the *MayGrad element-wise ops are recorded and executed as ONE sl_fused_chain launch
(forward) and ONE fused backward launch — results must equal the launch-per-op device exactly, which the rest of the suite pins
to the oracle.  Chains that do not fit the interpreter must silently run op by op.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def devices(cached=False):
    from sliced_b200.host import CUDA
    return CUDA(0, cached=cached), CUDA(0, cached=cached).set_fusion(True)


def test_chained_perf_graph_fused_vs_unfused():
    """examples/chained_perf.rs:86-114: out = x^2 * x + (b + x) * b, backward; 5 + 5 launches become 1 + 1"""
    plain, fused = devices()
    rng = np.random.default_rng(0)
    n = 100003
    xs, bs = rng.uniform(-2, 2, n).astype(np.float32), rng.uniform(-2, 2, n).astype(np.float32)
    res = []
    for dev in (plain, fused):
        x, b = dev.buffer(xs), dev.buffer(bs)
        l0 = dev.launches
        # registration order of chained_perf.rs:86-90; intermediates are dropped as soon as they are consumed
        mul_b = dev.mul(dev.add(b, x), b)
        out = dev.add(dev.mul(dev.square(x), x), mul_b)
        del mul_b
        o = out.read()
        l1 = dev.launches
        out.backward()
        res.append((o, x.grad().read(), b.grad().read(), l1 - l0, dev.launches - l1))
    (o0, xg0, bg0, f0, b0), (o1, xg1, bg1, f1, b1) = res
    assert np.array_equal(o0, o1) and np.array_equal(xg0, xg1) and np.array_equal(bg0, bg1)
    assert fused.fused_groups == 1 and fused.unfused_groups == 0
    assert f0 >= 5 and f1 <= 2, (f0, f1)
    assert b1 < b0, (b0, b1)
    x13, b21 = fused.buffer(np.full(4, 1.3, np.float32)), fused.buffer(np.full(4, 2.1, np.float32))
    out = fused.add(fused.mul(fused.square(x13), x13), fused.mul(fused.add(b21, x13), b21))
    assert np.all(out.read().view(np.uint32) == 0x41156459)   # chained_perf.rs:91
    plain.close(); fused.close()


@pytest.mark.parametrize("cached", [False, True])
def test_live_intermediates_no_grad_leaves_and_gradless_ops(cached):
    plain, fused = devices(cached)
    rng = np.random.default_rng(1)
    n = 4099
    xs, ys, zs = (rng.uniform(0.3, 1.7, n).astype(np.float32) for _ in range(3))
    res = []
    for dev in (plain, fused):
        x, y, z = dev.buffer(xs), dev.buffer(ys), dev.buffer(zs).no_grad()
        sq = dev.square(x)                       # a handle is kept: must be materialised, and its gradient must be complete
        # Clip registers no grad closure (ops.rs:419-425): nothing flows through it.  Two materialised results (sq, out), three leaves.
        out = dev.add(dev.sub(dev.mul(sq, y), z), dev.mul(dev.clip(dev.mul(sq, y), 0.5, 1.5), sq))
        o, s = out.read(), sq.read()
        out.backward()
        res.append((o, s, x.grad().read(), y.grad().read(), sq.grad().read()))
        assert dev.n_grads() >= 3
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    assert fused.fused_groups >= 1 and fused.unfused_groups == 0
    plain.close(); fused.close()


def test_chain_that_does_not_fit_runs_op_by_op():
    plain, fused = devices()
    rng = np.random.default_rng(2)
    n = 1000
    vs = [rng.uniform(-1, 1, n).astype(np.float32) for _ in range(5)]
    res = []
    for dev in (plain, fused):
        b = [dev.buffer(v) for v in vs]
        out = dev.add(dev.add(dev.mul(b[0], b[1]), dev.mul(b[2], b[3])), b[4])   # 5 leaves > 3
        out.backward()
        res.append([out.read()] + [x.grad().read() for x in b])
    for a, c in zip(*res):
        assert np.array_equal(a, c)
    assert fused.unfused_groups >= 1
    plain.close(); fused.close()


def test_mixed_with_other_ops_and_i32():
    """flush points: a gemm / reduction between element-wise ops, integer chains, aliasing (x * x)"""
    plain, fused = devices()
    res = []
    for dev in (plain, fused):
        lhs = dev.buffer(np.arange(1, 9, dtype=np.float32))
        rhs = dev.buffer(np.arange(1, 7, dtype=np.float32))
        h = dev.relu(dev.gemm(4, 2, 3, lhs, rhs))
        s = dev.sum_cols(3, dev.mul(h, h))
        out = dev.pow(s, 0.5)
        out.backward()
        xi = dev.buffer([1, 2, 3, 4, 5], np.int32)
        oi = dev.mul(dev.square(xi), xi)
        oi.backward()
        res.append((out.read(), lhs.grad().read(), rhs.grad().read(), oi.read(), xi.grad().read()))
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    assert res[1][3].tolist() == [1, 8, 27, 64, 125] and res[1][4].tolist() == [3, 12, 27, 48, 75]
    plain.close(); fused.close()


def test_sine_net_steps_with_fusion_are_bit_identical():
    """examples/sine_net.rs: (out - y)^2 and the relus go through the recorder; 30 steps must match the unfused device exactly"""
    from sliced_b200.host import CUDA, Mlp
    from tests import sine_replay as SR
    xs, ys, W, B = SR.problem()
    res = []
    for fusion in (False, True):
        dev = CUDA(0, cached=True)
        dev.set_fusion(fusion)
        mlp = Mlp(dev, SR.DIMS, 1)
        for l in range(3):
            mlp.weights(l).write(W[l])
        dx, dy = dev.buffer(xs).no_grad(), dev.buffer(ys).no_grad()
        hist = [mlp.step(dx, dy, None, 1000, 1e-4)[0] for _ in range(30)]
        res.append((hist, mlp.params().read(), dev.launches, dev.fused_groups))
        del mlp, dx, dy
        dev.close()
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1])
    assert res[1][3] > 0 and res[1][2] < res[0][2]
