"""Pins the CPU oracle against every known-answer test the reference holds for the path."""
import numpy as np
import pytest

from tests.golden.reference_kats import KATS
from tests.kat_runner import run_kat
from tests.oracle_backend import OracleBackend

import oracle as O


@pytest.mark.parametrize("kat", KATS, ids=[k["id"] for k in KATS])
def test_oracle_kat(kat):
    run_kat(OracleBackend(), kat)


# ---- tape-level known answers: the op, a second op for a non-trivial upstream gradient, backward() seeded with ones


def test_tape_max_cols_add():
    """tests/test_max_cols.rs:7-34"""
    lhs = O.arr([-3, 2, 3, 1, 1, 5, -5, 4, -9, -2, -4, -1], np.int32)
    rhs = O.arr([1, 4, 2], np.int32)
    mc = O.max_cols(4, lhs)
    assert mc.tolist() == [3, 5, -1]
    out = O.binary_ew(O.ADD, mc, rhs)
    g_out = np.ones_like(out)
    g_mc, g_rhs, g_lhs = np.zeros_like(mc), np.zeros_like(rhs), np.zeros_like(lhs)
    O.binary_ew_grad(O.ADD, mc, rhs, g_mc, g_rhs, g_out)
    O.max_cols_grad(4, mc, lhs, g_lhs, g_mc)
    assert g_lhs.tolist() == [0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1]
    assert g_rhs.tolist() == [1, 1, 1]


def test_tape_max_rows_mul():
    """tests/test_max_rows.rs:7-33"""
    lhs = O.arr([-3, 2, 3, 1, 1, 5, -5, 4, -9, -2, -4, -1], np.int32)
    rhs = O.arr([2, 3, 4, 1], np.int32)
    mr = O.max_rows(4, lhs)
    out = O.binary_ew(O.MUL, mr, rhs)
    g_mr, g_rhs, g_lhs = np.zeros_like(mr), np.zeros_like(rhs), np.zeros_like(lhs)
    O.binary_ew_grad(O.MUL, mr, rhs, g_mr, g_rhs, np.ones_like(out))
    O.max_rows_grad(4, mr, lhs, g_lhs, g_mr)
    assert g_lhs.tolist() == [0, 0, 4, 0, 2, 3, 0, 1, 0, 0, 0, 0]
    assert g_rhs.tolist() == [1, 5, 3, 4]


def test_tape_sum_cols_mul():
    """tests/test_sum_cols.rs:7-31"""
    x = O.arr([4, 2, 1, 3, 6, 1, 3, 1, 5, 4, 1, 1], np.int32)
    rhs = O.arr([1, 4, 2], np.int32)
    s = O.sum_cols(4, x)
    out = O.binary_ew(O.MUL, s, rhs)
    g_s, g_rhs, g_x = np.zeros_like(s), np.zeros_like(rhs), np.zeros_like(x)
    O.binary_ew_grad(O.MUL, s, rhs, g_s, g_rhs, np.ones_like(out))
    O.sum_cols_grad(4, g_x, g_s)
    assert g_x.tolist() == [1] * 4 + [4] * 4 + [2] * 4
    assert g_rhs.tolist() == [10, 11, 11]


def test_tape_sum_rows_mul():
    """tests/test_sum_rows.rs:6-30 (the reference panics at src/ops.rs:556; its vectors are still the spec)"""
    x = O.arr([4, 2, 1, 3, 6, 1, 3, 1, 5, 4, 1, 1], np.int32)
    rhs = O.arr([1, 4, 2, 3], np.int32)
    s = O.sum_rows(4, x)
    out = O.binary_ew(O.MUL, s, rhs)
    g_s, g_rhs, g_x = np.zeros_like(s), np.zeros_like(rhs), np.zeros_like(x)
    O.binary_ew_grad(O.MUL, s, rhs, g_s, g_rhs, np.ones_like(out))
    O.sum_rows_grad(4, g_x, g_s)
    assert g_x.tolist() == [1, 4, 2, 3] * 3
    assert g_rhs.tolist() == [15, 7, 5, 5]


def test_tape_mean_rows_sub():
    """tests/test_mean.rs:8-32"""
    lhs = O.arr([2., 3., 4.], np.float64)
    x = O.arr([1., 4., 3., 2., 3., 5., 2., 1., 6., -2., 1., 4.], np.float64)
    mr = O.mean_rows(3, x)
    out = O.binary_ew(O.SUB, lhs, mr)
    g_lhs, g_mr, g_x = np.zeros_like(lhs), np.zeros_like(mr), np.zeros_like(x)
    O.binary_ew_grad(O.SUB, lhs, mr, g_lhs, g_mr, np.ones_like(out))
    O.mean_rows_grad(3, g_x, g_mr)
    assert g_x.tolist() == [-0.25] * 12
    assert g_lhs.tolist() == [1., 1., 1.]


def test_tape_transpose_mul():
    """tests/matrix/transpose.rs:9-22: x.T * y ; x.grad == transpose(y). transpose_grad gets swapped dims (src/ops.rs:213-214)."""
    x = O.arr([1., 2., 3., 4., 5., 6.], np.float64)
    y = O.arr([-1., 1., 2., -5., -3., 2.], np.float64)
    xt = O.transpose(2, 3, x)
    out = O.binary_ew(O.MUL, y, xt)
    g_y, g_xt = np.zeros_like(y), np.zeros_like(xt)
    O.binary_ew_grad(O.MUL, y, xt, g_y, g_xt, np.ones_like(out))
    g_x = np.zeros_like(x)
    O.transpose(3, 2, g_xt, g_x, assign=True, quirk=True)  # slice_transpose::<T, Assign>(cols, rows, out_grad, x_grad)
    assert np.allclose(g_x, O.transpose(3, 2, y), atol=1e-2)


def test_tape_l2_norm_cols():
    """tests/matrix/l2_norm_cols.rs:11-21: squared -> sum_cols -> pow(1/2), grad of first row ~ [0.1826,0.3651,0.5477,0.7303]"""
    x = O.arr([1., 2., 3., 4., 5., 6., 7., 8.], np.float64)
    sq = O.unary(O.UN_SQUARE, x)
    sc = O.sum_cols(4, sq)
    out = O.unary(O.UN_POW, sc, 0.5)
    g_sc, g_sq, g_x = np.zeros_like(sc), np.zeros_like(sq), np.zeros_like(x)
    O.unary_grad(O.UN_POW, sc, g_sc, np.ones_like(out), 0.5)
    O.sum_cols_grad(4, g_sq, g_sc)
    O.unary_grad(O.UN_SQUARE, x, g_x, g_sq)
    assert np.allclose(g_x[:4], [0.1826, 0.3651, 0.5477, 0.7303], atol=1e-2)


def test_softmax_grad_closed_form_matches_jacobian():
    """the CUDA path uses s*(g-<s,g>); the reference builds the Jacobian (softmax/grad/cpu.rs:40-60): equal to rounding."""
    rng = np.random.default_rng(0)
    x = rng.uniform(-3, 3, (64, 10)).astype(np.float32).ravel()
    g = rng.uniform(-1, 1, (64, 10)).astype(np.float32).ravel()
    s = O.softmax(64, 10, x)
    a, b = np.zeros_like(x), np.zeros_like(x)
    O.softmax_grad(64, 10, a, s, g)
    O.softmax_grad(64, 10, b, s, g, closed=True)
    assert np.max(np.abs(a - b)) < 2e-7


def test_gemm_restatement_vs_numpy_and_truth():
    rng = np.random.default_rng(42)
    m, k, n = 37, 53, 29
    a = rng.uniform(-1, 1, m * k).astype(np.float32)
    b = rng.uniform(-1, 1, k * n).astype(np.float32)
    c = O.gemm(m, k, n, a, b)
    truth = O.gemm_truth(False, False, m, n, k, a, b)
    assert np.max(np.abs(c - truth)) < k * 2 ** -23 * 4
    assert np.allclose(truth, (a.reshape(m, k).astype(np.float64) @ b.reshape(k, n).astype(np.float64)).ravel(), atol=1e-12)
    og = rng.uniform(-1, 1, m * n).astype(np.float32)
    lg, rg = np.zeros_like(a), np.zeros_like(b)
    O.gemm_grad(m, k, n, a, b, lg, rg, og)
    assert np.allclose(lg.reshape(m, k), og.reshape(m, n) @ b.reshape(k, n).T, atol=1e-4)
    assert np.allclose(rg.reshape(k, n), a.reshape(m, k).T @ og.reshape(m, n), atol=1e-4)


def test_openblas_hook_matches_restatement():
    rng = np.random.default_rng(1)
    dims = [12, 16, 8, 10]
    B = 32
    x = rng.uniform(0, 1, B * dims[0]).astype(np.float32)
    labels = rng.integers(0, 10, B).astype(np.int32)
    y = np.zeros((B, 10), np.float32); y[np.arange(B), labels] = 1; y = y.ravel()
    def init():
        r = np.random.default_rng(5)
        W = [r.uniform(-0.1, 0.1, dims[i] * dims[i + 1]).astype(np.float32) for i in range(3)]
        Bs = [np.zeros(dims[i + 1], np.float32) for i in range(3)]
        return W, Bs
    W0, B0 = init()
    O.use_naive_gemm()
    l0, c0, dW0, dB0 = O.mlp_step(0, dims, x, y, labels, W0, B0, 0.1, want_grads=True)
    if not O.use_openblas(1):
        pytest.skip("no OpenBLAS in this image")
    try:
        W1, B1 = init()
        l1, c1, dW1, dB1 = O.mlp_step(0, dims, x, y, labels, W1, B1, 0.1, want_grads=True)
    finally:
        O.use_naive_gemm()
    assert abs(l0 - l1) < 1e-4 * abs(l0) and c0 == c1
    for a, b in zip(dW0 + dB0 + W0, dW1 + dB1 + W1):
        assert np.allclose(a, b, atol=1e-6)


def test_mlp_step_against_numpy_autodiff():
    """nn.rs step restated with float64 numpy math: loss, grads and the SGD update agree."""
    rng = np.random.default_rng(3)
    dims = [6, 7, 5, 4]
    B = 9
    x = rng.uniform(0, 1, (B, dims[0])).astype(np.float32)
    labels = rng.integers(0, 4, B).astype(np.int32)
    y = np.zeros((B, 4), np.float32); y[np.arange(B), labels] = 1
    W = [rng.uniform(-0.5, 0.5, (dims[i], dims[i + 1])).astype(np.float32) for i in range(3)]
    Bs = [rng.uniform(-0.1, 0.1, dims[i + 1]).astype(np.float32) for i in range(3)]
    Wc = [w.copy().ravel() for w in W]; Bc = [b.copy() for b in Bs]
    loss, correct, dW, dB = O.mlp_step(0, dims, x.ravel(), y.ravel(), labels, Wc, Bc, 0.1, want_grads=True)
    # float64 reference
    a = x.astype(np.float64); acts = [a]; zs = []
    for i in range(3):
        z = a @ W[i].astype(np.float64) + Bs[i]
        zs.append(z)
        a = np.maximum(z, 0) if i < 2 else np.exp(z - z.max(1, keepdims=True)) / np.exp(z - z.max(1, keepdims=True)).sum(1, keepdims=True)
        acts.append(a)
    p = acts[-1]
    ref_loss = -np.log(np.clip(p, 1e-7, 1 - 1e-7)[np.arange(B), labels]).sum()
    assert abs(loss - ref_loss) < 1e-4
    assert correct == int((p.argmax(1) == labels).sum())
    gz = (p - y) / B  # softmax + cce closed form
    for i in (2, 1, 0):
        gw = acts[i].T @ gz; gb = gz.sum(0)
        assert np.allclose(dW[i].reshape(gw.shape), gw, atol=2e-6), i
        assert np.allclose(dB[i], gb, atol=2e-6), i
        assert np.allclose(Wc[i].reshape(gw.shape), W[i] - 0.1 * gw, atol=2e-6)
        if i:
            gz = (gz @ W[i].astype(np.float64).T) * (zs[i - 1] >= 0)


def test_sine_net_trajectory_is_chaotic_evidence():
    """SURVEY 8(d) asks for rel 1e-4 on the sine_net loss after 1000 steps.  Evidence that no fp32 implementation with another
    summation order can meet that: the oracle (sequential loops) and a numpy/BLAS fp32 replay agree to 1e-6 over the first ten
    steps and are > 1e-3 apart at step 1000, both within 10 % of the fp64 replay.  (The GPU test gates on the same yardstick.)"""
    import oracle as O
    from tests import sine_replay as SR
    xs, ys, W, B = SR.problem()
    Wo, Bo = [w.copy() for w in W], [b.copy() for b in B]
    lo = np.array([O.mlp_step(1, SR.DIMS, xs, ys, None, Wo, Bo, 1e-4)[0] for _ in range(1001)])
    l64 = SR.replay(np.float64, 1001, W, xs, ys)
    l32 = SR.replay(np.float32, 1001, W, xs, ys)
    assert np.all(np.abs(lo[:10] - l64[:10]) <= 1e-6 * l64[:10]) and np.all(np.abs(l32[:10] - l64[:10]) <= 1e-6 * l64[:10])
    assert abs(lo[1000] - l32[1000]) > 1e-3 * l32[1000], "two fp32 summation orders stayed within 1e-3: the 1e-4 gate would be meaningful"
    for l in (lo, l32):
        assert abs(l[1000] - l64[1000]) <= 0.10 * l64[1000]
        assert l[1000] < 0.02 * l[0]
