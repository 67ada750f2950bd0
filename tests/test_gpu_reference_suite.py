"""
The reference's integration tests (tests/test_*.rs, tests/matrix/*.rs), transcribed against the CUDA device through the
host layer: build the device, run the *MayGrad op, compose with a second op for a non-trivial upstream gradient,
call backward(), assert the gradients.  Exact `assert_eq!` in the reference -> exact here; `roughly_equals` (abs 1e-2,
src/lib.rs:36-46) -> abs 1e-2 here, plus a tight bound.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def device():
    from sliced_b200.host import CUDA
    d = CUDA(0)          # CUDA::<Autograd<Base>>
    yield d
    d.close()


def roughly_equals(a, b, tol=1e-2):
    assert np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) <= tol


def test_add(device):
    """tests/test_add.rs:4-24"""
    lhs = device.buffer([1, 2, 3, 4, 5], np.int32)
    rhs = device.buffer([6, 7, 8, 9, 10], np.int32)
    out = device.add(lhs, rhs)
    assert out.read().tolist() == [7, 9, 11, 13, 15]
    out.backward()
    assert lhs.grad().read().tolist() == [1, 1, 1, 1, 1]
    assert rhs.grad().read().tolist() == [1, 1, 1, 1, 1]


def test_add2(device):
    lhs = device.buffer([1, 2, 3, 4, 5], np.int32)
    rhs = device.buffer([6, 7, 8, 9, 10], np.int32)
    out = device.add2(lhs, rhs)
    assert out.read().tolist() == [7, 9, 11, 13, 15]
    out.backward()
    assert lhs.grad().read().tolist() == [1] * 5 and rhs.grad().read().tolist() == [1] * 5


def test_sub(device):
    """tests/test_sub.rs:4-24"""
    lhs = device.buffer([1, 2, 3, 4, 5], np.int32)
    rhs = device.buffer([6, 7, 8, 9, 10], np.int32)
    out = device.sub(lhs, rhs)
    assert out.read().tolist() == [-5] * 5
    out.backward()
    assert lhs.grad().read().tolist() == [1] * 5
    assert rhs.grad().read().tolist() == [-1] * 5


def test_mul(device):
    """tests/test_mul.rs:4-24"""
    lhs = device.buffer([1, 2, 3, 4, 5], np.int32)
    rhs = device.buffer([6, 7, 8, 9, 10], np.int32)
    out = device.mul(lhs, rhs)
    assert out.read().tolist() == [6, 14, 24, 36, 50]
    out.backward()
    assert lhs.grad().read().tolist() == [6, 7, 8, 9, 10]
    assert rhs.grad().read().tolist() == [1, 2, 3, 4, 5]


def test_mul_by_itself_and_add_to_itself(device):
    """x.mul(&x) -> 2 x og, add(&x, &x) -> 2 og: both grad closures hit the same gradient buffer (ops.rs:125-126,163-164 on an
    aliased operand; custos hands out one grad buffer per id)"""
    x = device.buffer([1, 2, 3, 4, 5], np.int32)
    out = device.mul(x, x)
    assert out.read().tolist() == [1, 4, 9, 16, 25]
    out.backward()
    assert x.grad().read().tolist() == [2, 4, 6, 8, 10]
    y = device.buffer([1., -2., 3.5], np.float32)
    out = device.add(y, y)
    out.backward()
    assert y.grad().read().tolist() == [2., 2., 2.]
    z = device.buffer([4., 5.], np.float32)
    out = device.sub(z, z)
    out.backward()
    assert z.grad().read().tolist() == [0., 0.]


def test_gradients_of_dropped_buffers_are_released(device):
    """a non-Cached device allocates a fresh buffer (and, in backward, a fresh gradient) per op: the gradient map must not keep
    them alive once the buffers are gone (custos drops them with the buffer, OnDropBuffer †)"""
    x = device.buffer(np.ones(1024, np.float32))
    for _ in range(5):
        out = device.square(device.mul(x, x))
        out.backward()
        del out
    assert device.n_grads() <= 2, device.n_grads()


def test_pow(device):
    """tests/test_pow.rs:6-21 (f64, exact)"""
    x = device.buffer([1., 2., 3., 4., 5.], np.float64)
    out = device.pow(x, 3.)
    assert out.read().tolist() == [1., 8., 27., 64., 125.]
    out.backward()
    assert x.grad().read().tolist() == [3., 12., 27., 48., 75.]


def test_square(device):
    """tests/test_square.rs:4-20"""
    x = device.buffer([1, 2, 3, 4, 5], np.int32)
    out = device.square(x)
    assert out.read().tolist() == [1, 4, 9, 16, 25]
    out.backward()
    assert x.grad().read().tolist() == [2, 4, 6, 8, 10]


def test_gemm(device):
    """tests/test_gemm.rs:4-43"""
    lhs = device.buffer([1., 2., 3., 4., 4., 5., 6., 5.], np.float32)
    rhs = device.buffer([1., 2., 3., 4., 5., 6.], np.float32)
    out = device.gemm(4, 2, 3, lhs, rhs)
    assert out.read().tolist() == [9.0, 12.0, 15.0, 19.0, 26.0, 33.0, 24.0, 33.0, 42.0, 26.0, 37.0, 48.0]
    out.backward()
    assert lhs.grad().read().tolist() == [6.0, 15.0] * 4
    assert rhs.grad().read().tolist() == [14.0, 14.0, 14.0, 16.0, 16.0, 16.0]


def test_gemm_no_grad_operand(device):
    """gemm/grad/cpu_stack.rs:35-40: each side guarded by requires_grad() (x.no_grad() in examples/nn.rs:170)"""
    lhs = device.buffer([1., 2., 3., 4., 4., 5., 6., 5.], np.float32).no_grad()
    rhs = device.buffer([1., 2., 3., 4., 5., 6.], np.float32)
    out = device.gemm(4, 2, 3, lhs, rhs)
    out.backward()
    assert rhs.grad().read().tolist() == [14.0, 14.0, 14.0, 16.0, 16.0, 16.0]
    assert lhs.grad().read().tolist() == [0.0] * 8


def test_row_op(device):
    """tests/test_row_op.rs:3-38"""
    buf = device.buffer([-1, -3, 4, 5, 2, 2, 1, 5, -3, 2, 2, 4, 3, 2, -1], np.int32)
    row_add = device.buffer([1, 2, 3, 4, 5], np.int32)
    out = device.add_row(3, 5, buf, row_add)
    assert out.read().tolist() == [0, -1, 7, 9, 7, 3, 3, 8, 1, 7, 3, 6, 6, 6, 4]
    out.backward()
    assert row_add.grad().read().tolist() == [3] * 5
    assert buf.grad().read().tolist() == [1] * 15


def test_row_op_mut(device):
    """tests/test_row_op.rs:42-77"""
    buf = device.buffer([-1, -3, 4, 5, 2, 2, 1, 5, -3, 2, 2, 4, 3, 2, -1], np.int32)
    row_add = device.buffer([1, 2, 3, 4, 5], np.int32)
    device.add_row_mut(3, 5, buf, row_add)
    assert buf.read().tolist() == [0, -1, 7, 9, 7, 3, 3, 8, 1, 7, 3, 6, 6, 6, 4]
    buf.backward()
    assert row_add.grad().read().tolist() == [3] * 5
    assert buf.grad().read().tolist() == [1] * 15


def test_max_cols(device):
    """tests/test_max_cols.rs:3-35"""
    rhs = device.buffer([1, 4, 2], np.int32)
    lhs = device.buffer([-3, 2, 3, 1, 1, 5, -5, 4, -9, -2, -4, -1], np.int32)
    max_cols = device.max_cols(3, 4, lhs)
    assert max_cols.read().tolist() == [3, 5, -1]
    out = device.add(max_cols, rhs)
    out.backward()
    assert lhs.grad().read().tolist() == [0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1]
    assert rhs.grad().read().tolist() == [1, 1, 1]


def test_max_rows(device):
    """tests/test_max_rows.rs:3-34"""
    rhs = device.buffer([2, 3, 4, 1], np.int32)
    lhs = device.buffer([-3, 2, 3, 1, 1, 5, -5, 4, -9, -2, -4, -1], np.int32)
    max_rows = device.max_rows(4, lhs)
    out = device.mul(max_rows, rhs)
    out.backward()
    assert lhs.grad().read().tolist() == [0, 0, 4, 0, 2, 3, 0, 1, 0, 0, 0, 0]
    assert rhs.grad().read().tolist() == [1, 5, 3, 4]


def test_mean_rows(device):
    """tests/test_mean.rs:3-33"""
    lhs = device.buffer([2., 3., 4.], np.float64)
    to_mean_rows = device.buffer([1., 4., 3., 2., 3., 5., 2., 1., 6., -2., 1., 4.], np.float64)
    mean_rows = device.mean_rows(3, to_mean_rows)
    out = device.sub(lhs, mean_rows)
    out.backward()
    assert to_mean_rows.grad().read().tolist() == [-0.25] * 12
    assert lhs.grad().read().tolist() == [1., 1., 1.]


def test_sum_cols(device):
    """tests/test_sum_cols.rs:3-32"""
    rhs = device.buffer([1, 4, 2], np.int32)
    to_sum_cols = device.buffer([4, 2, 1, 3, 6, 1, 3, 1, 5, 4, 1, 1], np.int32)
    x = device.sum_cols(4, to_sum_cols)
    out = device.mul(x, rhs)
    out.backward()
    assert to_sum_cols.grad().read().tolist() == [1, 1, 1, 1, 4, 4, 4, 4, 2, 2, 2, 2]
    assert rhs.grad().read().tolist() == [10, 11, 11]


def test_sum_rows(device):
    """tests/test_sum_rows.rs:2-31 — the reference itself panics here (src/ops.rs:556 unimplemented!()); its vectors are the spec"""
    rhs = device.buffer([1, 4, 2, 3], np.int32)
    to_sum_rows = device.buffer([4, 2, 1, 3, 6, 1, 3, 1, 5, 4, 1, 1], np.int32)
    x = device.sum_rows(4, to_sum_rows)
    out = device.mul(x, rhs)
    out.backward()
    assert to_sum_rows.grad().read().tolist() == [1, 4, 2, 3] * 3
    assert rhs.grad().read().tolist() == [15, 7, 5, 5]


def test_transpose(device):
    """tests/test_transpose.rs:5-13"""
    x = device.buffer([1, 2, 3, 4, 5, 6], np.int32)
    out = device.transpose(2, 3, x)
    assert out.read().tolist() == [1, 4, 2, 5, 3, 6]


# ---------------------------------------------------------------- tests/matrix/*.rs
def test_matrix_transpose_grad(device):
    """tests/matrix/transpose.rs:5-23"""
    from sliced_b200.host import Matrix
    x = Matrix(device, 2, 3, [1., 2., 3., 4., 5., 6.], np.float64)
    out = x.T()
    y = Matrix(device, 3, 2, [-1., 1., 2., -5., -3., 2.], np.float64)
    out = device.mul(y.buf, out.buf)
    out.backward()
    expected = device.transpose(y.rows, y.cols, y.buf).read()
    roughly_equals(expected, x.grad().read(), 0.0)


def test_matrix_relu(device):
    """tests/matrix/relu.rs:3-21"""
    from sliced_b200.host import Matrix
    buf = Matrix(device, 1, 5, [-1., -3., 2., 5., -1.3], np.float64)
    out = buf.relu()
    assert out.read().tolist() == [0., 0., 2., 5., 0.]
    out.backward()
    assert buf.grad().read().tolist() == [0., 0., 1., 1., 0.]


def test_matrix_tanh(device):
    """tests/matrix/tanh.rs:3-33"""
    from sliced_b200.host import Matrix
    buf = Matrix(device, 1, 6, [-1., -3., 2., 5., -1.3, 0.], np.float32)
    out = buf.tanh()
    roughly_equals(out.read(), [-0.7615942, -0.9950548, 0.9640276, 0.9999092, -0.8617231, 0.], 2e-7)
    out.backward()
    roughly_equals(buf.grad().read(), [0.41997433, 0.009865999, 0.070650816, 0.00018155575, 0.25743324, 1.0], 2e-7)


def test_matrix_softmax_grad_is_zero_for_ones_upstream(device):
    """tests/matrix/softmax.rs:3-17"""
    from sliced_b200.host import Matrix
    x = Matrix(device, 2, 3, [1., 2., 3., 4., 5., 6.], np.float64)
    out = x.softmax()
    out.backward()
    roughly_equals([0.] * 6, x.grad().read(), 1e-15)


def test_matrix_l2_norm_cols(device):
    """tests/matrix/l2_norm_cols.rs:5-22 (device = Autograd<Cached<Base>> there)"""
    from sliced_b200.host import CUDA, Matrix
    dev = CUDA(0, cached=True)
    lhs = Matrix(dev, 2, 4, [1., 2., 3., 4., 5., 6., 7., 8.], np.float64)
    out = lhs.l2_norm_cols()
    out.backward()
    g = lhs.grad().read()
    roughly_equals([0.1826, 0.3651, 0.5477, 0.7303], g[:4])
    ref = np.array([1., 2., 3., 4., 5., 6., 7., 8.]).reshape(2, 4)
    roughly_equals((ref / np.sqrt((ref ** 2).sum(1, keepdims=True))).ravel(), g, 1e-12)
    dev.close()


def test_softmax_and_grad_known_answers(device):
    """src/ops2/softmax/cpu.rs:25-35, softmax/grad/cpu.rs:71-96 through the tape with a weighted upstream gradient"""
    x = device.buffer([1., 2., 3., 4., 5., 6.], np.float32)
    w = device.buffer([1., 2., 3., 4., 5., 6.], np.float32).no_grad()
    s = device.softmax(2, 3, x)
    roughly_equals(s.read(), [0.09003057, 0.24472847, 0.66524096] * 2, 2e-7)
    out = device.mul(s, w)       # d(out)/d(s) = w  -> softmax_grad sees out_grad = [1..6]
    out.backward()
    roughly_equals(x.grad().read(), [-0.1418171, -0.14077032, 0.28258747] * 2, 5e-7)


def test_sigmoid_matches_closed_form(device):
    """src/matrix.rs:234-262"""
    from sliced_b200.host import Matrix
    v = np.array([-3., -1., 0., 0.5, 2., 5.])
    m = Matrix(device, 1, 6, v, np.float64)
    out = m.sigmoid()
    s = 1 / (1 + np.exp(-v))
    roughly_equals(out.read(), s, 1e-15)
    out.backward()
    roughly_equals(m.grad().read(), s * (1 - s), 1e-15)


def test_gradient_descent_smoke(device):
    """tests/test_combination.rs:5-23 / tests/matrix/min_fn.rs:3-22: 100 steps of x -= 0.1 * grad(x^2) shrink |x|"""
    x = device.buffer([10., -10., 10., -5., 6., 3., 1.], np.float32)
    for _ in range(100):
        device.zero_grad()
        out = device.square(x)
        out.backward()
        device.sgd_step(x, 0.1)
    assert np.max(np.abs(x.read())) < 1e-7


def test_cached_device_reuses_buffers_and_stays_correct():
    """custos Cached + Cursor (examples/nn.rs:156,184): same buffers every iteration, SET ops must fully overwrite them"""
    from sliced_b200.host import CUDA
    dev = CUDA(0, cached=True)
    x = dev.buffer(np.full(1000, 1.3, np.float32))
    b = dev.buffer(np.full(1000, 2.1, np.float32))
    ptrs = None
    for _ in dev.range(5):          # examples/chained_perf.rs:73-91
        squared = dev.square(x)
        add = dev.add(b, x)
        mul_b = dev.mul(add, b)
        mul = dev.mul(squared, x)
        out = dev.add(mul, mul_b)
        assert out.read()[0] == np.float32(9.336999)
        cur = [t.ptr for t in (squared, add, mul_b, mul, out)]
        assert ptrs is None or cur == ptrs, "Cached device must hand out the same buffers each iteration"
        ptrs = cur
    dev.close()


def test_chained_graph_with_backward(device):
    """examples/chained_perf.rs:86-93 with backward() enabled (BASELINE config 4), against the oracle tape replay"""
    import oracle as O
    rng = np.random.default_rng(0)
    n = 10007
    xh, bh = rng.uniform(-2, 2, n).astype(np.float32), rng.uniform(-2, 2, n).astype(np.float32)
    x, b = device.buffer(xh), device.buffer(bh)
    squared = device.square(x)
    add = device.add(b, x)
    mul_b = device.mul(add, b)
    mul = device.mul(squared, x)
    out = device.add(mul, mul_b)
    assert np.array_equal(out.read(), O.chained_fwd(xh, bh))
    out.backward()
    xg, bg = np.zeros_like(xh), np.zeros_like(bh)
    O.chained_bwd(xh, bh, xg, bg, np.ones_like(xh))
    assert np.array_equal(x.grad().read(), xg) and np.array_equal(b.grad().read(), bg)
