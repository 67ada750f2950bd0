"""
Parity of the CUDA path (through the C ABI) with the CPU oracle on seeded random inputs, at sizes the oracle finishes in
seconds, including the edge cases (empty, single element, ragged tails, 16-byte-misaligned pointers, row/col counts that
defeat vectorisation) — and size-independent properties at the BASELINE sizes (16384 x 16384).

Tolerances (stated, per BASELINE.json north_star):
  * integer dtypes, transpose, argmax, every single-IEEE-op element-wise kernel (add/sub/mul/div, row_op, col_op, sgd,
    the ACC grads built from one mul + one add, the chained graph): BIT-EXACT vs the oracle (built with -fmad=false /
    -ffp-contract=off on both sides)
  * transcendental unary ops (pow/exp/ln/tanh/sigmoid): rel 1e-6 (CUDA libm vs glibc differ in the last ulps)
  * reductions (sum/mean rows/cols, scalar sum, softmax): K-scaled, |err| <= 4 * K * 2^-24 * max|x| (different, but
    deterministic, summation tree; the reference sums left to right)
"""
import numpy as np
import pytest

import oracle as O
from tests.oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu

DT = {"f32": np.float32, "f64": np.float64, "i32": np.int32}
EPS = {"f32": 2.0 ** -24, "f64": 2.0 ** -53}


@pytest.fixture(scope="module")
def be():
    from tests.cuda_backend import CudaBackend
    return CudaBackend()


@pytest.fixture(scope="module")
def ob():
    return OracleBackend()


def rnd(rng, n, dt, lo=-1.0, hi=1.0):
    if np.issubdtype(DT[dt], np.integer):
        return rng.integers(-50, 50, n).astype(np.int32)
    return rng.uniform(lo, hi, n).astype(DT[dt])


def exact(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    if a.dtype.kind == "f":
        assert np.array_equal(a.view(np.uint32 if a.dtype == np.float32 else np.uint64),
                              b.view(np.uint32 if b.dtype == np.float32 else np.uint64)), \
            f"{what}: not bit-exact, max abs diff {np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))}"
    else:
        assert np.array_equal(a, b), what


def close_rel(a, b, rel, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    assert np.all((err <= rel) | (np.abs(a - b) <= 1e-37)), f"{what}: max rel err {err.max()} > {rel}"


SIZES = [0, 1, 3, 4, 5, 1023, 1024, 4099, 262147]


@pytest.mark.parametrize("dt", ["f32", "f64", "i32"])
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("op", [O.ADD, O.SUB, O.MUL, O.DIV])
def test_binary_ew_and_grad(be, ob, dt, n, op):
    rng = np.random.default_rng(1234 + n + op)
    lhs, rhs, og = rnd(rng, n, dt), rnd(rng, n, dt), rnd(rng, n, dt)
    if op == O.DIV:
        rhs = np.where(np.abs(rhs) < (1 if dt == "i32" else 0.05), DT[dt](3), rhs).astype(DT[dt])
    exact(be.binary_ew(op, lhs, rhs), ob.binary_ew(op, lhs, rhs), "fwd")
    lg0, rg0 = rnd(rng, n, dt), rnd(rng, n, dt)  # non-zero: ACC vs SET mistakes are visible
    a = be.binary_ew_grad(op, lhs, rhs, lg0.copy(), rg0.copy(), og)
    b = ob.binary_ew_grad(op, lhs, rhs, lg0.copy(), rg0.copy(), og)
    exact(a[0], b[0], "lhs_grad"); exact(a[1], b[1], "rhs_grad")
    # one-sided (the other operand is no_grad)
    a = be.binary_ew_grad(op, lhs, rhs, lg0.copy(), None, og)
    b = ob.binary_ew_grad(op, lhs, rhs, lg0.copy(), None, og)
    exact(a[0], b[0], "lhs_grad only")
    a = be.binary_ew_grad(op, lhs, rhs, None, rg0.copy(), og)
    b = ob.binary_ew_grad(op, lhs, rhs, None, rg0.copy(), og)
    exact(a[1], b[1], "rhs_grad only")


@pytest.mark.parametrize("dt", ["f32", "i32"])
@pytest.mark.parametrize("op", [O.ADD, O.SUB, O.MUL, O.DIV])
def test_binary_ew_grad_same_buffer_for_both_grads(be, dt, op):
    """x.mul(x) / add(x, x): lhs_grad and rhs_grad are ONE buffer; the reference's loop adds both contributions in turn
    (binary_ew/grad/cpu_stack.rs:54-59).  Bit-exact vs the oracle called with the same aliased array."""
    import sliced_b200 as S
    ctx = be.ctx
    rng = np.random.default_rng(5 + op)
    for n in (1, 5, 4099, 262147):
        x, og, g0 = rnd(rng, n, dt), rnd(rng, n, dt), rnd(rng, n, dt)
        if op == O.DIV:
            x = np.where(np.abs(x) < (1 if dt == "i32" else 0.05), DT[dt](3), x).astype(DT[dt])
        ref = g0.copy()
        O.binary_ew_grad(op, x, x, ref, ref, og)
        dg, dx = ctx.array(g0), ctx.array(x)
        ctx.binary_ew_grad(op, dx, dx, dg, dg, ctx.array(og))
        exact(dg.numpy(), ref, f"aliased grads op {op} n {n}")
        ref2 = g0.copy()
        O.add_ew_grad(ref2, ref2, og)
        dg2 = ctx.array(g0)
        ctx.add_ew_grad(dg2, dg2, ctx.array(og))
        exact(dg2.numpy(), ref2, "aliased add_ew_grad")


def test_misaligned_pointers(be):
    """sub-buffers that start 4 bytes into an allocation take the scalar path and still match"""
    import sliced_b200 as S
    ctx = be.ctx
    rng = np.random.default_rng(7)
    n = 10007
    a, b = rnd(rng, n + 1, "f32"), rnd(rng, n + 1, "f32")
    da, db, dout = ctx.array(a), ctx.array(b), ctx.zeros(n + 1, np.float32)
    ctx.binary_ew(S.MUL, da.view(1, n), db.view(1, n), dout.view(1, n))
    exact(dout.numpy()[1:], a[1:] * b[1:])
    assert dout.numpy()[0] == 0
    dout2 = ctx.zeros(n + 1, np.float32)
    ctx.unary(S.UN_SQUARE, da.view(1, n), out=dout2.view(1, n))
    exact(dout2.numpy()[1:], a[1:] * a[1:])
    rows, cols = 37, 270  # row_op on a misaligned matrix
    m = rnd(rng, rows * cols + 1, "f32"); v = rnd(rng, cols, "f32")
    dm, dv, do = ctx.array(m), ctx.array(v), ctx.zeros(rows * cols + 1, np.float32)
    ctx.row_op(S.ADD, cols, dm.view(1, rows * cols), dv, do.view(1, rows * cols))
    exact(do.numpy()[1:], (m[1:].reshape(rows, cols) + v).ravel())


EXACT_UNOPS = [O.UN_SQUARE, O.UN_RELU, O.UN_CLIP, O.UN_NEG, O.UN_MUL_SCALAR, O.UN_NEG_DIV_SCALAR, O.UN_ADD_SCALAR]
LIBM_UNOPS = [O.UN_POW, O.UN_TANH, O.UN_SIGMOID, O.UN_EXP, O.UN_LN, O.UN_NEG_LN]


@pytest.mark.parametrize("dt", ["f32", "f64", "i32"])
@pytest.mark.parametrize("n", [0, 1, 7, 4096, 100003])
@pytest.mark.parametrize("op", EXACT_UNOPS)
def test_unary_exact(be, ob, dt, n, op):
    rng = np.random.default_rng(99 + n + op)
    x, og, xg0 = rnd(rng, n, dt), rnd(rng, n, dt), rnd(rng, n, dt)
    p0, p1 = (-0.25, 0.5) if op == O.UN_CLIP else ((3.0, 0.0) if dt == "i32" else (0.37, 0.0))
    if dt == "i32" and op == O.UN_CLIP:
        p0, p1 = -10.0, 20.0
    exact(be.unary(op, x, p0, p1), ob.unary(op, x, p0, p1), "fwd")
    exact(be.unary_grad(op, x, xg0.copy(), og, p0, p1), ob.unary_grad(op, x, xg0.copy(), og, p0, p1), "grad")


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("op", LIBM_UNOPS)
def test_unary_libm(be, ob, dt, op):
    rng = np.random.default_rng(5 + op)
    n = 50001
    x = rnd(rng, n, dt, 0.5, 2.0) if op in (O.UN_POW, O.UN_LN, O.UN_NEG_LN) else rnd(rng, n, dt, -4.0, 4.0)
    og = rnd(rng, n, dt)
    p0 = 3.0 if op == O.UN_POW else 0.0
    rel = 1e-6 if dt == "f32" else 1e-14
    close_rel(be.unary(op, x, p0), ob.unary(op, x, p0), rel, "fwd")
    z = np.zeros_like(x)
    ga, gb = be.unary_grad(op, x, z.copy(), og, p0), ob.unary_grad(op, x, z.copy(), og, p0)
    if op in (O.UN_TANH, O.UN_SIGMOID):
        # 1 - tanh(x)^2 and e/(1+e)^2 cancel: a 1-ulp difference in tanhf/expf is amplified relative to a result << 1,
        # so the derivative is held to 1e-6 of the function scale (|f'| <= 1), not of the (tiny) value
        assert np.max(np.abs(ga.astype(np.float64) - gb.astype(np.float64))) <= (1e-6 if dt == "f32" else 1e-14)
    else:
        close_rel(ga, gb, 2 * rel, "grad")
    if op == O.UN_POW:  # the exponent sine_net uses, and a fractional one
        for p in (2.0, 0.5):
            close_rel(be.unary(op, x, p), ob.unary(op, x, p), rel, f"pow {p}")


def test_unary_int_rejects_float_only_ops(be):
    import sliced_b200 as S
    x = be.ctx.array(np.arange(5, dtype=np.int32))
    with pytest.raises(S.SlicedError) as e:
        be.ctx.unary(S.UN_EXP, x)
    assert e.value.code == -3


SHAPES = [(1, 1), (1, 7), (7, 1), (3, 5), (33, 10), (17, 64), (64, 260), (129, 1000), (70, 4100), (5, 9001), (1000, 12)]


@pytest.mark.parametrize("dt", ["f32", "f64", "i32"])
@pytest.mark.parametrize("shape", SHAPES)
def test_row_col_ops(be, ob, dt, shape):
    rows, cols = shape
    rng = np.random.default_rng(rows * 131 + cols)
    lhs, rv, cv = rnd(rng, rows * cols, dt), rnd(rng, cols, dt), rnd(rng, rows, dt)
    og = rnd(rng, rows * cols, dt)
    for op in (O.ADD, O.SUB, O.MUL):
        exact(be.row_op(op, cols, lhs, rv), ob.row_op(op, cols, lhs, rv), f"row_op {op}")
        exact(be.col_op(op, cols, lhs, cv), ob.col_op(op, cols, lhs, cv), f"col_op {op}")
    cvd = np.where(cv == 0, DT[dt](2), cv).astype(DT[dt])
    exact(be.col_op(O.DIV, cols, lhs, cvd), ob.col_op(O.DIV, cols, lhs, cvd), "div_cols")
    exact(be.add_row_mut(rows, cols, lhs.copy(), rv), ob.add_row_mut(rows, cols, lhs.copy(), rv), "add_row_mut")
    # grads: lhs_grad is elementwise (exact); rhs_grad is a column sum (ints exact, floats K-scaled)
    lg0, rg0 = rnd(rng, rows * cols, dt), rnd(rng, cols, dt)
    a = be.add_row_grad(rows, cols, lg0.copy(), rg0.copy(), og)
    b = ob.add_row_grad(rows, cols, lg0.copy(), rg0.copy(), og)
    exact(a[0], b[0], "add_row_grad lhs (SET copy)")
    tol = 0 if dt == "i32" else 4 * rows * EPS[dt] * 1.0 + EPS[dt] * 4
    assert np.max(np.abs(a[1].astype(np.float64) - b[1].astype(np.float64)), initial=0) <= tol, "add_row_grad rhs"
    a = be.add_row_mut_grad(rows, cols, rg0.copy(), og); b = ob.add_row_mut_grad(rows, cols, rg0.copy(), og)
    assert np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)), initial=0) <= tol
    for op in (O.ADD, O.SUB, O.MUL):
        a = be.row_op_grad(op, cols, lhs, rv, lg0.copy(), rg0.copy(), og)
        b = ob.row_op_grad(op, cols, lhs, rv, lg0.copy(), rg0.copy(), og)
        exact(a[0], b[0], f"row_op_grad lhs {op}")
        assert np.max(np.abs(a[1].astype(np.float64) - b[1].astype(np.float64)), initial=0) <= (0 if dt == "i32" else tol * 2)
    if dt != "i32":
        cg0 = rnd(rng, rows, dt)
        a = be.col_op_grad(O.DIV, cols, lhs, cvd, lg0.copy(), cg0.copy(), og)
        b = ob.col_op_grad(O.DIV, cols, lhs, cvd, lg0.copy(), cg0.copy(), og)
        exact(a[0], b[0], "col_op_grad lhs")
        scale = np.max(np.abs(lhs)) / np.min(np.abs(cvd)) ** 2
        assert np.max(np.abs(a[1].astype(np.float64) - b[1].astype(np.float64))) <= 4 * cols * EPS[dt] * scale + 1e-30


@pytest.mark.parametrize("dt", ["f32", "f64", "i32"])
@pytest.mark.parametrize("shape", SHAPES + [(300, 33), (2, 70000)])
def test_reductions(be, ob, dt, shape):
    rows, cols = shape
    rng = np.random.default_rng(rows * 17 + cols)
    x = rnd(rng, rows * cols, dt)
    fl = dt != "i32"
    def chk(a, b, k, what):
        if fl:
            assert np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) <= 4 * k * EPS[dt] + 1e-30, what
        else:
            exact(a, b, what)
    chk(be.sum_rows(cols, x), ob.sum_rows(cols, x), rows, "sum_rows")
    chk(be.sum_cols(cols, x), ob.sum_cols(cols, x), cols, "sum_cols")
    chk(be.mean_rows(cols, x), ob.mean_rows(cols, x), rows, "mean_rows")
    chk(be.mean_cols(cols, x), ob.mean_cols(cols, x), cols, "mean_cols")
    exact(be.max_rows(cols, x), ob.max_rows(cols, x), "max_rows")
    exact(be.max_cols(cols, x), ob.max_cols(cols, x), "max_cols")
    chk([be.sum(x)], [ob.sum(x)], rows * cols, "sum")
    chk([be.mean(x)], [ob.mean(x)], rows * cols, "mean")
    exact(np.array([be.max(x)]), np.array([ob.max(x)]), "max")
    # grads: pure broadcast RMW -> exact
    xg0 = rnd(rng, rows * cols, dt); ogc = rnd(rng, cols, dt); ogr = rnd(rng, rows, dt)
    exact(be.sum_rows_grad(cols, xg0.copy(), ogc), ob.sum_rows_grad(cols, xg0.copy(), ogc), "sum_rows_grad")
    exact(be.sum_cols_grad(cols, xg0.copy(), ogr), ob.sum_cols_grad(cols, xg0.copy(), ogr), "sum_cols_grad")
    exact(be.mean_rows_grad(cols, xg0.copy(), ogc), ob.mean_rows_grad(cols, xg0.copy(), ogc), "mean_rows_grad")
    exact(be.mean_cols_grad(cols, xg0.copy(), ogr), ob.mean_cols_grad(cols, xg0.copy(), ogr), "mean_cols_grad")


@pytest.mark.parametrize("dt", ["f32", "i32"])
@pytest.mark.parametrize("shape", [(3, 4), (33, 10), (64, 260), (129, 1000), (40, 5000), (1000, 12)])
def test_max_ties_and_argmax(be, ob, dt, shape):
    """deliberate ties (values quantised to 1/8): max_rows_grad feeds EVERY tie, max_cols_grad the FIRST (Appendix A.8)"""
    rows, cols = shape
    rng = np.random.default_rng(rows + cols)
    x = (rng.integers(-8, 8, rows * cols) / (8 if dt == "f32" else 1)).astype(DT[dt])
    if dt == "i32":
        x = rng.integers(-3, 3, rows * cols).astype(np.int32)
    xg0 = rnd(rng, rows * cols, dt)
    mr, mc = ob.max_rows(cols, x), ob.max_cols(cols, x)
    ogc, ogr = rnd(rng, cols, dt), rnd(rng, rows, dt)
    exact(be.max_rows_grad(cols, mr, x, xg0.copy(), ogc), ob.max_rows_grad(cols, mr, x, xg0.copy(), ogc), "max_rows_grad")
    exact(be.max_cols_grad(cols, mc, x, xg0.copy(), ogr), ob.max_cols_grad(cols, mc, x, xg0.copy(), ogr), "max_cols_grad")
    ctx = be.ctx
    dx = ctx.array(x)
    out, idx = ctx.max_cols(rows, cols, dx, with_idx=True)
    exact(out.numpy(), mc)
    assert np.array_equal(idx.numpy(), x.reshape(rows, cols).argmax(1).astype(np.int32))  # numpy argmax = first maximum
    out, idx = ctx.max_rows(cols, dx, with_idx=True)
    exact(out.numpy(), mr)
    assert np.array_equal(idx.numpy(), x.reshape(rows, cols).argmax(0).astype(np.int32))
    # saved-argmax variant of the grad gives the same bits
    dxg = ctx.array(xg0)
    _, idx = ctx.max_cols(rows, cols, dx, with_idx=True)
    ctx.max_cols_grad_idx(cols, idx, dxg, ctx.array(ogr))
    exact(dxg.numpy(), ob.max_cols_grad(cols, mc, x, xg0.copy(), ogr), "max_cols_grad_idx")


@pytest.mark.parametrize("dt", ["f32", "f64", "i32"])
@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (64, 64), (65, 63), (1, 1000), (1000, 1), (257, 129), (300, 4100)])
def test_transpose(be, ob, dt, shape):
    rows, cols = shape
    rng = np.random.default_rng(3)
    x = rnd(rng, rows * cols, dt)
    exact(be.transpose(rows, cols, x), ob.transpose(rows, cols, x), "SET")
    exact(be.transpose(rows, cols, x), x.reshape(rows, cols).T.ravel(), "SET vs numpy")
    base = rnd(rng, rows * cols, dt)
    exact(be.transpose(rows, cols, x, base.copy(), accumulate=True), ob.transpose(rows, cols, x, base.copy(), accumulate=True), "ACC")


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (65536 // 64, 10), (100, 33), (64, 1000), (20, 1024), (9, 4096), (5, 16384), (3, 5003), (2, 40000)])
def test_softmax(be, ob, dt, shape):
    samples, features = shape
    rng = np.random.default_rng(samples + features)
    x = rnd(rng, samples * features, dt, -5, 5)
    g = rnd(rng, samples * features, dt)
    s_ref = ob.softmax(samples, features, x)
    s = be.softmax(samples, features, x)
    tol = 4 * features * EPS[dt] * np.max(s_ref) + 4 * EPS[dt]
    assert np.max(np.abs(s.astype(np.float64) - s_ref.astype(np.float64))) <= tol
    assert np.allclose(s.reshape(samples, features).sum(1), 1.0, atol=4 * features * EPS[dt] + 1e-6)
    xg = be.softmax_grad(samples, features, rnd(rng, samples * features, dt), s_ref, g)  # SET: the junk prefill must vanish
    ref = np.zeros_like(x)
    if features <= 512:
        O.softmax_grad(samples, features, ref, s_ref, g)          # the reference's Jacobian form
    else:
        O.softmax_grad(samples, features, ref, s_ref, g, closed=True)  # O(F^2) per row is infeasible
    assert np.max(np.abs(xg.astype(np.float64) - ref.astype(np.float64))) <= 8 * features * EPS[dt] * np.max(s_ref) + 4 * EPS[dt]


@pytest.mark.parametrize("n", [0, 1, 5, 100003])
def test_sgd_and_chained(be, ob, n):
    rng = np.random.default_rng(n)
    w, g = rnd(rng, n, "f32"), rnd(rng, n, "f32")
    exact(be.sgd_step(w.copy(), g, 0.1), ob.sgd_step(w.copy(), g, 0.1), "sgd")
    x, b, og = rnd(rng, n, "f32", -2, 2), rnd(rng, n, "f32", -2, 2), rnd(rng, n, "f32")
    exact(be.chained_fwd(x, b), ob.chained_fwd(x, b), "chained fwd (5 reference passes fused)")
    xg0, bg0 = rnd(rng, n, "f32"), rnd(rng, n, "f32")
    a = be.chained_bwd(x, b, xg0.copy(), bg0.copy(), og); r = ob.chained_bwd(x, b, xg0.copy(), bg0.copy(), og)
    exact(a[0], r[0], "chained bwd x"); exact(a[1], r[1], "chained bwd b")


def test_misc_ops(be, ob):
    rng = np.random.default_rng(0)
    x = rnd(rng, 37, "f32")
    exact(be.diagflat(x), ob.diagflat(x))
    og = rnd(rng, 37 * 37, "f32"); xg0 = rnd(rng, 37, "f32")
    exact(be.diagflat_grad(xg0.copy(), og), ob.diagflat_grad(xg0.copy(), og))
    cl = rng.integers(0, 10, 1000).astype(np.float32)
    exact(be.onehot(cl), ob.onehot(cl))
    og = rnd(rng, 1000 * 10, "f32"); cg0 = rnd(rng, 1000, "f32")
    exact(be.onehot_grad(10, cl, cg0.copy(), og), ob.onehot_grad(10, cl, cg0.copy(), og))
    preds = rnd(rng, 1000 * 10, "f32"); labels = rng.integers(0, 10, 1000).astype(np.int32)
    ctx = be.ctx
    assert ctx.count_correct(1000, 10, ctx.array(preds), ctx.array(labels)) == int((preds.reshape(1000, 10).argmax(1) == labels).sum())


# ---------------------------------------------------------------- BASELINE sizes: 16384 x 16384 f32 through properties
R = C = 16384


@pytest.fixture(scope="module")
def big(be):
    rng = np.random.default_rng(1234)
    x = rng.uniform(-1, 1, R * C).astype(np.float32)
    return x, be.ctx.array(x)


def test_full_size_elementwise_and_transpose(be, big):
    import sliced_b200 as S
    x, dx = big
    ctx = be.ctx
    t = ctx.transpose(R, C, dx)
    tt = ctx.transpose(C, R, t)
    d = ctx.binary_ew(S.SUB, tt, dx)                      # transpose is an involution, bit-exact
    assert ctx.max(ctx.unary(S.UN_SQUARE, d)) == 0.0
    # spot-check a strip against numpy
    ht = t.numpy().reshape(C, R)
    assert np.array_equal(ht[:4], x.reshape(R, C)[:, :4].T)
    del ht
    y = ctx.binary_ew(S.ADD, dx, t)
    z = ctx.binary_ew(S.SUB, y, t)                        # (x + t) - t == x within 1 ulp of |x+t|
    err = ctx.max(ctx.unary(S.UN_SQUARE, ctx.binary_ew(S.SUB, z, dx)))
    assert err <= (2.0 ** -23) ** 2


def test_full_size_reductions_agree(be, big):
    import sliced_b200 as S
    x, dx = big
    ctx = be.ctx
    sr = ctx.sum_rows(C, dx).numpy()                      # reduce over rows
    sc_t = ctx.sum_cols(R, ctx.transpose(R, C, dx)).numpy()  # the same numbers through the row-direction kernel
    ref = x.reshape(R, C).sum(0, dtype=np.float64)
    tol = 4 * R * 2.0 ** -24
    assert np.max(np.abs(sr - ref)) <= tol and np.max(np.abs(sc_t - ref)) <= tol
    total = float(ctx.sum(dx))
    assert abs(total - float(x.sum(dtype=np.float64))) <= 4 * R * C * 2.0 ** -24 * 1e-2 + 1.0  # checksum of checksums
    assert abs(float(sr.sum(dtype=np.float64)) - total) <= 2.0
    mr, idx = ctx.max_rows(C, dx, with_idx=True)
    assert np.array_equal(mr.numpy(), x.reshape(R, C).max(0))
    assert np.array_equal(idx.numpy(), x.reshape(R, C).argmax(0).astype(np.int32))
    mc, idx = ctx.max_cols(R, C, dx, with_idx=True)
    assert np.array_equal(mc.numpy(), x.reshape(R, C).max(1))
    assert np.array_equal(idx.numpy(), x.reshape(R, C).argmax(1).astype(np.int32))


def test_full_size_softmax_rows_sum_to_one(be, big):
    x, dx = big
    ctx = be.ctx
    s = ctx.softmax(R, C, dx)
    sums = ctx.sum_cols(C, s).numpy()
    assert np.max(np.abs(sums - 1.0)) <= 4 * C * 2.0 ** -24
    g = ctx.full(R * C, 1.0)
    xg = ctx.full(R * C, 7.0)
    ctx.softmax_grad(R, C, xg, s, g)                      # Jacobian x ones = 0 (tests/matrix/softmax.rs), SET
    assert float(ctx.max(ctx.unary(0, xg))) <= (4 * C * 2.0 ** -24) ** 2
    row = ctx.softmax(1, C, dx.view(5 * C, C)).numpy()    # one row against the oracle
    assert np.max(np.abs(row - O.softmax(1, C, x[5 * C:6 * C].copy()))) <= 4 * C * 2.0 ** -24 * row.max() + 1e-9


def _cce_chain_oracle(samples, features, z, y, labels, rows):
    """examples/nn.rs:190-233 tail replayed with the oracle's single ops (the reference's operation order)"""
    s = O.softmax(samples, features, z)
    preds = O.binary_ew(O.MUL, O.unary(O.UN_CLIP, s, 1e-7, 1. - 1e-7), y)
    loss = O.unary(O.UN_NEG_LN, O.sum_cols(features, preds))
    g = O.unary(O.UN_NEG_DIV_SCALAR, O.binary_ew(O.DIV, y, s), float(rows))
    dz = np.zeros_like(z)
    O.softmax_grad(samples, features, dz, s, g, closed=True)
    correct = int(np.sum(np.argmax(s.reshape(samples, features), axis=1) == labels))
    return s, dz, loss, correct


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("shape", [(1, 10), (7, 3), (1000, 10), (65536, 10), (333, 16), (100, 32), (5, 1)])
def test_softmax_cce_fused_equals_chain(be, dt, shape):
    """sl_softmax_cce (one launch) against the nine-launch chain it replaces — BIT-IDENTICAL for features <= 32 — and against
    the oracle's op-by-op replay (exact loss/accuracy structure, libm-level tolerance on values)"""
    import sliced_b200 as S
    ctx = be.ctx
    samples, features = shape
    rng = np.random.default_rng(samples + features)
    z = rnd(rng, samples * features, dt, -4, 4)
    labels = rng.integers(0, features, samples).astype(np.int32)
    y = np.zeros((samples, features), DT[dt]); y[np.arange(samples), labels] = 1; y = y.ravel()
    rows = 4 * samples   # global batch under data parallelism
    dz_, dy, dl = ctx.array(z), ctx.array(y), ctx.array(labels)
    probs, dz, loss, correct = ctx.softmax_cce(samples, features, dz_, dy, dl, rows)
    # the chain, through the same C ABI
    s = ctx.softmax(samples, features, dz_)
    t = ctx.binary_ew(S.MUL, ctx.unary(S.UN_CLIP, s, 1e-7, 1. - 1e-7), dy)
    loss_c = ctx.unary(S.UN_NEG_LN, ctx.sum_cols(features, t))
    g = ctx.unary(S.UN_NEG_DIV_SCALAR, ctx.binary_ew(S.DIV, dy, s), float(rows))
    dz_c = ctx.array(rnd(rng, samples * features, dt))
    ctx.softmax_grad(samples, features, dz_c, s, g)
    exact(probs.numpy(), s.numpy(), "probs")
    exact(loss.numpy(), loss_c.numpy(), "loss")
    exact(dz.numpy(), dz_c.numpy(), "dz")
    assert correct == ctx.count_correct(samples, features, s, dl)
    # the oracle
    so, dzo, lo, co = _cce_chain_oracle(samples, features, z, y, labels, rows)
    rel = 1e-6 if dt == "f32" else 1e-13
    close_rel(probs.numpy(), so, rel, "probs vs oracle")
    close_rel(loss.numpy(), lo, rel * 4, "loss vs oracle")
    assert np.max(np.abs(dz.numpy() - dzo)) <= rel * 8 * max(np.max(np.abs(dzo)), 1e-30)
    if dt == "f64":
        assert correct == co


@pytest.mark.parametrize("shape", [(64, 1000), (48, 16384), (5, 33)])
def test_softmax_cce_fused_wide_rows(be, shape):
    """features > 32: block-per-row form, K-scaled tolerance against the oracle's replay; soft (non one-hot) targets"""
    ctx = be.ctx
    samples, features = shape
    rng = np.random.default_rng(features)
    z = rnd(rng, samples * features, "f32", -3, 3)
    y = rng.uniform(0, 1, (samples, features)).astype(np.float32)
    y = (y / y.sum(1, keepdims=True)).astype(np.float32).ravel()
    labels = rng.integers(0, features, samples).astype(np.int32)
    probs, dz, loss, correct = ctx.softmax_cce(samples, features, ctx.array(z), ctx.array(y), ctx.array(labels), samples)
    so, dzo, lo, co = _cce_chain_oracle(samples, features, z, y, labels, samples)
    tol = 4 * features * EPS["f32"]
    assert np.max(np.abs(probs.numpy() - so)) <= tol * np.max(so)
    assert np.max(np.abs(loss.numpy() - lo)) <= tol * np.max(np.abs(lo)) + 1e-6
    assert np.max(np.abs(dz.numpy() - dzo)) <= tol * max(np.max(np.abs(dzo)), 1e-30) + 1e-9
    assert correct == int(np.sum(np.argmax(probs.numpy().reshape(samples, features), axis=1) == labels))
