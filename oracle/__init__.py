"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

numpy front-end of the CPU oracle (oracle/sliced_oracle.c), the restatement of the reference's custos-CPU
implementation of sliced's forward + backward ops.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this package; nothing under sliced_b200/ does.

The reference (Rust, un-vendored custos dependency) cannot be built here, so this oracle is pinned by the
reference's own golden vectors: tests/test_oracle_golden.py.

Function names follow the reference's slice functions (src/ops2/<op>/cpu.rs); semantics (SET vs ACC, loop
order, tie handling) are documented per function in sliced_oracle_typed.inc with file:line citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsliced_oracle.so")

ADD, SUB, MUL, DIV = 0, 1, 2, 3
(UN_SQUARE, UN_POW, UN_RELU, UN_TANH, UN_SIGMOID, UN_EXP, UN_LN, UN_NEG_LN, UN_CLIP, UN_NEG, UN_MUL_SCALAR,
 UN_NEG_DIV_SCALAR, UN_ADD_SCALAR) = range(13)

_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.int32): "i32"}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ("sliced_oracle.c", "sliced_oracle_typed.inc")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _chk(*arrs):
    dt = None
    for a in arrs:
        if a is None:
            continue
        assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "oracle wants contiguous numpy arrays"
        dt = dt or a.dtype
        assert a.dtype == dt, (a.dtype, dt)
    return _SUF[np.dtype(dt)]


def _fn(name, suf, restype=None):
    f = getattr(lib(), f"{name}_{suf}")
    f.restype = restype
    return f


def arr(x, dtype):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


_sz = C.c_size_t
_d = C.c_double
_i = C.c_int

# ---------------------------------------------------------------- binary / unary


def binary_ew(op, lhs, rhs):
    s = _chk(lhs, rhs)
    out = np.empty_like(lhs)
    _fn("orc_binary_ew", s)(_i(op), _p(lhs), _p(rhs), _p(out), _sz(lhs.size))
    return out


def binary_ew_grad(op, lhs, rhs, lhs_grad, rhs_grad, out_grad):
    s = _chk(lhs, rhs, lhs_grad, rhs_grad, out_grad)
    n = min(lhs.size, rhs.size, out_grad.size)
    _fn("orc_binary_ew_grad", s)(_i(op), _p(lhs), _p(rhs), _p(lhs_grad), _p(rhs_grad), _p(out_grad), _sz(n))


def add_ew_grad(lhs_grad, rhs_grad, out_grad):
    s = _chk(lhs_grad, rhs_grad, out_grad)
    _fn("orc_add_ew_grad", s)(_p(lhs_grad), _p(rhs_grad), _p(out_grad), _sz(out_grad.size))


def unary(op, x, p0=0.0, p1=0.0):
    s = _chk(x)
    out = np.empty_like(x)
    _fn("orc_unary", s)(_i(op), _d(p0), _d(p1), _p(x), _p(out), _sz(x.size))
    return out


def unary_grad(op, x, x_grad, out_grad, p0=0.0, p1=0.0):
    s = _chk(x, x_grad, out_grad)
    _fn("orc_unary_grad", s)(_i(op), _d(p0), _d(p1), _p(x), _p(x_grad), _p(out_grad), _sz(x.size))

def unary_deriv(op, x, p0=0.0, p1=0.0):
    s = _chk(x)
    out = np.empty_like(x)
    _fn("orc_unary_deriv", s)(_i(op), _d(p0), _d(p1), _p(x), _p(out), _sz(x.size))
    return out


def chain_replay(listing, inputs, outputs, scale=None):
    """Op-by-op replay of a fused-chain program (sliced_b200.chain.Prog.listing()) with the oracle's single-op loops: every
    instruction is one full pass over the arrays, exactly the launch-per-op sequence the fused kernel replaces.  outputs: arrays
    updated in place (SET, or ACC when the program says so).  scale (optional float64 array, updated in place): element-wise largest
    magnitude any register held — the yardstick for comparing programs that contain libm calls."""
    CH_RDIV_IMM, CH_CONST, CH_COPY, CH_UNARY_F, CH_UNARY_D = 4, 5, 6, 16, 48
    dt = inputs[0].dtype if inputs else outputs[0].dtype
    n = outputs[0].size
    regs = [None] * listing["n_regs"]
    for r in range(listing["n_in"]):
        regs[r] = inputs[r].copy()
    for (op, dst, a, b, p0, p1) in listing["instr"]:
        if op <= 3:
            v = binary_ew(op, regs[a], regs[b])
        elif op == CH_RDIV_IMM:
            v = binary_ew(DIV, np.full(n, p0, dt), regs[a])
        elif op == CH_CONST:
            v = np.full(n, p0, dt)
        elif op == CH_COPY:
            v = regs[a].copy()
        elif op >= CH_UNARY_D:
            v = unary_deriv(op - CH_UNARY_D, regs[a], p0, p1)
        else:
            v = unary(op - CH_UNARY_F, regs[a], p0, p1)
        regs[dst] = v
        if scale is not None:
            np.maximum(scale, np.abs(v.astype(np.float64)), out=scale)
    for j, (r, acc) in enumerate(zip(listing["out_reg"], listing["out_acc"])):
        if acc:
            outputs[j][:] = binary_ew(ADD, outputs[j].copy(), regs[r])
        else:
            outputs[j][:] = regs[r]
    return outputs



# ---------------------------------------------------------------- row_op / col_op


def row_op(op, cols, lhs, rhs):
    s = _chk(lhs, rhs)
    out = np.empty_like(lhs)
    _fn("orc_row_op", s)(_i(op), _sz(lhs.size // cols), _sz(cols), _p(lhs), _p(rhs), _p(out))
    return out


def add_row(cols, lhs, rhs):
    return row_op(ADD, cols, lhs, rhs)


def add_row_mut(rows, cols, lhs, rhs):
    s = _chk(lhs, rhs)
    _fn("orc_add_row_mut", s)(_sz(rows), _sz(cols), _p(lhs), _p(rhs))


def add_row_grad(rows, cols, lhs_grad, rhs_grad, out_grad):
    s = _chk(lhs_grad, rhs_grad, out_grad)
    _fn("orc_add_row_grad", s)(_sz(rows), _sz(cols), _p(lhs_grad), _p(rhs_grad), _p(out_grad))


def add_row_mut_grad(rows, cols, rhs_grad, out_grad):
    s = _chk(rhs_grad, out_grad)
    _fn("orc_add_row_mut_grad", s)(_sz(rows), _sz(cols), _p(rhs_grad), _p(out_grad))


def row_op_grad(op, cols, lhs, rhs, lhs_grad, rhs_grad, out_grad):
    s = _chk(lhs, rhs, lhs_grad, rhs_grad, out_grad)
    _fn("orc_row_op_grad", s)(_i(op), _sz(lhs.size // cols), _sz(cols), _p(lhs), _p(rhs), _p(lhs_grad), _p(rhs_grad),
                              _p(out_grad))


def col_op(op, cols, lhs, rhs):
    s = _chk(lhs, rhs)
    out = np.empty_like(lhs)
    _fn("orc_col_op", s)(_i(op), _sz(lhs.size // cols), _sz(cols), _p(lhs), _p(rhs), _p(out))
    return out


def col_op_grad(op, cols, lhs, rhs, lhs_grad, rhs_grad, out_grad):
    s = _chk(lhs, rhs, lhs_grad, rhs_grad, out_grad)
    _fn("orc_col_op_grad", s)(_i(op), _sz(lhs.size // cols), _sz(cols), _p(lhs), _p(rhs), _p(lhs_grad), _p(rhs_grad),
                              _p(out_grad))


# ---------------------------------------------------------------- reductions

_CT = {"f32": C.c_float, "f64": C.c_double, "i32": C.c_int32}


def sum_(x):
    s = _chk(x)
    return x.dtype.type(_fn("orc_sum", s, _CT[s])(_p(x), _sz(x.size)))


def mean(x):
    s = _chk(x)
    return x.dtype.type(_fn("orc_mean", s, _CT[s])(_p(x), _sz(x.size)))


def max_(x):
    s = _chk(x)
    return x.dtype.type(_fn("orc_max", s, _CT[s])(_p(x), _sz(x.size)))


def _reduce(name, rows, cols, x, out_len, zero=False):
    s = _chk(x)
    out = np.zeros(out_len, dtype=x.dtype) if zero else np.empty(out_len, dtype=x.dtype)
    _fn(name, s)(_sz(rows), _sz(cols), _p(x), _p(out))
    return out


def sum_rows(cols, x):
    return _reduce("orc_sum_rows", x.size // cols, cols, x, cols, zero=True)


def sum_cols(cols, x):
    return _reduce("orc_sum_cols", x.size // cols, cols, x, x.size // cols)


def mean_rows(cols, x):
    return _reduce("orc_mean_rows", x.size // cols, cols, x, cols, zero=True)


def mean_cols(cols, x):
    return _reduce("orc_mean_cols", x.size // cols, cols, x, x.size // cols)


def max_rows(cols, x):
    return _reduce("orc_max_rows", x.size // cols, cols, x, cols)


def max_rows_noinit(cols, x, out):
    s = _chk(x, out)
    _fn("orc_max_rows_noinit", s)(_sz(x.size // cols), _sz(cols), _p(x), _p(out))


def max_cols(cols, x):
    return _reduce("orc_max_cols", x.size // cols, cols, x, x.size // cols)


def _rgrad(name, cols, x_grad, out_grad):
    s = _chk(x_grad, out_grad)
    _fn(name, s)(_sz(x_grad.size // cols), _sz(cols), _p(x_grad), _p(out_grad))


def sum_rows_grad(cols, x_grad, out_grad):
    _rgrad("orc_sum_rows_grad", cols, x_grad, out_grad)


def sum_cols_grad(cols, x_grad, out_grad):
    _rgrad("orc_sum_cols_grad", cols, x_grad, out_grad)


def mean_rows_grad(cols, x_grad, out_grad):
    _rgrad("orc_mean_rows_grad", cols, x_grad, out_grad)


def mean_cols_grad(cols, x_grad, out_grad):
    _rgrad("orc_mean_cols_grad", cols, x_grad, out_grad)


def max_rows_grad(cols, out, x, x_grad, out_grad):
    s = _chk(out, x, x_grad, out_grad)
    _fn("orc_max_rows_grad", s)(_sz(x.size // cols), _sz(cols), _p(out), _p(x), _p(x_grad), _p(out_grad))


def max_cols_grad(cols, out, x, x_grad, out_grad):
    s = _chk(out, x, x_grad, out_grad)
    rc = _fn("orc_max_cols_grad", s, C.c_int)(_sz(x.size // cols), _sz(cols), _p(out), _p(x), _p(x_grad), _p(out_grad))
    if rc != 0:
        raise RuntimeError("Could not find maximum in gradient calculation")  # max/grad/cpu.rs:66


def max_grad(out, x, x_grad):
    s = _chk(x, x_grad)
    rc = _fn("orc_max_grad", s, C.c_int)(_CT[s](out), _p(x), _p(x_grad), _sz(x.size))
    if rc != 0:
        raise RuntimeError("max not found")


# ---------------------------------------------------------------- transpose / diagflat / onehot


def transpose(rows, cols, x, out=None, assign=False, quirk=True):
    s = _chk(x, out)
    if out is None:
        out = np.zeros_like(x)
    _fn("orc_transpose", s)(_sz(rows), _sz(cols), _p(x), _p(out), _i(int(assign)), _i(int(quirk)))
    return out


def diagflat(x):
    s = _chk(x)
    out = np.zeros(x.size * x.size, dtype=x.dtype)
    _fn("orc_diagflat", s)(_p(x), _sz(x.size), _p(out))
    return out


def diagflat_grad(x_grad, out_grad):
    s = _chk(x_grad, out_grad)
    _fn("orc_diagflat_grad", s)(_p(x_grad), _sz(x_grad.size), _p(out_grad))


def onehot(classes):
    s = _chk(classes)
    hc = int(max_(classes)) + 1  # onehot/cpu.rs:8
    out = np.zeros(classes.size * hc, dtype=classes.dtype)
    _fn("orc_onehot", s)(_sz(hc), _p(classes), _sz(classes.size), _p(out))
    return out


def onehot_grad(highest_class, classes, classes_grad, out_grad):
    s = _chk(classes, classes_grad, out_grad)
    _fn("orc_onehot_grad", s)(_sz(highest_class), _p(classes), _sz(classes.size), _p(classes_grad), _p(out_grad))


# ---------------------------------------------------------------- gemm / softmax


def gemm_ex(trans_a, trans_b, m, n, k, a, b, c=None, accumulate=False):
    s = _chk(a, b, c)
    if c is None:
        c = np.zeros(m * n, dtype=a.dtype)
    _fn("orc_gemm_ex", s)(_i(int(trans_a)), _i(int(trans_b)), _sz(m), _sz(n), _sz(k), _p(a), _p(b), _p(c),
                          _i(int(accumulate)))
    return c


def gemm(m, k, n, lhs, rhs):
    """Gemm::gemm(m, k, n, lhs, rhs) — src/ops2/gemm/mod.rs:24-31."""
    return gemm_ex(False, False, m, n, k, lhs, rhs)


def blas_gemm(m, n, k, a, b, c=None):
    """custos GenericBlas::gemm(m, n, k, a, b, c)."""
    return gemm_ex(False, False, m, n, k, a, b, c)


def blas_gemmT(m, n, k, a, b, c=None):
    return gemm_ex(False, True, m, n, k, a, b, c)


def blas_Tgemm(m, n, k, a, b, c=None):
    return gemm_ex(True, False, m, n, k, a, b, c)


def gemm_grad(m, k, n, lhs, rhs, lhs_grad, rhs_grad, out_grad, accumulate=False):
    s = _chk(lhs, rhs, lhs_grad, rhs_grad, out_grad)
    _fn("orc_gemm_grad", s)(_sz(m), _sz(k), _sz(n), _p(lhs), _p(rhs), _p(lhs_grad), _p(rhs_grad), _p(out_grad),
                            _i(int(accumulate)))


def gemm_truth(trans_a, trans_b, m, n, k, a, b):
    out = np.zeros(m * n, dtype=np.float64)
    f = lib().orc_gemm_truth_f32
    f.restype = None
    f(_i(int(trans_a)), _i(int(trans_b)), _sz(m), _sz(n), _sz(k), _p(a), _p(b), _p(out))
    return out


def softmax(samples, features, x):
    s = _chk(x)
    out = np.empty_like(x)
    _fn("orc_softmax", s)(_sz(samples), _sz(features), _p(x), _p(out))
    return out


def softmax_grad(samples, features, x_grad, out, out_grad, closed=False):
    s = _chk(x_grad, out, out_grad)
    name = "orc_softmax_grad_closed" if closed else "orc_softmax_grad"
    _fn(name, s)(_sz(samples), _sz(features), _p(x_grad), _p(out), _p(out_grad))


def sgd_step(w, g, lr):
    s = _chk(w, g)
    _fn("orc_sgd_step", s)(_p(w), _p(g), _d(lr), _sz(w.size))


def chained_fwd(x, b):
    s = _chk(x, b)
    out = np.empty_like(x)
    _fn("orc_chained_fwd", s)(_p(x), _p(b), _p(out), _sz(x.size))
    return out


def chained_bwd(x, b, x_grad, b_grad, out_grad):
    s = _chk(x, b, x_grad, b_grad, out_grad)
    _fn("orc_chained_bwd", s)(_p(x), _p(b), _p(x_grad), _p(b_grad), _p(out_grad), _sz(x.size))


# ---------------------------------------------------------------- BLAS hook + MLP step

_blas_keepalive = []


def use_openblas(threads: int | None = None) -> bool:
    """Plug OpenBLAS' cblas_sgemm (bundled in the scipy wheel) into the f32 gemm of the MLP replay — the reference
    links a system CBLAS under its `blas` feature (Cargo.toml:26).  Returns False when no OpenBLAS is found."""
    import glob
    cands = []
    try:
        import scipy
        cands += glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
    except Exception:
        pass
    for path in cands:
        try:
            bl = C.CDLL(path)
            fn = getattr(bl, "scipy_cblas_sgemm", None)
            if fn is None:
                continue
            if threads is not None and hasattr(bl, "scipy_openblas_set_num_threads"):
                bl.scipy_openblas_set_num_threads(int(threads))
            lib().orc_set_sgemm(C.cast(fn, C.c_void_p))
            _blas_keepalive.append(bl)
            return True
        except OSError:
            continue
    return False


def use_naive_gemm():
    lib().orc_set_sgemm(C.c_void_p(0))


def mlp_step(loss_kind, dims, x, y, labels, W, B, lr, grad_rows=None, apply_sgd=True, want_grads=False):
    """One training step of examples/nn.rs (loss_kind=0) / examples/sine_net.rs (loss_kind=1), replayed op by op.
    W, B: lists of float32 arrays, updated in place when apply_sgd.  Returns (loss_sum, correct, dW, dB)."""
    n_layers = len(dims) - 1
    batch = x.size // dims[0]
    grad_rows = batch if grad_rows is None else grad_rows
    PF = C.POINTER(C.c_float)
    Wp = (PF * n_layers)(*[w.ctypes.data_as(PF) for w in W])
    Bp = (PF * n_layers)(*[b.ctypes.data_as(PF) for b in B])
    dW = [np.zeros_like(w) for w in W] if want_grads else None
    dB = [np.zeros_like(b) for b in B] if want_grads else None
    dWp = (PF * n_layers)(*[w.ctypes.data_as(PF) for w in dW]) if want_grads else None
    dBp = (PF * n_layers)(*[b.ctypes.data_as(PF) for b in dB]) if want_grads else None
    dims_a = (C.c_size_t * len(dims))(*dims)
    loss = C.c_double(0)
    correct = C.c_int64(0)
    f = lib().orc_mlp_step_f32
    f.restype = C.c_int
    rc = f(_i(loss_kind), _i(n_layers), dims_a, _sz(batch), _p(x), _p(y), _p(labels), Wp, Bp, C.c_float(lr),
           _sz(grad_rows), _i(int(apply_sgd)), dWp, dBp, C.byref(loss), C.byref(correct))
    assert rc == 0
    return loss.value, correct.value, dW, dB
