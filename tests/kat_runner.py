"""
Replays one known-answer entry of tests/golden/reference_kats.py on a backend.

A backend is an object with the oracle's function vocabulary (oracle/__init__.py); the CUDA path is adapted to the same
vocabulary by tests/cuda_backend.py so that both are checked against the very same expectations.
"""
import numpy as np

DT = {"f32": np.float32, "f64": np.float64, "i32": np.int32}


def _a(v, dt):
    return np.ascontiguousarray(np.asarray(v, dtype=dt))


def _check(got, want, dt, tol, tight=None, what=""):
    got = np.asarray(got)
    want = np.asarray(want, dtype=np.float64 if np.issubdtype(dt, np.floating) else dt)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if tol == 0:
        assert np.array_equal(got.astype(want.dtype), want), f"{what}: exact mismatch\n got  {got}\n want {want}"
    else:
        err = np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) if got.size else 0.0
        assert err <= tol, f"{what}: abs err {err} > {tol}"
        if tight is not None:  # the reference's tolerance is loose (1e-2); we also hold a tight one
            assert err <= tight, f"{what}: abs err {err} > tight {tight}"


def run_kat(be, k):
    dt = DT[k["dtype"]]
    kind = k["kind"]
    exp = k.get("expect", {})
    tol = k.get("tol", 0)
    tight = k.get("tight")
    ck = lambda got, key: _check(got, exp[key], dt, tol, tight, f'{k["id"]}.{key} ({k["src"]})')

    if kind == "binary_ew":
        ck(be.binary_ew(k["op"], _a(k["lhs"], dt), _a(k["rhs"], dt)), "out")
    elif kind == "binary_ew_grad":
        lhs, rhs, og = _a(k["lhs"], dt), _a(k["rhs"], dt), _a(k["out_grad"], dt)
        lg, rg = np.zeros_like(lhs), np.zeros_like(rhs)
        lg, rg = be.binary_ew_grad(k["op"], lhs, rhs, lg, rg, og) or (lg, rg)
        ck(lg, "lhs_grad"); ck(rg, "rhs_grad")
    elif kind == "unary":
        ck(be.unary(k["op"], _a(k["x"], dt), k.get("p0", 0.0), k.get("p1", 0.0)), "out")
    elif kind == "unary_grad":
        x, og = _a(k["x"], dt), _a(k["out_grad"], dt)
        xg = np.zeros_like(x)
        xg = be.unary_grad(k["op"], x, xg, og, k.get("p0", 0.0), k.get("p1", 0.0))
        ck(xg if xg is not None else None, "x_grad") if xg is not None else None
    elif kind == "row_op":
        ck(be.row_op(k["op"], k["cols"], _a(k["lhs"], dt), _a(k["rhs"], dt)), "out")
    elif kind == "add_row_mut":
        lhs = _a(k["lhs"], dt)
        lhs = be.add_row_mut(k["rows"], k["cols"], lhs, _a(k["rhs"], dt))
        ck(lhs, "lhs")
    elif kind == "add_row_grad":
        og = _a(k["out_grad"], dt)
        lg, rg = np.zeros(k["rows"] * k["cols"], dt), np.zeros(k["cols"], dt)
        lg, rg = be.add_row_grad(k["rows"], k["cols"], lg, rg, og)
        ck(lg, "lhs_grad"); ck(rg, "rhs_grad")
    elif kind == "add_row_mut_grad":
        og = _a(k["out_grad"], dt)
        rg = be.add_row_mut_grad(k["rows"], k["cols"], np.zeros(k["cols"], dt), og)
        ck(rg, "rhs_grad")
    elif kind == "row_op_grad":
        lhs, rhs, og = _a(k["lhs"], dt), _a(k["rhs"], dt), _a(k["out_grad"], dt)
        lg, rg = be.row_op_grad(k["op"], k["cols"], lhs, rhs, np.zeros_like(lhs), np.zeros_like(rhs), og)
        ck(lg, "lhs_grad"); ck(rg, "rhs_grad")
    elif kind == "col_op":
        ck(be.col_op(k["op"], k["cols"], _a(k["lhs"], dt), _a(k["rhs"], dt)), "out")
    elif kind == "col_op_grad":
        lhs, rhs, og = _a(k["lhs"], dt), _a(k["rhs"], dt), _a(k["out_grad"], dt)
        lg, rg = be.col_op_grad(k["op"], k["cols"], lhs, rhs, np.zeros_like(lhs), np.zeros_like(rhs), og)
        ck(lg, "lhs_grad"); ck(rg, "rhs_grad")
    elif kind == "max_rows_noinit":
        if not hasattr(be, "max_rows_noinit"):
            return "skipped: slice-only entry"
        out = _a(k["init"], dt)
        be.max_rows_noinit(k["cols"], _a(k["x"], dt), out)
        ck(out, "out")
    elif kind in ("max_rows", "max_cols", "sum_rows", "sum_cols", "mean_rows", "mean_cols"):
        ck(getattr(be, kind)(k["cols"], _a(k["x"], dt)), "out")
    elif kind in ("max_cols_grad", "max_rows_grad"):
        x, out, og = _a(k["x"], dt), _a(k["out"], dt), _a(k["out_grad"], dt)
        xg = getattr(be, kind)(k["cols"], out, x, np.zeros_like(x), og)
        ck(xg, "x_grad")
    elif kind == "max_grad":
        if not hasattr(be, "max_grad"):
            return "skipped: slice-only entry"
        x = _a(k["x"], dt)
        xg = np.zeros_like(x)
        be.max_grad(k["out"], x, xg)
        ck(xg, "x_grad")
    elif kind in ("sum_cols_grad", "sum_rows_grad", "mean_rows_grad", "mean_cols_grad"):
        og = _a(k["out_grad"], dt)
        xg = getattr(be, kind)(k["cols"], np.zeros(k["rows"] * k["cols"], dt), og)
        ck(xg, "x_grad")
    elif kind == "transpose":
        ck(be.transpose(k["rows"], k["cols"], _a(k["x"], dt)), "out")
    elif kind == "softmax":
        ck(be.softmax(k["samples"], k["features"], _a(k["x"], dt)), "out")
    elif kind == "softmax_grad":
        x, og = _a(k["x"], dt), _a(k["out_grad"], dt)
        out = be.softmax(k["samples"], k["features"], x)
        xg = be.softmax_grad(k["samples"], k["features"], np.zeros_like(x), out, og)
        ck(xg, "x_grad")
    elif kind == "diagflat":
        ck(be.diagflat(_a(k["x"], dt)), "out")
    elif kind == "diagflat_grad":
        og = _a(k["out_grad"], dt)
        xg = be.diagflat_grad(np.zeros(k["n"], dt), og)
        ck(xg, "x_grad")
    elif kind == "onehot":
        ck(be.onehot(_a(k["classes"], dt)), "out")
    elif kind == "onehot_grad":
        cl, og = _a(k["classes"], dt), _a(k["out_grad"], dt)
        cg = be.onehot_grad(k["highest_class"], cl, np.zeros_like(cl), og)
        ck(cg, "classes_grad")
    elif kind == "gemm":
        ck(be.gemm(k["m"], k["k"], k["n"], _a(k["lhs"], dt), _a(k["rhs"], dt)), "out")
    elif kind == "gemm_grad":
        lhs, rhs, og = _a(k["lhs"], dt), _a(k["rhs"], dt), _a(k["out_grad"], dt)
        lg, rg = be.gemm_grad(k["m"], k["k"], k["n"], lhs, rhs, np.zeros_like(lhs), np.zeros_like(rhs), og)
        ck(lg, "lhs_grad"); ck(rg, "rhs_grad")
    elif kind == "tgemm_equiv":
        # Tgemm(m,n,k,a,b) with a stored [k x m]  ==  gemm(m,n,k, transpose(a), b)
        a, b = _a(k["a"], dt), _a(k["b"], dt)
        m, n, kk = k["m"], k["n"], k["k"]
        got = be.blas_Tgemm(m, n, kk, a, b)
        ta = be.transpose(kk, m, a)
        want = be.blas_gemm(m, n, kk, ta, b)
        _check(got, want, dt, tol, None, k["id"])
        _check(got, (a.reshape(kk, m).T.astype(np.float64) @ b.reshape(kk, n).astype(np.float64)).ravel(), dt, 1e-9, None, k["id"])
    elif kind == "gemmt_equiv":
        a, b = _a(k["a"], dt), _a(k["b"], dt)
        m, n, kk = k["m"], k["n"], k["k"]
        got = be.blas_gemmT(m, n, kk, a, b)
        tb = be.transpose(n, kk, b)
        want = be.blas_gemm(m, n, kk, a, tb)
        _check(got, want, dt, tol, None, k["id"])
        _check(got, (a.reshape(m, kk).astype(np.float64) @ b.reshape(n, kk).T.astype(np.float64)).ravel(), dt, 1e-9, None, k["id"])
    elif kind == "chained":
        out = be.chained_fwd(_a(k["x"], dt), _a(k["b"], dt))
        bits = out.view(np.uint32)
        assert np.all(bits == exp["out_bits"]), f'{k["id"]}: {hex(int(bits[0]))} != {hex(exp["out_bits"])}'
        assert out[0] == np.float32(9.336999)  # examples/chained_perf.rs:91
    else:
        raise AssertionError(f"unknown KAT kind {kind}")
    return None
