"""
Raw device-op layer: a Context (one CUDA device + stream, `sl_ctx`) and DeviceArray (a typed device allocation),
with one method per L2 op trait of the reference (src/ops2/<op>/mod.rs, grad.rs).  Argument order follows the
reference traits.  Everything runs through the C ABI; there is no host fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import F32, F64, I32, SlicedError, check, load

_DT = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32}
_NP = {F32: np.float32, F64: np.float64, I32: np.int32}


def dtype_code(dt) -> int:
    try:
        return _DT[np.dtype(dt)]
    except KeyError:
        raise SlicedError(capi.SL_ERR_UNSUPPORTED, f"dtype {dt} unsupported (f32, f64, i32)")


class DeviceArray:
    """A flat, typed device buffer (the analogue of custos `Buffer<T, CUDA>` †, without autograd)."""

    def __init__(self, ctx: "Context", n: int, dtype, ptr: int | None = None, owner=None):
        self.ctx = ctx
        self.size = int(n)
        self.dtype = np.dtype(dtype)
        self.code = dtype_code(dtype)
        self._owned = ptr is None
        self._owner = owner
        if ptr is None:
            p = C.c_void_p()
            check(ctx.h, ctx.lib.sl_malloc(ctx.h, max(self.nbytes, 1), C.byref(p)))
            ptr = p.value
        self.ptr = ptr

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    def __len__(self):
        return self.size

    def numpy(self) -> np.ndarray:
        out = np.empty(self.size, dtype=self.dtype)
        if self.size:
            check(self.ctx.h, self.ctx.lib.sl_read(self.ctx.h, out.ctypes.data, self.ptr, self.nbytes))
        return out

    read = numpy

    def write(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype).ravel()
        assert host.size == self.size
        if self.size:
            check(self.ctx.h, self.ctx.lib.sl_write(self.ctx.h, self.ptr, host.ctypes.data, self.nbytes))
            self.ctx.sync()  # pageable source: make the call safe to return from
        return self

    def clear(self):
        check(self.ctx.h, self.ctx.lib.sl_clear(self.ctx.h, self.ptr, self.nbytes))
        return self

    def view(self, offset: int, n: int) -> "DeviceArray":
        """Sub-buffer (used by the parity tests to produce 16-byte-misaligned pointers)."""
        assert 0 <= offset and offset + n <= self.size
        return DeviceArray(self.ctx, n, self.dtype, ptr=self.ptr + offset * self.dtype.itemsize, owner=self)

    def free(self):
        if self._owned and self.ptr:
            self.ctx.lib.sl_free(self.ctx.h, self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


class Context:
    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load()
        h = C.c_void_p()
        if stream is None:
            rc = self.lib.sl_ctx_create(device, C.byref(h))
        else:
            rc = self.lib.sl_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h))
        check(None, rc)
        self.h = h

    # ---------------------------------------------------------------- plumbing
    def close(self):
        if self.h:
            self.lib.sl_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        check(self.h, self.lib.sl_sync(self.h))

    @property
    def stream(self) -> int:
        return self.lib.sl_ctx_stream(self.h) or 0

    @property
    def launches(self) -> int:
        return int(self.lib.sl_ctx_launch_count(self.h))

    def set_gemm_mode(self, mode: int):
        check(self.h, self.lib.sl_ctx_set_gemm_mode(self.h, mode))

    def empty(self, n, dtype=np.float32) -> DeviceArray:
        return DeviceArray(self, n, dtype)

    def zeros(self, n, dtype=np.float32) -> DeviceArray:
        return DeviceArray(self, n, dtype).clear()

    def array(self, host, dtype=None) -> DeviceArray:
        host = np.ascontiguousarray(host, dtype=dtype).ravel()
        return DeviceArray(self, host.size, host.dtype).write(host)

    def full(self, n, value, dtype=np.float32) -> DeviceArray:
        a = DeviceArray(self, n, dtype)
        check(self.h, self.lib.sl_fill(self.h, a.code, a.ptr, float(value), n))
        return a

    def copy(self, src: DeviceArray) -> DeviceArray:
        dst = DeviceArray(self, src.size, src.dtype)
        check(self.h, self.lib.sl_copy(self.h, dst.ptr, src.ptr, src.nbytes))
        return dst

    def _c(self, rc):
        check(self.h, rc)

    @staticmethod
    def _p(a):
        return None if a is None else a.ptr

    # ---------------------------------------------------------------- E
    def binary_ew(self, op, lhs, rhs, out=None):
        out = out or self.empty(lhs.size, lhs.dtype)
        self._c(self.lib.sl_binary_ew(self.h, lhs.code, op, lhs.ptr, rhs.ptr, out.ptr, lhs.size))
        return out

    def binary_ew_grad(self, op, lhs, rhs, lhs_grad, rhs_grad, out_grad):
        n = min(x.size for x in (lhs, rhs, out_grad) if x is not None)
        self._c(self.lib.sl_binary_ew_grad(self.h, out_grad.code, op, self._p(lhs), self._p(rhs), self._p(lhs_grad), self._p(rhs_grad),
                                           out_grad.ptr, n))

    def add_ew_grad(self, lhs_grad, rhs_grad, out_grad):
        self._c(self.lib.sl_add_ew_grad(self.h, out_grad.code, lhs_grad.ptr, rhs_grad.ptr, out_grad.ptr, out_grad.size))

    def unary(self, op, x, p0=0.0, p1=0.0, out=None):
        out = out or self.empty(x.size, x.dtype)
        self._c(self.lib.sl_unary(self.h, x.code, op, p0, p1, x.ptr, out.ptr, x.size))
        return out

    def unary_grad(self, op, x, x_grad, out_grad, p0=0.0, p1=0.0):
        self._c(self.lib.sl_unary_grad(self.h, x.code, op, p0, p1, x.ptr, x_grad.ptr, out_grad.ptr, x.size))

    def row_op(self, op, cols, lhs, rhs, out=None):
        out = out or self.empty(lhs.size, lhs.dtype)
        self._c(self.lib.sl_row_op(self.h, lhs.code, op, lhs.size // cols, cols, lhs.ptr, rhs.ptr, out.ptr))
        return out

    def add_row(self, cols, lhs, rhs):
        out = self.empty(lhs.size, lhs.dtype)
        self._c(self.lib.sl_add_row(self.h, lhs.code, lhs.size // cols, cols, lhs.ptr, rhs.ptr, out.ptr))
        return out

    def add_row_mut(self, rows, cols, lhs, rhs):
        self._c(self.lib.sl_add_row_mut(self.h, lhs.code, rows, cols, lhs.ptr, rhs.ptr))

    def add_row_grad(self, rows, cols, lhs_grad, rhs_grad, out_grad):
        self._c(self.lib.sl_add_row_grad(self.h, out_grad.code, rows, cols, lhs_grad.ptr, rhs_grad.ptr, out_grad.ptr))

    def add_row_mut_grad(self, rows, cols, rhs_grad, out_grad):
        self._c(self.lib.sl_add_row_mut_grad(self.h, out_grad.code, rows, cols, rhs_grad.ptr, out_grad.ptr))

    def row_op_grad(self, op, cols, lhs, rhs, lhs_grad, rhs_grad, out_grad):
        self._c(self.lib.sl_row_op_grad(self.h, out_grad.code, op, out_grad.size // cols, cols, self._p(lhs), self._p(rhs), self._p(lhs_grad),
                                        self._p(rhs_grad), out_grad.ptr))

    def col_op(self, op, cols, lhs, rhs, out=None):
        out = out or self.empty(lhs.size, lhs.dtype)
        self._c(self.lib.sl_col_op(self.h, lhs.code, op, lhs.size // cols, cols, lhs.ptr, rhs.ptr, out.ptr))
        return out

    def col_op_grad(self, op, cols, lhs, rhs, lhs_grad, rhs_grad, out_grad):
        self._c(self.lib.sl_col_op_grad(self.h, lhs.code, op, lhs.size // cols, cols, lhs.ptr, rhs.ptr, self._p(lhs_grad), self._p(rhs_grad),
                                        out_grad.ptr))

    def sgd_step(self, w, g, lr):
        self._c(self.lib.sl_sgd_step(self.h, w.code, w.ptr, g.ptr, lr, w.size))

    def chained_fwd(self, x, b, out=None):
        out = out or self.empty(x.size, x.dtype)
        self._c(self.lib.sl_chained_fwd(self.h, x.code, x.ptr, b.ptr, out.ptr, x.size))
        return out

    def chained_bwd(self, x, b, x_grad, b_grad, out_grad):
        self._c(self.lib.sl_chained_bwd(self.h, x.code, x.ptr, b.ptr, x_grad.ptr, b_grad.ptr, out_grad.ptr, x.size))

    # ---------------------------------------------------------------- G
    def gemm(self, m, k, n, lhs, rhs, out=None, mode=-1):
        out = out or self.empty(m * n, lhs.dtype)
        self._c(self.lib.sl_gemm(self.h, lhs.code, m, k, n, lhs.ptr, rhs.ptr, out.ptr, mode))
        return out

    def gemm_ex(self, trans_a, trans_b, m, n, k, a, b, c=None, accumulate=False, mode=-1):
        c = c or self.empty(m * n, a.dtype)
        self._c(self.lib.sl_gemm_ex(self.h, a.code, int(trans_a), int(trans_b), m, n, k, a.ptr, b.ptr, c.ptr, int(accumulate), mode))
        return c

    def gemm_nt(self, m, n, k, a, b, c=None, mode=-1):
        c = c or self.empty(m * n, a.dtype)
        self._c(self.lib.sl_gemm_nt(self.h, a.code, m, n, k, a.ptr, b.ptr, c.ptr, mode))
        return c

    def gemm_tn(self, m, n, k, a, b, c=None, mode=-1):
        c = c or self.empty(m * n, a.dtype)
        self._c(self.lib.sl_gemm_tn(self.h, a.code, m, n, k, a.ptr, b.ptr, c.ptr, mode))
        return c

    def gemm_grad(self, m, k, n, lhs, rhs, lhs_grad, rhs_grad, out_grad, accumulate=False, mode=-1):
        self._c(self.lib.sl_gemm_grad(self.h, out_grad.code, m, k, n, lhs.ptr, rhs.ptr, self._p(lhs_grad), self._p(rhs_grad), out_grad.ptr,
                                      int(accumulate), mode))

    def gemm_scope_begin(self):
        self._c(self.lib.sl_gemm_scope_begin(self.h))

    def gemm_scope_end(self):
        self._c(self.lib.sl_gemm_scope_end(self.h))

    # ---------------------------------------------------------------- R
    def _scalar(self, fn, x):
        out = self.empty(1, x.dtype)
        self._c(fn(self.h, x.code, x.ptr, x.size, out.ptr))
        return out.numpy()[0]

    def sum(self, x):
        return self._scalar(self.lib.sl_sum, x)

    def mean(self, x):
        return self._scalar(self.lib.sl_mean, x)

    def max(self, x):
        return self._scalar(self.lib.sl_max, x)

    def sum_rows(self, cols, x):
        out = self.empty(cols, x.dtype)
        self._c(self.lib.sl_sum_rows(self.h, x.code, x.size // cols, cols, x.ptr, out.ptr))
        return out

    def sum_cols(self, cols, x):
        out = self.empty(x.size // cols, x.dtype)
        self._c(self.lib.sl_sum_cols(self.h, x.code, x.size // cols, cols, x.ptr, out.ptr))
        return out

    def mean_rows(self, cols, x):
        out = self.empty(cols, x.dtype)
        self._c(self.lib.sl_mean_rows(self.h, x.code, x.size // cols, cols, x.ptr, out.ptr))
        return out

    def mean_cols(self, cols, x):
        out = self.empty(x.size // cols, x.dtype)
        self._c(self.lib.sl_mean_cols(self.h, x.code, x.size // cols, cols, x.ptr, out.ptr))
        return out

    def max_rows(self, cols, x, with_idx=False):
        out = self.empty(cols, x.dtype)
        idx = self.empty(cols, np.int32) if with_idx else None
        self._c(self.lib.sl_max_rows(self.h, x.code, x.size // cols, cols, x.ptr, out.ptr, self._p(idx)))
        return (out, idx) if with_idx else out

    def max_cols(self, rows, cols, x, with_idx=False):
        out = self.empty(rows, x.dtype)
        idx = self.empty(rows, np.int32) if with_idx else None
        self._c(self.lib.sl_max_cols(self.h, x.code, rows, cols, x.ptr, out.ptr, self._p(idx)))
        return (out, idx) if with_idx else out

    def sum_rows_grad(self, cols, x_grad, out_grad):
        self._c(self.lib.sl_sum_rows_grad(self.h, x_grad.code, x_grad.size // cols, cols, x_grad.ptr, out_grad.ptr))

    def sum_cols_grad(self, cols, x_grad, out_grad):
        self._c(self.lib.sl_sum_cols_grad(self.h, x_grad.code, x_grad.size // cols, cols, x_grad.ptr, out_grad.ptr))

    def mean_rows_grad(self, cols, x_grad, out_grad):
        self._c(self.lib.sl_mean_rows_grad(self.h, x_grad.code, x_grad.size // cols, cols, x_grad.ptr, out_grad.ptr))

    def mean_cols_grad(self, cols, x_grad, out_grad):
        self._c(self.lib.sl_mean_cols_grad(self.h, x_grad.code, x_grad.size // cols, cols, x_grad.ptr, out_grad.ptr))

    def max_rows_grad(self, cols, out, x, x_grad, out_grad):
        self._c(self.lib.sl_max_rows_grad(self.h, x.code, x.size // cols, cols, out.ptr, x.ptr, x_grad.ptr, out_grad.ptr))

    def max_cols_grad(self, cols, out, x, x_grad, out_grad):
        self._c(self.lib.sl_max_cols_grad(self.h, x.code, x.size // cols, cols, out.ptr, x.ptr, x_grad.ptr, out_grad.ptr))

    def max_cols_grad_idx(self, cols, idx, x_grad, out_grad):
        self._c(self.lib.sl_max_cols_grad_idx(self.h, x_grad.code, x_grad.size // cols, cols, idx.ptr, x_grad.ptr, out_grad.ptr))

    # ---------------------------------------------------------------- T / S / misc
    def transpose(self, rows, cols, x, out=None, accumulate=False):
        out = out or self.empty(x.size, x.dtype)
        self._c(self.lib.sl_transpose(self.h, x.code, rows, cols, x.ptr, out.ptr, int(accumulate)))
        return out

    def softmax(self, samples, features, x, out=None):
        out = out or self.empty(x.size, x.dtype)
        self._c(self.lib.sl_softmax(self.h, x.code, samples, features, x.ptr, out.ptr))
        return out

    def softmax_grad(self, samples, features, x_grad, out, out_grad):
        self._c(self.lib.sl_softmax_grad(self.h, out.code, samples, features, x_grad.ptr, out.ptr, out_grad.ptr))

    def softmax_cce(self, samples, features, logits, targets, labels=None, grad_rows=None):
        """fused softmax + cce + cce_grad + softmax_grad (+ accuracy): returns (probs, logits_grad, loss_per_sample, correct)"""
        probs, dz = self.empty(logits.size, logits.dtype), self.empty(logits.size, logits.dtype)
        loss = self.empty(samples, logits.dtype)
        cnt = self.zeros(1, np.int32)
        self._c(self.lib.sl_softmax_cce(self.h, logits.code, samples, features, logits.ptr, targets.ptr, self._p(labels), grad_rows or samples,
                                        probs.ptr, dz.ptr, loss.ptr, cnt.ptr))
        return probs, dz, loss, int(cnt.numpy()[0])

    def diagflat(self, x):
        out = self.zeros(x.size * x.size, x.dtype)
        self._c(self.lib.sl_diagflat(self.h, x.code, x.size, x.ptr, out.ptr))
        return out

    def diagflat_grad(self, x_grad, out_grad):
        self._c(self.lib.sl_diagflat_grad(self.h, x_grad.code, x_grad.size, x_grad.ptr, out_grad.ptr))

    def onehot(self, classes):
        hc = int(self.max(classes)) + 1  # src/ops2/onehot/cpu.rs:8
        out = self.zeros(classes.size * hc, classes.dtype)
        self._c(self.lib.sl_onehot(self.h, classes.code, classes.size, hc, classes.ptr, out.ptr))
        return out

    def onehot_grad(self, highest_class, classes, classes_grad, out_grad):
        self._c(self.lib.sl_onehot_grad(self.h, classes.code, classes.size, highest_class, classes.ptr, classes_grad.ptr, out_grad.ptr))

    def count_correct(self, rows, cols, preds, labels):
        cnt = self.empty(1, np.int32)
        self._c(self.lib.sl_count_correct(self.h, preds.code, rows, cols, preds.ptr, labels.ptr, cnt.ptr))
        return int(cnt.numpy()[0])
