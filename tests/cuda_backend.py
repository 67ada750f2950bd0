"""Adapter: the CUDA path (through the C ABI) in the vocabulary tests/kat_runner.py and the parity tests use:
numpy in, numpy out; in-place (ACC) arguments are uploaded, run on the device and returned."""
import numpy as np

import sliced_b200 as S


class CudaBackend:
    name = "cuda"

    def __init__(self, ctx=None, gemm_mode=-1):
        self.ctx = ctx or S.Context(0)
        self.gemm_mode = gemm_mode

    def _up(self, a):
        return None if a is None else self.ctx.array(a)

    def binary_ew(self, op, lhs, rhs): return self.ctx.binary_ew(op, self._up(lhs), self._up(rhs)).numpy()
    def binary_ew_grad(self, op, lhs, rhs, lg, rg, og):
        dl, dr = self._up(lg), self._up(rg)
        self.ctx.binary_ew_grad(op, self._up(lhs), self._up(rhs), dl, dr, self._up(og))
        return (None if dl is None else dl.numpy()), (None if dr is None else dr.numpy())
    def add_ew_grad(self, lg, rg, og):
        dl, dr = self._up(lg), self._up(rg)
        self.ctx.add_ew_grad(dl, dr, self._up(og)); return dl.numpy(), dr.numpy()
    def unary(self, op, x, p0=0.0, p1=0.0): return self.ctx.unary(op, self._up(x), p0, p1).numpy()
    def unary_grad(self, op, x, xg, og, p0=0.0, p1=0.0):
        d = self._up(xg); self.ctx.unary_grad(op, self._up(x), d, self._up(og), p0, p1); return d.numpy()
    def row_op(self, op, cols, lhs, rhs): return self.ctx.row_op(op, cols, self._up(lhs), self._up(rhs)).numpy()
    def add_row_mut(self, rows, cols, lhs, rhs):
        d = self._up(lhs); self.ctx.add_row_mut(rows, cols, d, self._up(rhs)); return d.numpy()
    def add_row_grad(self, rows, cols, lg, rg, og):
        dl, dr = self._up(lg), self._up(rg)
        self.ctx.add_row_grad(rows, cols, dl, dr, self._up(og)); return dl.numpy(), dr.numpy()
    def add_row_mut_grad(self, rows, cols, rg, og):
        dr = self._up(rg); self.ctx.add_row_mut_grad(rows, cols, dr, self._up(og)); return dr.numpy()
    def row_op_grad(self, op, cols, lhs, rhs, lg, rg, og):
        dl, dr = self._up(lg), self._up(rg)
        self.ctx.row_op_grad(op, cols, self._up(lhs), self._up(rhs), dl, dr, self._up(og)); return dl.numpy(), dr.numpy()
    def col_op(self, op, cols, lhs, rhs): return self.ctx.col_op(op, cols, self._up(lhs), self._up(rhs)).numpy()
    def col_op_grad(self, op, cols, lhs, rhs, lg, rg, og):
        dl, dr = self._up(lg), self._up(rg)
        self.ctx.col_op_grad(op, cols, self._up(lhs), self._up(rhs), dl, dr, self._up(og)); return dl.numpy(), dr.numpy()
    def sum(self, x): return self.ctx.sum(self._up(x))
    def mean(self, x): return self.ctx.mean(self._up(x))
    def max(self, x): return self.ctx.max(self._up(x))
    def sum_rows(self, cols, x): return self.ctx.sum_rows(cols, self._up(x)).numpy()
    def sum_cols(self, cols, x): return self.ctx.sum_cols(cols, self._up(x)).numpy()
    def mean_rows(self, cols, x): return self.ctx.mean_rows(cols, self._up(x)).numpy()
    def mean_cols(self, cols, x): return self.ctx.mean_cols(cols, self._up(x)).numpy()
    def max_rows(self, cols, x): return self.ctx.max_rows(cols, self._up(x)).numpy()
    def max_cols(self, cols, x): return self.ctx.max_cols(x.size // cols, cols, self._up(x)).numpy()
    def _xg(self, fn, cols, xg, og):
        d = self._up(xg); fn(cols, d, self._up(og)); return d.numpy()
    def sum_rows_grad(self, cols, xg, og): return self._xg(self.ctx.sum_rows_grad, cols, xg, og)
    def sum_cols_grad(self, cols, xg, og): return self._xg(self.ctx.sum_cols_grad, cols, xg, og)
    def mean_rows_grad(self, cols, xg, og): return self._xg(self.ctx.mean_rows_grad, cols, xg, og)
    def mean_cols_grad(self, cols, xg, og): return self._xg(self.ctx.mean_cols_grad, cols, xg, og)
    def max_rows_grad(self, cols, out, x, xg, og):
        d = self._up(xg); self.ctx.max_rows_grad(cols, self._up(out), self._up(x), d, self._up(og)); return d.numpy()
    def max_cols_grad(self, cols, out, x, xg, og):
        d = self._up(xg); self.ctx.max_cols_grad(cols, self._up(out), self._up(x), d, self._up(og)); return d.numpy()
    def transpose(self, rows, cols, x, out=None, accumulate=False):
        d = self._up(out) if out is not None else None
        return self.ctx.transpose(rows, cols, self._up(x), d, accumulate).numpy()
    def softmax(self, samples, features, x): return self.ctx.softmax(samples, features, self._up(x)).numpy()
    def softmax_grad(self, samples, features, xg, out, og):
        d = self._up(xg); self.ctx.softmax_grad(samples, features, d, self._up(out), self._up(og)); return d.numpy()
    def diagflat(self, x): return self.ctx.diagflat(self._up(x)).numpy()
    def diagflat_grad(self, xg, og):
        d = self._up(xg); self.ctx.diagflat_grad(d, self._up(og)); return d.numpy()
    def onehot(self, classes): return self.ctx.onehot(self._up(classes)).numpy()
    def onehot_grad(self, hc, classes, cg, og):
        d = self._up(cg); self.ctx.onehot_grad(hc, self._up(classes), d, self._up(og)); return d.numpy()
    def gemm(self, m, k, n, lhs, rhs): return self.ctx.gemm(m, k, n, self._up(lhs), self._up(rhs), mode=self.gemm_mode).numpy()
    def gemm_grad(self, m, k, n, lhs, rhs, lg, rg, og, accumulate=False):
        dl, dr = self._up(lg), self._up(rg)
        self.ctx.gemm_grad(m, k, n, self._up(lhs), self._up(rhs), dl, dr, self._up(og), accumulate, self.gemm_mode)
        return (None if dl is None else dl.numpy()), (None if dr is None else dr.numpy())
    def blas_gemm(self, m, n, k, a, b): return self.ctx.gemm_ex(False, False, m, n, k, self._up(a), self._up(b), mode=self.gemm_mode).numpy()
    def blas_gemmT(self, m, n, k, a, b): return self.ctx.gemm_nt(m, n, k, self._up(a), self._up(b), mode=self.gemm_mode).numpy()
    def blas_Tgemm(self, m, n, k, a, b): return self.ctx.gemm_tn(m, n, k, self._up(a), self._up(b), mode=self.gemm_mode).numpy()
    def gemm_ex(self, ta, tb, m, n, k, a, b, c=None, accumulate=False):
        d = self._up(c) if c is not None else None
        return self.ctx.gemm_ex(ta, tb, m, n, k, self._up(a), self._up(b), d, accumulate, self.gemm_mode).numpy()
    def sgd_step(self, w, g, lr):
        d = self._up(w); self.ctx.sgd_step(d, self._up(g), lr); return d.numpy()
    def chained_fwd(self, x, b): return self.ctx.chained_fwd(self._up(x), self._up(b)).numpy()
    def chained_bwd(self, x, b, xg, bg, og):
        dx, db = self._up(xg), self._up(bg)
        self.ctx.chained_bwd(self._up(x), self._up(b), dx, db, self._up(og)); return dx.numpy(), db.numpy()
