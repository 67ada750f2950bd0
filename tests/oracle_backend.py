"""Adapter: the CPU oracle with 'return the result' calling convention used by tests/kat_runner.py."""
import numpy as np

import oracle as O


class OracleBackend:
    name = "oracle"

    def binary_ew(self, op, lhs, rhs): return O.binary_ew(op, lhs, rhs)
    def binary_ew_grad(self, op, lhs, rhs, lg, rg, og):
        O.binary_ew_grad(op, lhs, rhs, lg, rg, og); return lg, rg
    def add_ew_grad(self, lg, rg, og):
        O.add_ew_grad(lg, rg, og); return lg, rg
    def unary(self, op, x, p0=0.0, p1=0.0): return O.unary(op, x, p0, p1)
    def unary_grad(self, op, x, xg, og, p0=0.0, p1=0.0):
        O.unary_grad(op, x, xg, og, p0, p1); return xg
    def row_op(self, op, cols, lhs, rhs): return O.row_op(op, cols, lhs, rhs)
    def add_row_mut(self, rows, cols, lhs, rhs):
        O.add_row_mut(rows, cols, lhs, rhs); return lhs
    def add_row_grad(self, rows, cols, lg, rg, og):
        O.add_row_grad(rows, cols, lg, rg, og); return lg, rg
    def add_row_mut_grad(self, rows, cols, rg, og):
        O.add_row_mut_grad(rows, cols, rg, og); return rg
    def row_op_grad(self, op, cols, lhs, rhs, lg, rg, og):
        O.row_op_grad(op, cols, lhs, rhs, lg, rg, og); return lg, rg
    def col_op(self, op, cols, lhs, rhs): return O.col_op(op, cols, lhs, rhs)
    def col_op_grad(self, op, cols, lhs, rhs, lg, rg, og):
        O.col_op_grad(op, cols, lhs, rhs, lg, rg, og); return lg, rg
    def sum(self, x): return O.sum_(x)
    def mean(self, x): return O.mean(x)
    def max(self, x): return O.max_(x)
    def sum_rows(self, cols, x): return O.sum_rows(cols, x)
    def sum_cols(self, cols, x): return O.sum_cols(cols, x)
    def mean_rows(self, cols, x): return O.mean_rows(cols, x)
    def mean_cols(self, cols, x): return O.mean_cols(cols, x)
    def max_rows(self, cols, x): return O.max_rows(cols, x)
    def max_rows_noinit(self, cols, x, out): O.max_rows_noinit(cols, x, out)
    def max_cols(self, cols, x): return O.max_cols(cols, x)
    def max_grad(self, out, x, xg): O.max_grad(out, x, xg)
    def sum_rows_grad(self, cols, xg, og):
        O.sum_rows_grad(cols, xg, og); return xg
    def sum_cols_grad(self, cols, xg, og):
        O.sum_cols_grad(cols, xg, og); return xg
    def mean_rows_grad(self, cols, xg, og):
        O.mean_rows_grad(cols, xg, og); return xg
    def mean_cols_grad(self, cols, xg, og):
        O.mean_cols_grad(cols, xg, og); return xg
    def max_rows_grad(self, cols, out, x, xg, og):
        O.max_rows_grad(cols, out, x, xg, og); return xg
    def max_cols_grad(self, cols, out, x, xg, og):
        O.max_cols_grad(cols, out, x, xg, og); return xg
    def transpose(self, rows, cols, x, out=None, accumulate=False):
        # accumulate=True -> AOS=Assign without the CPU quirk line (the OpenCL reference's behaviour)
        return O.transpose(rows, cols, x, out, assign=accumulate, quirk=not accumulate)
    def softmax(self, samples, features, x): return O.softmax(samples, features, x)
    def softmax_grad(self, samples, features, xg, out, og):
        O.softmax_grad(samples, features, xg, out, og); return xg
    def diagflat(self, x): return O.diagflat(x)
    def diagflat_grad(self, xg, og):
        O.diagflat_grad(xg, og); return xg
    def onehot(self, classes): return O.onehot(classes)
    def onehot_grad(self, hc, classes, cg, og):
        O.onehot_grad(hc, classes, cg, og); return cg
    def gemm(self, m, k, n, lhs, rhs): return O.gemm(m, k, n, lhs, rhs)
    def gemm_grad(self, m, k, n, lhs, rhs, lg, rg, og, accumulate=False):
        O.gemm_grad(m, k, n, lhs, rhs, lg, rg, og, accumulate); return lg, rg
    def blas_gemm(self, m, n, k, a, b): return O.blas_gemm(m, n, k, a, b)
    def blas_gemmT(self, m, n, k, a, b): return O.blas_gemmT(m, n, k, a, b)
    def blas_Tgemm(self, m, n, k, a, b): return O.blas_Tgemm(m, n, k, a, b)
    def sgd_step(self, w, g, lr):
        O.sgd_step(w, g, lr); return w
    def chained_fwd(self, x, b): return O.chained_fwd(x, b)
    def chained_bwd(self, x, b, xg, bg, og):
        O.chained_bwd(x, b, xg, bg, og); return xg, bg
