//! `RowOpGrad` for `CUDA<Mods>` (trait: src/ops2/row_op/grad.rs:12-41): `add_row_grad` copies out_grad into lhs_grad (SET) and adds
//! its column sums to rhs_grad (ACC) in one pass; `add_row_mut_grad` only the column sums.
use custos::{Buffer, OnDropBuffer, Shape, CUDA};
use sliced_b200_sys::*;

use super::RowOpGrad;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, LS: Shape, RS: Shape, Mods: OnDropBuffer> RowOpGrad<T, LS, RS> for CUDA<Mods> {
    fn row_op_grad(
        &self, _cols: usize, _lhs: &Buffer<T, Self, LS>, _rhs: &Buffer<T, Self, RS>, _lhs_grad: &mut Buffer<T, Self, LS>,
        _rhs_grad: &mut Buffer<T, Self, RS>, _out_grad: &Buffer<T, Self, LS>, _lhs_grad_fn: impl Fn(T) -> T, _rhs_grad_fn: impl Fn(T) -> T,
    ) {
        unimplemented!("sliced_b200: opaque derivative closures are not supported on CUDA; the arithmetic ops are sl_row_op_grad(op)")
    }

    fn add_row_grad(&self, rows: usize, cols: usize, lhs_grad: &mut Buffer<T, Self, LS>, rhs_grad: &mut Buffer<T, Self, RS>, out_grad: &Buffer<T, Self, LS>) {
        let rc = unsafe { sl_add_row_grad(self.ctx(), T::CODE, rows, cols, mptr(lhs_grad), mptr(rhs_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }

    fn add_row_mut_grad(&self, rows: usize, cols: usize, rhs_grad: &mut Buffer<T, Self, RS>, out_grad: &Buffer<T, Self, LS>) {
        let rc = unsafe { sl_add_row_mut_grad(self.ctx(), T::CODE, rows, cols, mptr(rhs_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}
