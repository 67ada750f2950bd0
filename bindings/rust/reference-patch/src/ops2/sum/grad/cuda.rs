//! `SumRowsGrad` / `SumColsGrad` for `CUDA<Mods>` (traits: src/ops2/sum/grad.rs:12-28), ACC: `x_grad[r,c] += out_grad[c]` /
//! `x_grad[r,c] += out_grad[r]`.
use custos::{Buffer, OnDropBuffer, Shape, CUDA};
use sliced_b200_sys::*;

use super::{SumColsGrad, SumRowsGrad};
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: OnDropBuffer> SumRowsGrad<T, IS, OS> for CUDA<Mods> {
    fn sum_rows_grad(&self, cols: usize, x_grad: &mut Buffer<T, Self, IS>, out_grad: &Buffer<T, Self, OS>) {
        let rows = x_grad.len() / cols;
        let rc = unsafe { sl_sum_rows_grad(self.ctx(), T::CODE, rows, cols, mptr(x_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: OnDropBuffer> SumColsGrad<T, IS, OS> for CUDA<Mods> {
    fn sum_cols_grad(&self, cols: usize, x_grad: &mut Buffer<T, Self, IS>, out_grad: &Buffer<T, Self, OS>) {
        let rows = x_grad.len() / cols;
        let rc = unsafe { sl_sum_cols_grad(self.ctx(), T::CODE, rows, cols, mptr(x_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}
