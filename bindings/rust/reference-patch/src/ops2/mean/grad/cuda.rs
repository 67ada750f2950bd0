//! `MeanRowsGrad` / `MeanColsGrad` for `CUDA<Mods>` (traits: src/ops2/mean/grad.rs:12-28), ACC:
//! rows: `x_grad[r,c] += (cols / len) * out_grad[c]` (factor formed first, mean/grad/cpu.rs:38-47); cols: `x_grad[r,c] += out_grad[r] / cols`.
use custos::{Buffer, OnDropBuffer, Shape, CUDA};
use sliced_b200_sys::*;

use super::{MeanColsGrad, MeanRowsGrad};
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: OnDropBuffer> MeanRowsGrad<T, IS, OS> for CUDA<Mods> {
    fn mean_rows_grad(&self, cols: usize, x_grad: &mut Buffer<T, Self, IS>, out_grad: &Buffer<T, Self, OS>) {
        let rows = x_grad.len() / cols;
        let rc = unsafe { sl_mean_rows_grad(self.ctx(), T::CODE, rows, cols, mptr(x_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: OnDropBuffer> MeanColsGrad<T, IS, OS> for CUDA<Mods> {
    fn mean_cols_grad(&self, cols: usize, x_grad: &mut Buffer<T, Self, IS>, out_grad: &Buffer<T, Self, OS>) {
        let rows = x_grad.len() / cols;
        let rc = unsafe { sl_mean_cols_grad(self.ctx(), T::CODE, rows, cols, mptr(x_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}
