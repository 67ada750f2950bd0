//! `MaxRows` / `MaxCols` and their gradients for `CUDA<Mods>` (traits: src/ops2/max/mod.rs:16-26, max/grad.rs:13-33).
//! Tie semantics of the CPU backend are kept: `max_rows_grad` feeds EVERY row equal to the column maximum, `max_cols_grad` the
//! FIRST column equal to the row maximum (max/grad/cpu.rs:46-54, 61-68).
use custos::{Buffer, OnDropBuffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::{MaxCols, MaxColsGrad, MaxRows, MaxRowsGrad};
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: Retrieve<Self, T, OS>> MaxRows<T, IS, OS> for CUDA<Mods> {
    fn max_rows(&self, cols: usize, x: &Buffer<T, Self, IS>) -> Buffer<T, Self, OS> {
        let mut out = self.retrieve(cols, x).unwrap();
        let rc = unsafe { sl_max_rows(self.ctx(), T::CODE, x.len() / cols, cols, cptr(x), mptr(&mut out), core::ptr::null_mut()) };
        self.check(rc).unwrap();
        out
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: Retrieve<Self, T, OS>> MaxCols<T, IS, OS> for CUDA<Mods> {
    fn max_cols(&self, rows: usize, cols: usize, x: &Buffer<T, Self, IS>) -> Buffer<T, Self, OS> {
        let mut out = self.retrieve(rows, x).unwrap();
        let rc = unsafe { sl_max_cols(self.ctx(), T::CODE, rows, cols, cptr(x), mptr(&mut out), core::ptr::null_mut()) };
        self.check(rc).unwrap();
        out
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: OnDropBuffer> MaxRowsGrad<T, IS, OS> for CUDA<Mods> {
    fn max_rows_grad(&self, cols: usize, out: &Buffer<T, Self, OS>, x: &Buffer<T, Self, IS>, x_grad: &mut Buffer<T, Self, IS>, out_grad: &Buffer<T, Self, OS>) {
        let rc = unsafe { sl_max_rows_grad(self.ctx(), T::CODE, x.len() / cols, cols, cptr(out), cptr(x), mptr(x_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: OnDropBuffer> MaxColsGrad<T, IS, OS> for CUDA<Mods> {
    fn max_cols_grad(&self, cols: usize, out: &Buffer<T, Self, OS>, x: &Buffer<T, Self, IS>, x_grad: &mut Buffer<T, Self, IS>, out_grad: &Buffer<T, Self, OS>) {
        let rc = unsafe { sl_max_cols_grad(self.ctx(), T::CODE, x.len() / cols, cols, cptr(out), cptr(x), mptr(x_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}
