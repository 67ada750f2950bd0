//! `Transpose` / `TranposeGrad` for `CUDA<Mods>` (traits: src/ops2/transpose/mod.rs:17-21, grad.rs:13-26).  The gradient is SET,
//! as on the CPU backend (`accumulate = 0`); the OpenCL backend's `+=` is `accumulate = 1`.
use custos::{Buffer, OnDropBuffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::{TranposeGrad, Transpose};
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: Retrieve<Self, T, OS>> Transpose<T, IS, OS> for CUDA<Mods> {
    fn transpose(&self, rows: usize, cols: usize, x: &Buffer<T, Self, IS>) -> Buffer<T, Self, OS> {
        let mut out = self.retrieve(x.len(), x).unwrap();
        let rc = unsafe { sl_transpose(self.ctx(), T::CODE, rows, cols, cptr(x), mptr(&mut out), 0) };
        self.check(rc).unwrap();
        out
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: OnDropBuffer> TranposeGrad<T, IS, OS> for CUDA<Mods> {
    fn transpose_grad(&self, rows: usize, cols: usize, x_grad: &mut Buffer<T, Self, IS>, out_grad: &Buffer<T, Self, OS>) {
        // out_grad is cols x rows; transposing it back gives the rows x cols gradient
        let rc = unsafe { sl_transpose(self.ctx(), T::CODE, cols, rows, cptr(out_grad), mptr(x_grad), 0) };
        self.check(rc).unwrap();
    }
}
