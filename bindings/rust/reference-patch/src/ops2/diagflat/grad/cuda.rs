//! `DiagflatGrad` for `CUDA<Mods>` (trait: src/ops2/diagflat/grad.rs:13-15; CPU impl diagflat/grad/cpu.rs:32-36): `x_grad[i] += out_grad[i*n + i]`.
use custos::{Buffer, OnDropBuffer, Shape, CUDA};
use sliced_b200_sys::*;

use super::DiagflatGrad;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: OnDropBuffer> DiagflatGrad<T, IS, OS> for CUDA<Mods> {
    fn diagflat_grad(&self, x_grad: &mut Buffer<T, Self, IS>, out_grad: &Buffer<T, Self, OS>) {
        let n = x_grad.len();
        self.check(unsafe { sl_diagflat_grad(self.ctx(), T::CODE, n, mptr(x_grad), cptr(out_grad)) }).unwrap();
    }
}
