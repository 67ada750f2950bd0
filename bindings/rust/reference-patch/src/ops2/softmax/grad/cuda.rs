//! `SoftmaxGrad` for `CUDA<Mods>` (trait: src/ops2/softmax/grad.rs:15-24): `x_grad[row] = J(row) * out_grad[row]` (SET) in closed
//! form `s * (g - <s, g>)` — the CPU backend materialises the F x F Jacobian per row (softmax/grad/cpu.rs:45-60).
use custos::{Buffer, OnDropBuffer, Shape, CUDA};
use sliced_b200_sys::*;

use super::SoftmaxGrad;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, S: Shape, Mods: OnDropBuffer> SoftmaxGrad<T, S> for CUDA<Mods> {
    fn softmax_grad(&self, samples: usize, features: usize, x_grad: &mut Buffer<T, Self, S>, out: &Buffer<T, Self, S>, out_grad: &Buffer<T, Self, S>) {
        let rc = unsafe { sl_softmax_grad(self.ctx(), T::CODE, samples, features, mptr(x_grad), cptr(out), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}
