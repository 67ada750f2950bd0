//! `GemmGrad` for `CUDA<Mods>` (trait: src/ops2/gemm/grad.rs:14-29): `lhs_grad = out_grad * rhs^T`, `rhs_grad = lhs^T * out_grad`,
//! both SET like the CPU backend's BLAS calls (pass `accumulate = 1` for the OpenCL backend's `+=`).  One call: the library splits
//! `out_grad` once for both products.
use custos::{Buffer, OnDropBuffer, Shape, CUDA};
use sliced_b200_sys::*;

use super::GemmGrad;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, LS: Shape, RS: Shape, OS: Shape, Mods: OnDropBuffer> GemmGrad<T, LS, RS, OS> for CUDA<Mods> {
    fn gemm_grad(
        &self, m: usize, k: usize, n: usize, lhs: &Buffer<T, Self, LS>, rhs: &Buffer<T, Self, RS>, lhs_grad: &mut Buffer<T, Self, LS>,
        rhs_grad: &mut Buffer<T, Self, RS>, out_grad: &Buffer<T, Self, OS>,
    ) {
        // `requires_grad() == false` -> NULL: that half is skipped (gemm/grad/cpu_stack.rs:35,38)
        let lg = if lhs.requires_grad() { mptr(lhs_grad) } else { core::ptr::null_mut() };
        let rg = if rhs.requires_grad() { mptr(rhs_grad) } else { core::ptr::null_mut() };
        let rc = unsafe { sl_gemm_grad(self.ctx(), T::CODE, m, k, n, cptr(lhs), cptr(rhs), lg, rg, cptr(out_grad), 0, -1) };
        self.check(rc).unwrap();
    }
}
