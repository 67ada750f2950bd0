//! `ColOp` for `CUDA<Mods>` (trait: src/ops2/col_op/mod.rs:20-58): `out[r,c] = lhs[r,c] (op) rhs[r]`.  `sub_cols` / `div_cols` — the
//! two call sites (softmax's building blocks) — are overridden; the opaque `Fn(T, T) -> T` of `col_op` cannot cross a C ABI.
use custos::{Buffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::ColOp;
use crate::cuda_device::{cptr, mptr, SlDevice};

fn launch<T: SlDtype, LS: Shape, RS: Shape, Mods: Retrieve<CUDA<Mods>, T, LS>>(
    dev: &CUDA<Mods>, op: core::ffi::c_int, cols: usize, lhs: &Buffer<T, CUDA<Mods>, LS>, rhs: &Buffer<T, CUDA<Mods>, RS>,
) -> Buffer<T, CUDA<Mods>, LS> {
    let mut out = dev.retrieve(lhs.len(), (lhs, rhs)).unwrap();
    let rc = unsafe { sl_col_op(dev.ctx(), T::CODE, op, lhs.len() / cols, cols, cptr(lhs), cptr(rhs), mptr(&mut out)) };
    dev.check(rc).unwrap();
    out
}

impl<T: SlDtype + 'static, LS: Shape, RS: Shape, Mods: Retrieve<Self, T, LS>> ColOp<T, LS, RS> for CUDA<Mods> {
    fn col_op<F>(&self, _cols: usize, _lhs: &Buffer<T, Self, LS>, _rhs: &Buffer<T, Self, RS>, _f: F) -> Buffer<T, Self, LS>
    where
        F: Fn(T, T) -> T + Copy + 'static,
    {
        unimplemented!("sliced_b200: opaque col_op closures are not supported on CUDA; use sub_cols / div_cols")
    }
    fn sub_cols(&self, cols: usize, lhs: &Buffer<T, Self, LS>, rhs: &Buffer<T, Self, RS>) -> Buffer<T, Self, LS> where T: core::ops::Sub<Output = T> {
        launch(self, SL_SUB, cols, lhs, rhs)
    }
    fn div_cols(&self, cols: usize, lhs: &Buffer<T, Self, LS>, rhs: &Buffer<T, Self, RS>) -> Buffer<T, Self, LS> where T: core::ops::Div<Output = T> {
        launch(self, SL_DIV, cols, lhs, rhs)
    }
}
