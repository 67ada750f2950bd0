//! `ColOpGrad` for `CUDA<Mods>` (trait: src/ops2/col_op/grad.rs:12-27; CPU impl col_op/grad/cpu.rs:7-89):
//! `lhs_grad[r,c] += lgf(l, rhs[r]) * og ; rhs_grad[r] += sum_c rgf(l, rhs[r]) * og` (the trait method is spelled `row_op_grad`).
//! As for `BinaryElementWiseGrad`, the pair of gradient closures is recognised by evaluating it on marker operands
//! (`MayToCLSource`-style source text): the pairs the crate and its tests use are those of sub (1, -1), add (1, 1), mul (r, l) and
//! div (1/r, l/-(r*r)) — col_op/grad/cpu.rs:97-161.
use custos::{Buffer, Eval, OnDropBuffer, Resolve, Shape, ToMarker, CUDA};
use sliced_b200_sys::*;

use super::ColOpGrad;
use crate::cuda_device::{cptr, mptr, SlDevice};

fn norm(s: String) -> String {
    s.chars().filter(|c| !c.is_whitespace()).collect::<String>().trim_matches(|c| c == '(' || c == ')').to_string()
}

fn classify(l: &str, r: &str) -> Option<core::ffi::c_int> {
    match (l, r) {
        ("1", "1") => Some(SL_ADD),
        ("1", "-1") => Some(SL_SUB),
        ("b", "a") => Some(SL_MUL),
        ("1/b", "a/-(b*b)") | ("1/b", "a/(-(b*b))") => Some(SL_DIV),
        _ => None,
    }
}

impl<T: SlDtype, LS: Shape, RS: Shape, Mods: OnDropBuffer> ColOpGrad<T, LS, RS> for CUDA<Mods> {
    fn row_op_grad<LhsGrad, RhsGrad>(
        &self, cols: usize, lhs: &Buffer<T, Self, LS>, rhs: &Buffer<T, Self, RS>, lhs_grad: &mut Buffer<T, Self, LS>,
        rhs_grad: &mut Buffer<T, Self, RS>, out_grad: &Buffer<T, Self, LS>, lhs_grad_fn: impl Fn(Resolve<T>, Resolve<T>) -> LhsGrad,
        rhs_grad_fn: impl Fn(Resolve<T>, Resolve<T>) -> RhsGrad,
    ) where
        LhsGrad: Eval<T> + ToString,
        RhsGrad: Eval<T> + ToString,
    {
        let l = norm(lhs_grad_fn("a".to_marker(), "b".to_marker()).to_string());
        let r = norm(rhs_grad_fn("a".to_marker(), "b".to_marker()).to_string());
        let op = classify(&l, &r).expect("sliced_b200: col_op gradient closures are not those of add/sub/mul/div (SL_ERR_UNSUPPORTED)");
        let rows = lhs.len() / cols;
        let rc = unsafe { sl_col_op_grad(self.ctx(), T::CODE, op, rows, cols, cptr(lhs), cptr(rhs), mptr(lhs_grad), mptr(rhs_grad), cptr(out_grad)) };
        self.check(rc).unwrap();
    }
}
