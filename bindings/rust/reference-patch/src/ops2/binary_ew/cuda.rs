//! `BinaryElementWise` / `BinaryElementWiseGrad` / `AddElementWiseGrad` for `CUDA<Mods>` (trait: src/ops2/binary_ew/mod.rs:55-100,
//! grad.rs:22-45).  The provided methods add/sub/mul/div are overridden, so the common path never sees a closure; the generic
//! closure entry point recognises the four arithmetic ops the way the OpenCL backend turns the closure into source text
//! (binary_ew/opencl.rs:49) and refuses anything else — a precompiled library cannot run an arbitrary Rust closure.
use custos::{Buffer, Eval, MayToCLSource, Resolve, Retrieve, Retriever, Shape, ToMarker, CUDA};
use sliced_b200_sys::*;

use super::{AddElementWiseGrad, BinaryElementWise, BinaryElementWiseGrad};
use crate::cuda_device::{cptr, mptr, SlDevice};

fn launch<T: SlDtype, S: Shape, Mods: Retrieve<CUDA<Mods>, T, S>>(
    dev: &CUDA<Mods>, op: core::ffi::c_int, lhs: &Buffer<T, CUDA<Mods>, S>, rhs: &Buffer<T, CUDA<Mods>, S>,
) -> Buffer<T, CUDA<Mods>, S> {
    let mut out = dev.retrieve(lhs.len(), (lhs, rhs)).unwrap();
    let rc = unsafe { sl_binary_ew(dev.ctx(), T::CODE, op, cptr(lhs), cptr(rhs), mptr(&mut out), lhs.len()) };
    dev.check(rc).unwrap();
    out
}

/// which of `a+b`, `a-b`, `a*b`, `a/b` is this closure?  (evaluated on marker operands, as `to_cl_source` does)
fn classify<T, O: MayToCLSource>(f: impl Fn(Resolve<T>, Resolve<T>) -> O) -> Option<core::ffi::c_int> {
    let src: String = f("a".to_marker(), "b".to_marker()).to_cl_source().chars().filter(|c| !c.is_whitespace()).collect();
    match src.trim_matches(|c| c == '(' || c == ')') {
        "a+b" => Some(SL_ADD),
        "a-b" => Some(SL_SUB),
        "a*b" => Some(SL_MUL),
        "a/b" => Some(SL_DIV),
        _ => None,
    }
}

impl<T, S, Mods> BinaryElementWise<T, S> for CUDA<Mods>
where
    T: SlDtype + 'static,
    S: Shape,
    Mods: Retrieve<Self, T, S>,
{
    fn binary_ew<O>(&self, lhs: &Buffer<T, Self, S>, rhs: &Buffer<T, Self, S>, f: impl Fn(Resolve<T>, Resolve<T>) -> O + Copy + 'static) -> Buffer<T, Self, S>
    where
        O: Eval<T> + MayToCLSource,
    {
        let op = classify(f).expect("sliced_b200: binary_ew closure is not one of add/sub/mul/div (SL_ERR_UNSUPPORTED)");
        launch(self, op, lhs, rhs)
    }
    fn add(&self, lhs: &Buffer<T, Self, S>, rhs: &Buffer<T, Self, S>) -> Buffer<T, Self, S> where T: core::ops::Add<T, Output = T> {
        launch(self, SL_ADD, lhs, rhs)
    }
    fn sub(&self, lhs: &Buffer<T, Self, S>, rhs: &Buffer<T, Self, S>) -> Buffer<T, Self, S> where T: core::ops::Sub<T, Output = T> {
        launch(self, SL_SUB, lhs, rhs)
    }
    fn mul(&self, lhs: &Buffer<T, Self, S>, rhs: &Buffer<T, Self, S>) -> Buffer<T, Self, S> where T: core::ops::Mul<T, Output = T> {
        launch(self, SL_MUL, lhs, rhs)
    }
    fn div(&self, lhs: &Buffer<T, Self, S>, rhs: &Buffer<T, Self, S>) -> Buffer<T, Self, S> where T: core::ops::Div<T, Output = T> {
        launch(self, SL_DIV, lhs, rhs)
    }
}

impl<T: SlDtype, S: Shape, Mods: custos::OnDropBuffer> BinaryElementWiseGrad<T, S> for CUDA<Mods> {
    /// `lhs_grad += lgf(l, r) * og ; rhs_grad += rgf(l, r) * og` (ACC).  The tape closures of src/ops.rs:125-164 only ever pass the
    /// derivative pairs of add (1, 1), sub (1, -1) and mul (rhs, lhs); the op code travels with the closure registration.
    fn binary_ew_grad<LO, RO>(
        &self, lhs: &Buffer<T, Self, S>, rhs: &Buffer<T, Self, S>, lhs_grad: &mut Buffer<T, Self, S>, rhs_grad: &mut Buffer<T, Self, S>,
        out_grad: &Buffer<T, Self, S>, lhs_grad_fn: impl Fn(Resolve<T>, Resolve<T>) -> LO, rhs_grad_fn: impl Fn(Resolve<T>, Resolve<T>) -> RO,
    ) where
        LO: Eval<T> + MayToCLSource,
        RO: Eval<T> + MayToCLSource,
    {
        let l: String = lhs_grad_fn("a".to_marker(), "b".to_marker()).to_cl_source().chars().filter(|c| !c.is_whitespace()).collect();
        let r: String = rhs_grad_fn("a".to_marker(), "b".to_marker()).to_cl_source().chars().filter(|c| !c.is_whitespace()).collect();
        let op = match (l.as_str(), r.as_str()) {
            ("1", "1") => SL_ADD,
            ("1", "-1") | ("1", "(-1)") => SL_SUB,
            ("b", "a") => SL_MUL,
            _ => panic!("sliced_b200: unsupported binary_ew_grad closure pair ({l}, {r})"),
        };
        let rc = unsafe {
            sl_binary_ew_grad(self.ctx(), T::CODE, op, cptr(lhs), cptr(rhs), mptr(lhs_grad), mptr(rhs_grad), cptr(out_grad), out_grad.len())
        };
        self.check(rc).unwrap();
    }
}

impl<T: SlDtype, S: Shape, Mods: custos::OnDropBuffer> AddElementWiseGrad<T, S> for CUDA<Mods> {
    fn add_ew_grad(&self, lhs_grad: &mut Buffer<T, Self, S>, rhs_grad: &mut Buffer<T, Self, S>, out_grad: &Buffer<T, Self, S>) {
        let rc = unsafe { sl_add_ew_grad(self.ctx(), T::CODE, mptr(lhs_grad), mptr(rhs_grad), cptr(out_grad), out_grad.len()) };
        self.check(rc).unwrap();
    }
}
