//! `Softmax` for `CUDA<Mods>` (trait: src/ops2/softmax/mod.rs:16-18): one kernel instead of the CPU backend's five passes
//! (max_cols, sub_cols, exp, sum_cols, div_cols: softmax/cpu.rs:11-16).
use custos::{Buffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::Softmax;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, S: Shape, Mods: Retrieve<Self, T, S>> Softmax<T, S> for CUDA<Mods> {
    fn softmax(&self, samples: usize, features: usize, x: &Buffer<T, Self, S>) -> Buffer<T, Self, S> {
        let mut out = self.retrieve(x.len(), x).unwrap();
        let rc = unsafe { sl_softmax(self.ctx(), T::CODE, samples, features, cptr(x), mptr(&mut out)) };
        self.check(rc).unwrap();
        out
    }
}
