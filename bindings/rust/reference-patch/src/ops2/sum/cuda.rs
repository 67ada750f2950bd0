//! `Sum` / `SumRows` / `SumCols` for `CUDA<Mods>` (traits: src/ops2/sum/mod.rs:15-25).  `sum_rows(cols, x)` has length `cols`
//! (reduces over the rows), `sum_cols(cols, x)` length `rows`.  Deterministic (no float atomics); outputs are fully overwritten even
//! when `Cached` hands back a stale buffer.  Mean* / Max* follow the same pattern with `sl_mean_*` / `sl_max_*`.
use custos::{Buffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::{Sum, SumCols, SumRows};
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype + Default, S: Shape, Mods: Retrieve<Self, T, ()>> Sum<T, S> for CUDA<Mods> {
    fn sum(&self, x: &Buffer<T, Self, S>) -> T {
        let mut out: Buffer<T, Self, ()> = self.retrieve(1, x).unwrap();
        let rc = unsafe { sl_sum(self.ctx(), T::CODE, cptr(x), x.len(), mptr(&mut out)) };
        self.check(rc).unwrap();
        out.read()[0] // sl_read: the one blocking call
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: Retrieve<Self, T, OS>> SumRows<T, IS, OS> for CUDA<Mods> {
    fn sum_rows(&self, cols: usize, x: &Buffer<T, Self, IS>) -> Buffer<T, Self, OS> {
        let mut out = self.retrieve(cols, x).unwrap();
        let rc = unsafe { sl_sum_rows(self.ctx(), T::CODE, x.len() / cols, cols, cptr(x), mptr(&mut out)) };
        self.check(rc).unwrap();
        out
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: Retrieve<Self, T, OS>> SumCols<T, IS, OS> for CUDA<Mods> {
    fn sum_cols(&self, cols: usize, x: &Buffer<T, Self, IS>) -> Buffer<T, Self, OS> {
        let rows = x.len() / cols;
        let mut out = self.retrieve(rows, x).unwrap();
        let rc = unsafe { sl_sum_cols(self.ctx(), T::CODE, rows, cols, cptr(x), mptr(&mut out)) };
        self.check(rc).unwrap();
        out
    }
}
