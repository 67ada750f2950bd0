//! `Mean` / `MeanRows` / `MeanCols` for `CUDA<Mods>` (traits: src/ops2/mean/mod.rs:15-64): sums divided by the reduced extent
//! (integer division for integer element types, like mean/cpu.rs:71-78).
use custos::{Buffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::{Mean, MeanCols, MeanRows};
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype + Default, S: Shape, Mods: Retrieve<Self, T, ()>> Mean<T, S> for CUDA<Mods> {
    fn mean(&self, x: &Buffer<T, Self, S>) -> T {
        let mut out: Buffer<T, Self, ()> = self.retrieve(1, x).unwrap();
        let rc = unsafe { sl_mean(self.ctx(), T::CODE, cptr(x), x.len(), mptr(&mut out)) };
        self.check(rc).unwrap();
        out.read()[0]
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: Retrieve<Self, T, OS>> MeanRows<T, IS, OS> for CUDA<Mods> {
    fn mean_rows(&self, cols: usize, x: &Buffer<T, Self, IS>) -> Buffer<T, Self, OS> {
        let mut out = self.retrieve(cols, x).unwrap();
        let rc = unsafe { sl_mean_rows(self.ctx(), T::CODE, x.len() / cols, cols, cptr(x), mptr(&mut out)) };
        self.check(rc).unwrap();
        out
    }
}

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: Retrieve<Self, T, OS>> MeanCols<T, IS, OS> for CUDA<Mods> {
    fn mean_cols(&self, cols: usize, x: &Buffer<T, Self, IS>) -> Buffer<T, Self, OS> {
        let rows = x.len() / cols;
        let mut out = self.retrieve(rows, x).unwrap();
        let rc = unsafe { sl_mean_cols(self.ctx(), T::CODE, rows, cols, cptr(x), mptr(&mut out)) };
        self.check(rc).unwrap();
        out
    }
}
