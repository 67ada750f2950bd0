//! `Diagflat` for `CUDA<Mods>` (trait: src/ops2/diagflat/mod.rs:21-41; CPU impl diagflat/cpu.rs:5-11,42-46).
//! Only the diagonal is written (diagflat/cpu.rs:42-46), so the retrieved buffer is cleared first: a `Cached` device hands back stale memory.
use custos::{Buffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::Diagflat;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype, IS: Shape, OS: Shape, Mods: Retrieve<Self, T, OS>> Diagflat<T, IS, OS> for CUDA<Mods> {
    fn diagflat(&self, x: &Buffer<T, Self, IS>) -> Buffer<T, Self, OS> {
        let mut out = self.retrieve(x.len() * x.len(), x).unwrap();
        let bytes = x.len() * x.len() * core::mem::size_of::<T>();
        self.check(unsafe { sl_clear(self.ctx(), mptr(&mut out), bytes) }).unwrap();
        self.check(unsafe { sl_diagflat(self.ctx(), T::CODE, x.len(), cptr(x), mptr(&mut out)) }).unwrap();
        out
    }
}
