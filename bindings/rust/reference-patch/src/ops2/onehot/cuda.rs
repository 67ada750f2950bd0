//! `Onehot` for `CUDA<Mods>` (trait: src/ops2/onehot/mod.rs:17-40; CPU impl onehot/cpu.rs:5-16,40-44).
//! `highest_class = max(classes) + 1` (onehot/cpu.rs:8) is a device reduction + one scalar read; only the ones are written, so the
//! retrieved buffer is cleared first.
use custos::{prelude::Number, Buffer, Retrieve, Retriever, CUDA};
use sliced_b200_sys::*;

use super::Onehot;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T: SlDtype + PartialOrd + Number, Mods: Retrieve<Self, T>> Onehot<T> for CUDA<Mods> {
    fn onehot(&self, classes: &Buffer<T, Self>) -> Buffer<T, Self> {
        let mut scalar: Buffer<T, Self> = self.retrieve(1, classes).unwrap();
        self.check(unsafe { sl_max(self.ctx(), T::CODE, cptr(classes), classes.len(), mptr(&mut scalar)) }).unwrap();
        let mut host = [T::default(); 1];
        self.check(unsafe { sl_read(self.ctx(), host.as_mut_ptr() as *mut core::ffi::c_void, cptr(&scalar), core::mem::size_of::<T>()) }).unwrap();
        let highest_class = host[0].as_usize() + 1;
        let mut out = self.retrieve(classes.len() * highest_class, classes).unwrap();
        self.check(unsafe { sl_clear(self.ctx(), mptr(&mut out), classes.len() * highest_class * core::mem::size_of::<T>()) }).unwrap();
        self.check(unsafe { sl_onehot(self.ctx(), T::CODE, classes.len(), highest_class, cptr(classes), mptr(&mut out)) }).unwrap();
        out
    }
}
