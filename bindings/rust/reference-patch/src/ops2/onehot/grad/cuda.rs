//! `onehot_grad` for `CUDA<Mods>` (CPU: src/ops2/onehot/grad/cpu.rs:3-12): `classes_grad[i] += out_grad[i*highest_class + classes[i]]`.
//! The reference exposes this as a free slice function; the device form takes the same arguments on `Buffer`s.
use custos::{Buffer, OnDropBuffer, CUDA};
use sliced_b200_sys::*;

use crate::cuda_device::{cptr, mptr, SlDevice};

pub fn cuda_onehot_grad<T: SlDtype, Mods: OnDropBuffer>(
    device: &CUDA<Mods>, highest_class: usize, classes: &Buffer<T, CUDA<Mods>>, classes_grad: &mut Buffer<T, CUDA<Mods>>, out_grad: &Buffer<T, CUDA<Mods>>,
) {
    let rc = unsafe { sl_onehot_grad(device.ctx(), T::CODE, classes.len(), highest_class, cptr(classes), mptr(classes_grad), cptr(out_grad)) };
    device.check(rc).unwrap();
}
