//! `Gemm` for `CUDA<Mods>` (trait: src/ops2/gemm/mod.rs:21-32).  `mode = -1`: the context default (3xFP16, fp32-class).
use custos::{AddOperation, AsNoId, Buffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::Gemm;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T, LS, RS, OS, Mods> Gemm<T, LS, RS, OS> for CUDA<Mods>
where
    T: SlDtype + 'static,
    LS: Shape,
    RS: Shape,
    OS: Shape,
    Mods: Retrieve<Self, T, OS> + AddOperation + 'static,
{
    fn gemm(&self, m: usize, k: usize, n: usize, lhs: &Buffer<T, Self, LS>, rhs: &Buffer<T, Self, RS>) -> Buffer<T, Self, OS> {
        let mut out = self.retrieve(m * n, (lhs, rhs)).unwrap();
        // lazily executable, like the CPU impl: the closure only forwards raw pointers
        self.add_op((m.no_id(), k.no_id(), n.no_id(), lhs, rhs, &mut out), |(m, k, n, lhs, rhs, out)| {
            let dev = lhs.device();
            let rc = unsafe { sl_gemm(dev.ctx(), T::CODE, **m, **k, **n, cptr(lhs), cptr(rhs), mptr(out), -1) };
            dev.check(rc)
        })
        .unwrap();
        out
    }
}
