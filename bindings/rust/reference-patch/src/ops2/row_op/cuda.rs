//! `RowOp` for `CUDA<Mods>` (trait: src/ops2/row_op/mod.rs:17-47).  `add_row` / `add_row_mut` are the only call sites in the crate
//! and the examples; the opaque `Fn(&mut T, T, T)` of `row_op` cannot cross a C ABI, the four arithmetic ops are `sl_row_op(op)`.
use custos::{AddOperation, AsNoId, Buffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use super::RowOp;
use crate::cuda_device::{cptr, mptr, SlDevice};

impl<T, LS, RS, Mods> RowOp<T, LS, RS> for CUDA<Mods>
where
    T: SlDtype + 'static,
    LS: Shape,
    RS: Shape,
    Mods: Retrieve<Self, T, LS> + AddOperation + 'static,
{
    fn row_op<F: Fn(&mut T, T, T) + Copy>(&self, _cols: usize, _lhs: &Buffer<T, Self, LS>, _rhs: &Buffer<T, Self, RS>, _f: F) -> Buffer<T, Self, LS> {
        unimplemented!("sliced_b200: opaque row_op closures are not supported on CUDA; use add_row / sl_row_op(op)")
    }

    fn add_row(&self, cols: usize, lhs: &Buffer<T, Self, LS>, rhs: &Buffer<T, Self, RS>) -> Buffer<T, Self, LS>
    where
        T: core::ops::Add<Output = T>,
    {
        let mut out = self.retrieve(lhs.len(), (lhs, rhs)).unwrap();
        let rc = unsafe { sl_add_row(self.ctx(), T::CODE, lhs.len() / cols, cols, cptr(lhs), cptr(rhs), mptr(&mut out)) };
        self.check(rc).unwrap();
        out
    }

    fn add_row_mut(&self, rows: usize, cols: usize, lhs: &mut Buffer<T, Self, LS>, rhs: &Buffer<T, Self, RS>) {
        self.add_op((rows.no_id(), cols.no_id(), lhs, rhs), |(rows, cols, lhs, rhs)| {
            let dev = rhs.device();
            let rc = unsafe { sl_add_row_mut(dev.ctx(), T::CODE, **rows, **cols, mptr(lhs), cptr(rhs)) };
            dev.check(rc)
        })
        .unwrap();
    }
}
