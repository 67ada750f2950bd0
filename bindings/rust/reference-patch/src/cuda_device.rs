//! Glue between custos' `CUDA<Mods>` device and libsliced_b200: the context handle and the error convention.
//! Every op impl in `src/ops2/*/cuda.rs` goes through these two methods.
use custos::{Buffer, Shape, CUDA};
use sliced_b200_sys::*;

pub trait SlDevice {
    /// the `*mut sl_ctx` created in `CUDA::new(idx)` (`sl_ctx_create`) and destroyed on drop (`sl_ctx_destroy`)
    fn ctx(&self) -> *mut sl_ctx;

    /// status code -> `custos::Result`; the op impls `.unwrap()` it like the CPU / OpenCL ones do
    fn check(&self, rc: core::ffi::c_int) -> custos::Result<()> {
        if rc == SL_OK {
            return Ok(());
        }
        let msg = unsafe { core::ffi::CStr::from_ptr(sl_last_error_string(self.ctx())) };
        Err(custos::Error::from(std::io::Error::new(std::io::ErrorKind::Other, format!("sliced_b200 error {rc}: {}", msg.to_string_lossy()))))
    }
}

impl<Mods> SlDevice for CUDA<Mods> {
    #[inline]
    fn ctx(&self) -> *mut sl_ctx {
        self.sliced_ctx() // accessor added to custos' CUDA device next to its stream / handle getters
    }
}

/// raw device pointers of a buffer (custos `CUDAPtr::ptr` is a `u64` device address)
#[inline]
pub fn cptr<T, D: custos::Device, S: Shape>(buf: &Buffer<T, D, S>) -> *const core::ffi::c_void
where
    D::Base<T, S>: custos::PtrType,
{
    buf.base().ptr() as *const core::ffi::c_void
}
#[inline]
pub fn mptr<T, D: custos::Device, S: Shape>(buf: &mut Buffer<T, D, S>) -> *mut core::ffi::c_void
where
    D::Base<T, S>: custos::PtrType,
{
    buf.base().ptr() as *mut core::ffi::c_void
}
