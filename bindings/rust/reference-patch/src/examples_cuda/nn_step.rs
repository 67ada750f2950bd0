//! What `examples/nn.rs` and `examples/sine_net.rs` call on a `CUDA` device when they want the fused forms instead of the op-by-op
//! tape (optional: the op-by-op chain through the `cuda.rs` trait impls gives the same results; see tests/test_gpu_mlp.py).
//! Authored, not compiled (see ../../README.md); argument lists are checked against the header by tests/test_abi.py.
use core::ffi::c_void;

use custos::{Buffer, CUDA};
use sliced_b200_sys::*;

use crate::cuda_device::{cptr, mptr, SlDevice};

/// `Linear<I, O>` of examples/nn.rs:13-46 followed by `relu` (src/matrix.rs:181): one gemm whose epilogue adds the bias row, applies
/// the relu and leaves the relu mask as ONE BIT per element — all that `relu`'s grad closure (`x.geq(0)`, matrix.rs:186) needs of `z`.
pub struct LinearRelu<'a, Mods> {
    pub weights: &'a Buffer<'a, f32, CUDA<Mods>>, // [I x O]
    pub bias: &'a Buffer<'a, f32, CUDA<Mods>>,    // [O]
}

impl<'a, Mods> LinearRelu<'a, Mods> {
    /// act[batch x O] = relu(x W + b); mask_bits: u32 [batch x ceil(O / 32)]
    pub fn forward(&self, dev: &CUDA<Mods>, batch: usize, i: usize, o: usize, x: &Buffer<f32, CUDA<Mods>>, act: &mut Buffer<f32, CUDA<Mods>>,
                   mask_bits: &mut Buffer<i32, CUDA<Mods>>) {
        let rc = unsafe {
            sl_linear_fwd_bits(dev.ctx(), SL_F32, batch, i, o, cptr(x), cptr(self.weights), cptr(self.bias), mptr(act), mptr(mask_bits) as *mut u32, -1)
        };
        dev.check(rc).unwrap();
    }

    /// the two grad closures the tape would run for this layer (gemm_grad + relu grad + add_row_mut grad), fused:
    ///   x_grad = mask(prev layer) * (out_grad W^T)      (skipped for the first layer: its input is `.no_grad()`, nn.rs:170)
    ///   W.grad = x^T out_grad (SET), b.grad += column sums of out_grad
    #[allow(clippy::too_many_arguments)]
    pub fn backward(&self, dev: &CUDA<Mods>, batch: usize, i: usize, o: usize, x: &Buffer<f32, CUDA<Mods>>, out_grad: &Buffer<f32, CUDA<Mods>>,
                    prev_mask_bits: Option<&Buffer<i32, CUDA<Mods>>>, x_grad: Option<&mut Buffer<f32, CUDA<Mods>>>,
                    w_grad: &mut Buffer<f32, CUDA<Mods>>, b_grad: &mut Buffer<f32, CUDA<Mods>>) {
        let rc = unsafe { sl_linear_bwd_params(dev.ctx(), SL_F32, batch, i, o, cptr(x), cptr(out_grad), mptr(w_grad), mptr(b_grad), -1) };
        dev.check(rc).unwrap();
        if let (Some(bits), Some(xg)) = (prev_mask_bits, x_grad) {
            let rc = unsafe {
                sl_linear_bwd_input_relu_bits(dev.ctx(), SL_F32, batch, i, o, cptr(self.weights), cptr(out_grad), cptr(bits) as *const u32, mptr(xg), -1)
            };
            dev.check(rc).unwrap();
        }
    }
}

/// examples/nn.rs:190-233 after the last `Linear`: softmax, accuracy, `cce`, `cce_grad` and `softmax_grad` in one launch.
/// Returns nothing: the probabilities, d logits, the per-sample losses and the correct-prediction counter stay on the device.
#[allow(clippy::too_many_arguments)]
pub fn softmax_cce_step<Mods>(dev: &CUDA<Mods>, batch: usize, classes: usize, global_batch: usize, logits: &Buffer<f32, CUDA<Mods>>,
                              targets: &Buffer<f32, CUDA<Mods>>, labels: &Buffer<i32, CUDA<Mods>>, probs: &mut Buffer<f32, CUDA<Mods>>,
                              logits_grad: &mut Buffer<f32, CUDA<Mods>>, loss_per_sample: &mut Buffer<f32, CUDA<Mods>>,
                              correct: &mut Buffer<i32, CUDA<Mods>>) {
    // `grad_rows` is the `rows` of cce_grad (nn.rs:151): the GLOBAL batch when the step is sharded over several GPUs
    let rc = unsafe {
        sl_softmax_cce(dev.ctx(), SL_F32, batch, classes, cptr(logits), cptr(targets), cptr(labels) as *const i32, global_batch, mptr(probs),
                       mptr(logits_grad), mptr(loss_per_sample), mptr(correct) as *mut i32)
    };
    dev.check(rc).unwrap();
}

/// `SGD::step` (nn.rs:108-119): the host loop `*value -= *grad * self.lr` needs `Deref<[T]>`, which a device buffer does not offer.
pub fn sgd_step<Mods>(dev: &CUDA<Mods>, lr: f64, param: &mut Buffer<f32, CUDA<Mods>>, grad: &Buffer<f32, CUDA<Mods>>) {
    let n = param.len();
    let rc = unsafe { sl_sgd_step(dev.ctx(), SL_F32, mptr(param), cptr(grad), lr, n) };
    dev.check(rc).unwrap();
}

/// One epoch body of nn.rs (:184-237) between `sl_gemm_scope_begin` / `_end`: inside the scope every activation is split into its
/// fp16 planes once (the forward gemm's planes serve the weight-gradient gemm of the same step).
pub fn with_step_scope<Mods, F: FnOnce()>(dev: &CUDA<Mods>, step: F) {
    dev.check(unsafe { sl_gemm_scope_begin(dev.ctx()) }).unwrap();
    step();
    dev.check(unsafe { sl_gemm_scope_end(dev.ctx()) }).unwrap();
}

/// examples/sine_net.rs:135-163 (1-64-64-1 on 1000 samples): forward, squared-error loss, backward and SGD of the whole net as ONE
/// launch when the shape fits (`sl_mlp_small_fits`); `params` / `grads` are the flat `[W0 | b0 | W1 | b1 | ...]` buffers the
/// `Linear`s are views of, `seg_off[2l]` / `seg_off[2l + 1]` the float offsets of layer l's weights / bias.
#[allow(clippy::too_many_arguments)]
pub fn sine_net_step<Mods>(dev: &CUDA<Mods>, dims: &[usize], seg_off: &[usize], batch: usize, x: &Buffer<f32, CUDA<Mods>>, y: &Buffer<f32, CUDA<Mods>>,
                           params: &mut Buffer<f32, CUDA<Mods>>, grads: &mut Buffer<f32, CUDA<Mods>>, lr: f64,
                           loss_sum: &mut Buffer<f32, CUDA<Mods>>) -> bool {
    let n_layers = (dims.len() - 1) as core::ffi::c_int;
    if unsafe { sl_mlp_small_fits(dev.ctx(), n_layers, dims.as_ptr(), batch) } == 0 {
        return false; // wider / deeper nets: run the op-by-op tape (or capture it once with sl_graph_begin / sl_graph_end)
    }
    let rc = unsafe {
        sl_mlp_small_step(dev.ctx(), SL_F32, n_layers, dims.as_ptr(), seg_off.as_ptr(), batch, cptr(x), cptr(y), mptr(params), mptr(grads), lr,
                          mptr(loss_sum) as *mut c_void)
    };
    dev.check(rc).unwrap();
    true
}
