// mlp.hpp — the training loops of the reference's examples, on the device, written against slh::Device / slh::Matrix
// exactly as the Rust examples are written against custos + sliced:
//
//   examples/nn.rs:13-54      Linear<I, O> { weights, bias }, forward = gemm + add_row_mut
//   examples/nn.rs:86-119     SGD { lr }, step: w -= grad * lr
//   examples/nn.rs:124-152    cce / cce_grad
//   examples/nn.rs:184-237    one epoch: zero_grad -> 3 x (Linear, relu | softmax) -> accuracy -> loss -> backward_with -> sgd
//   examples/sine_net.rs:135-163  the same with loss = (out - y)^2 and backward()
//
// Data-parallel extension (new functionality): every rank runs the step on its batch shard with
// cce_grad scaled by the GLOBAL batch, the parameter gradients live in one flat bucket that is sum-all-reduced
// (NCCL over NVLink) before the SGD step, so all replicas stay bit-identical.
#pragma once

#include <vector>

#include "sliced_host.hpp"

namespace slh {

struct Linear {  // examples/nn.rs:13-16
    Matrix weights;  // [I x O], require_grad
    Matrix bias;     // [1 x O], require_grad
    Matrix forward(const Matrix& inputs) const {  // examples/nn.rs:38-46
        Matrix out = inputs.gemm(weights);
        out.add_row_mut(bias);
        return out;
    }
};

enum LossKind { LOSS_SOFTMAX_CCE = 0 /* nn.rs */, LOSS_SQUARED = 1 /* sine_net.rs */ };

struct StepResult {
    double loss_sum = 0;   // sum over the local batch of the per-sample loss (mean = / batch)
    long long correct = 0; // argmax == label count (nn.rs:195-211)
};

class Mlp {
  public:
    Mlp(Device& dev, const std::vector<size_t>& dims, LossKind loss);
    ~Mlp();
    Device& device() { return dev_; }
    size_t n_layers() const { return layers_.size(); }
    const std::vector<size_t>& dims() const { return dims_; }
    Linear& layer(size_t l) { return layers_[l]; }
    // all parameter gradients, contiguous: [W0 | b0 | W1 | b1 | ...] — the all-reduce bucket
    Buf grad_bucket() { return bucket_; }
    Buf params() { return params_; }
    size_t n_params() const { return n_params_; }
    void* metrics_dptr() const { return metrics_dev_; }   // device [loss_sum f32][correct i32] of the last step (for sl_read_async)

    // One training step, op by op like the reference.  x: [batch x dims[0]] (no_grad), y: [batch x dims.back()],
    // labels: int32 [batch] or null.  grad_rows: the `rows` of cce_grad (nn.rs:151) — the global batch under DP.
    // Phases can be run separately so that a data-parallel driver can put the exchange between backward and sgd.
    StepResult forward_backward(const Buf& x, const Buf& y, const Buf& labels, size_t batch, size_t grad_rows, bool want_metrics);
    // The same step with the chains the tape exposes fused (what custos' `Lazy` graph + `optimize()` hook is for, nn.rs:302):
    //   gemm -> add_row_mut -> relu            => one gemm with a bias + relu epilogue that also emits the mask bits (sl_linear_fwd_bits)
    //   gemm_grad(lhs) -> relu grad            => one gemm with a mask-bit epilogue (sl_linear_bwd_input_relu_bits)
    //   zero_grad of activation gradients      => dropped: every activation gradient is produced by a SET kernel
    // Every per-element operation and its order are those of the unfused step, so losses, gradients and weights are
    // bit-identical to forward_backward() (tests/test_gpu_mlp.py::test_fused_step_is_bit_identical).
    StepResult forward_backward_fused(const Buf& x, const Buf& y, const Buf& labels, size_t batch, size_t grad_rows, bool want_metrics);
    // Data-parallel pipelining across steps (fused step with a communicator only; default off).  The gradient exchange of the weight
    // gradient computed LAST in backward has nothing left in its own step to hide behind.  Deferred mode (a) orders backward as the
    // dX chain first, then the weight gradients from the input layer up, so that the exposed exchange belongs to a late layer, and
    // (b) moves every layer's join + SGD update from the end of the step into the NEXT forward pass, right before that layer's gemm:
    // the late layers' exchanges then overlap the next step's input upload, operand split and first gemms.  Same arithmetic in the
    // same order on every parameter (update before use), so results are bit-identical; parameters are only current after flush()
    // (every accessor of the C bridge, predict(), forward_backward(), sgd() and allreduce_grads() flush first).
    void set_deferred(bool on) { flush(); deferred_ = on; }
    void flush();
    void set_fused(bool on) { flush(); fused_ = on; }
    bool fused() const { return fused_; }
    bool fused_active() const { return fused_ && loss_ == LOSS_SOFTMAX_CCE; }
    void allreduce_grads();          // sl_allreduce_sum over the bucket (no-op for a world of one)
    void sgd(double lr);             // SGD::step on every Linear (nn.rs:235-237)
    void exchange_and_sgd(double lr);  // both, with the update of each layer issued as soon as its exchange has completed
    // fused + squared loss + a shape sl_mlp_small_step takes (<= 4 layers of width <= 64, one rank): the WHOLE step, SGD included, is one
    // launch of one thread-block cluster (csrc/mlp_small.cu) — examples/sine_net.rs at its shipped sizes
    bool small_active(size_t batch) const;
    StepResult step_small(const Buf& x, const Buf& y, size_t batch, double lr, bool want_metrics);
    StepResult step(const Buf& x, const Buf& y, const Buf& labels, size_t batch, size_t grad_rows, double lr, bool want_metrics) {
        if (small_active(batch) && grad_rows == batch) return step_small(x, y, batch, lr, want_metrics);
        StepResult r = (fused_ && loss_ == LOSS_SOFTMAX_CCE) ? forward_backward_fused(x, y, labels, batch, grad_rows, want_metrics)
                                                                : forward_backward(x, y, labels, batch, grad_rows, want_metrics);
        exchange_and_sgd(lr);
        return r;
    }
    // The same step recorded once as a CUDA graph and replayed (custos `Lazy` + `run()`, examples/sine_net.rs:178-233): for the
    // launch-latency-bound shipped sizes (1000 x 64 matrices, ~60 kernels per step).  The first call runs eagerly (and allocates
    // every buffer), records the graph, later calls with the same buffers / sizes / lr are one cudaGraphLaunch.
    StepResult step_replay(const Buf& x, const Buf& y, const Buf& labels, size_t batch, size_t grad_rows, double lr, bool want_metrics);
    Matrix predict(const Buf& x, size_t batch);  // forward only, tape disabled

  private:
    Device& dev_;
    std::vector<size_t> dims_;
    LossKind loss_;
    std::vector<Linear> layers_;
    Buf params_;
    Buf bucket_;
    size_t n_params_ = 0;
    void* metrics_dev_ = nullptr;  // [loss_sum f32][correct i32]
    bool fused_ = false;
    bool exchanged_ = false;          // per-layer async all-reduces are in flight (fused backward)
    bool deferred_ = false;           // set_deferred
    bool have_pending_ = false;       // deferred mode: joins + SGD updates of the previous step are still owed
    double pending_lr_ = 0;
    std::vector<char> pending_;       // per layer
    bool deferred_active() const;
    void apply_pending_layer(size_t li);
    void* graph_ = nullptr;           // captured step (step_replay)
    struct GraphKey { const void *x, *y, *l; size_t batch, rows; double lr; bool fused; } gkey_{};
    std::vector<size_t> seg_off_;
    std::vector<int> layer_exchanges_;   // fused backward: number of exchanges issued up to and including layer l's
    std::vector<size_t> sgd_order_;      // layers in the order their exchanges were issued
    // persistent activations / activation gradients of the fused step (sized for the last batch seen)
    size_t fused_batch_ = 0;
    std::vector<Buf> z_, a_, gz_, mbits_;   // (z_ only for the last layer; mbits_: relu mask bits of the hidden layers)
    Buf loss_tmp_[4];
    StepResult read_metrics(bool want);
};

}  // namespace slh
