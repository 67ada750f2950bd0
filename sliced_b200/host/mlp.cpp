// mlp.cpp — examples/nn.rs and examples/sine_net.rs training steps on the device (see mlp.hpp).
#include "mlp.hpp"

#include <cstdlib>

namespace slh {

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

Mlp::Mlp(Device& dev, const std::vector<size_t>& dims, LossKind loss) : dev_(dev), dims_(dims), loss_(loss) {
    if (dims.size() < 2) throw Error(SL_ERR_INVALID_ARG, "Mlp: need at least one layer");
    // one flat parameter buffer and one flat gradient bucket; Linear's weights/bias (and their grads) are views into them
    std::vector<size_t> off;
    size_t total = 0;
    for (size_t l = 0; l + 1 < dims.size(); ++l) {
        off.push_back(total);
        total += round_up(dims[l] * dims[l + 1], 64);
        off.push_back(total);
        total += round_up(dims[l + 1], 64);
    }
    n_params_ = total;
    off.push_back(total);
    seg_off_ = off;  // [W0, b0, W1, b1, ..., end]: layer l owns bucket floats [seg_off_[2l], seg_off_[2l+2])
    params_ = dev_.buffer(total, SL_F32);  // bias = zeros (Matrix::new, nn.rs:33); weights are written by the caller (rand, nn.rs:26)
    bucket_ = dev_.buffer(total, SL_F32);
    bucket_->requires_grad = false;
    for (size_t l = 0; l + 1 < dims.size(); ++l) {
        const size_t I = dims[l], O = dims[l + 1];
        Buf w = dev_.wrap((float*)params_->dptr + off[2 * l], I * O, SL_F32);
        Buf b = dev_.wrap((float*)params_->dptr + off[2 * l + 1], O, SL_F32);
        Buf gw = dev_.wrap((float*)bucket_->dptr + off[2 * l], I * O, SL_F32);
        Buf gb = dev_.wrap((float*)bucket_->dptr + off[2 * l + 1], O, SL_F32);
        gw->requires_grad = gb->requires_grad = false;
        dev_.bind_grad(w, gw);
        dev_.bind_grad(b, gb);
        layers_.push_back(Linear{Matrix(w, I, O), Matrix(b, 1, O)});
    }
    dev_.check(sl_malloc(dev_.ctx(), 16, &metrics_dev_));
    dev_.check(sl_clear(dev_.ctx(), metrics_dev_, 16));
}

bool Mlp::small_active(size_t batch) const {
    return fused_ && loss_ == LOSS_SQUARED && sl_comm_nranks(dev_.ctx()) <= 1 &&
           sl_mlp_small_fits(dev_.ctx(), (int)layers_.size(), dims_.data(), batch) != 0;
}

StepResult Mlp::step_small(const Buf& x, const Buf& y, size_t batch, double lr, bool want_metrics) {
    flush();
    dev_.flush_pending();
    dev_.check(sl_mlp_small_step(dev_.ctx(), SL_F32, (int)layers_.size(), dims_.data(), seg_off_.data(), batch, x->dptr, y->dptr, params_->dptr,
                                 bucket_->dptr, lr, metrics_dev_));
    return read_metrics(want_metrics);
}

Mlp::~Mlp() {
    try { flush(); } catch (...) {}
    if (graph_) sl_graph_destroy(dev_.ctx(), graph_);
    if (metrics_dev_) sl_free(dev_.ctx(), metrics_dev_);
}

StepResult Mlp::step_replay(const Buf& x, const Buf& y, const Buf& labels, size_t batch, size_t grad_rows, double lr, bool want_metrics) {
    if (deferred_active()) throw Error(SL_ERR_INVALID_ARG, "Mlp::step_replay: a deferred (cross-step pipelined) data-parallel step cannot be captured");
    flush();
    const GraphKey key{x->dptr, y->dptr, labels ? labels->dptr : nullptr, batch, grad_rows, lr, fused_};
    const bool same = graph_ && key.x == gkey_.x && key.y == gkey_.y && key.l == gkey_.l && key.batch == gkey_.batch &&
                      key.rows == gkey_.rows && key.lr == gkey_.lr && key.fused == gkey_.fused;
    if (same) {
        dev_.check(sl_graph_launch(dev_.ctx(), graph_));
        return read_metrics(want_metrics);
    }
    if (graph_) {
        dev_.check(sl_graph_destroy(dev_.ctx(), graph_));
        graph_ = nullptr;
    }
    // a captured step must be allocation-free: the op-by-op tape only is on a Cached device (same buffers every iteration);
    // the fused softmax-cce step keeps its own persistent activations
    if (!dev_.cached() && !(fused_ && loss_ == LOSS_SOFTMAX_CCE) && !small_active(batch))
        throw Error(SL_ERR_INVALID_ARG, "Mlp::step_replay: needs a Cached device (or the fused step): a captured step may not allocate or free");
    StepResult r = step(x, y, labels, batch, grad_rows, lr, want_metrics);   // this call's step, eagerly (creates every buffer)
    dev_.check(sl_graph_begin(dev_.ctx()));
    try {
        step(x, y, labels, batch, grad_rows, lr, false);                     // recorded, not executed
    } catch (...) {
        void* g = nullptr;
        sl_graph_end(dev_.ctx(), &g);
        if (g) sl_graph_destroy(dev_.ctx(), g);
        throw;
    }
    dev_.check(sl_graph_end(dev_.ctx(), &graph_));
    gkey_ = key;
    return r;
}

StepResult Mlp::forward_backward(const Buf& x, const Buf& y, const Buf& labels, size_t batch, size_t grad_rows, bool want_metrics) {
    flush();
    Device& d = dev_;
    const size_t L = layers_.size();
    const size_t out_cols = dims_.back();
    d.range_begin();   // `for epoch in device.range(..)` (nn.rs:184): rewind the Cached cursor
    d.zero_grad();     // nn.rs:186-188
    x->requires_grad = false;  // `.no_grad()` (nn.rs:170)
    y->requires_grad = false;

    Matrix out(x, batch, dims_[0]);
    for (size_t l = 0; l < L; ++l) {                       // nn.rs:190-193
        out = layers_[l].forward(out);
        if (l + 1 < L) out = out.relu();
        else if (loss_ == LOSS_SOFTMAX_CCE) out = out.softmax();
    }
    const size_t on = batch * out_cols;
    d.flush_pending();   // (fusion: the raw device pointers used below must hold their values)
    d.check(sl_clear(d.ctx(), metrics_dev_, 16));
    if (loss_ == LOSS_SOFTMAX_CCE) {
        if (labels)  // accuracy: nn.rs:195-211 (a host loop in the reference; a kernel + one int here)
            d.check(sl_count_correct(d.ctx(), SL_F32, batch, out_cols, out.data->dptr, (const int32_t*)labels->dptr, (int32_t*)metrics_dev_ + 1));
        // cce(preds, targets, cols): nn.rs:124-138 — L2 ops, nothing goes on the tape
        d.set_tape_enabled(false);
        Buf preds = d.clip(out.data, 1E-7, 1. - 1E-7);
        preds = d.binary_ew(SL_MUL, preds, y);
        Buf per_sample = d.sum_cols(out_cols, preds);
        Buf loss = d.apply_fn(per_sample, SL_UN_NEG_LN);
        // cce_grad(preds, targets, rows): nn.rs:140-152
        Buf grad = d.binary_ew(SL_DIV, y, out.data);
        grad = d.apply_fn(grad, SL_UN_NEG_DIV_SCALAR, (double)grad_rows);
        d.flush_pending();
        d.set_tape_enabled(true);
        d.check(sl_sum(d.ctx(), SL_F32, loss->dptr, batch, metrics_dev_));   // device.mean(&loss) * batch (nn.rs:224)
        d.backward_with(out.data, grad);                                      // nn.rs:233
    } else {
        Matrix ym(y, batch, out_cols);
        Matrix loss = out.sub(ym).pow(2.);                                    // sine_net.rs:150
        d.flush_pending();
        d.check(sl_sum(d.ctx(), SL_F32, loss.data->dptr, on, metrics_dev_));  // dev.mean(&loss) * len (sine_net.rs:151)
        d.backward(loss.data);                                                // sine_net.rs:156
    }
    StepResult r;
    if (want_metrics) {
        struct { float loss; int32_t correct; } m;
        d.check(sl_read(d.ctx(), &m, metrics_dev_, 8));  // the step's device -> host read
        r.loss_sum = m.loss;
        r.correct = m.correct;
    }
    return r;
}

StepResult Mlp::read_metrics(bool want) {
    StepResult r;
    if (want) {
        struct { float loss; int32_t correct; } m;
        dev_.check(sl_read(dev_.ctx(), &m, metrics_dev_, 8));
        r.loss_sum = m.loss;
        r.correct = m.correct;
    }
    return r;
}

StepResult Mlp::forward_backward_fused(const Buf& x, const Buf& y, const Buf& labels, size_t batch, size_t grad_rows, bool want_metrics) {
    Device& d = dev_;
    sl_ctx* c = d.ctx();
    const size_t L = layers_.size();
    const size_t oc = dims_.back();
    if (fused_batch_ != batch) {  // (re)allocate the persistent activations: z_l, a_l = relu(z_l), gz_l = d loss / d z_l
        z_.clear(); a_.clear(); gz_.clear(); mbits_.clear();
        for (size_t l = 0; l < L; ++l) {
            // hidden layers keep a_l = relu(z_l) and ONE BIT per element of (z_l >= 0) — all backward needs of z_l (the relu gradient,
            // src/matrix.rs:181-188); the logits z_{L-1} of the last layer are kept whole for the softmax
            z_.push_back(l + 1 == L ? d.buffer(batch * dims_[l + 1], SL_F32) : Buf());
            mbits_.push_back(l + 1 == L ? Buf() : d.buffer(batch * ((dims_[l + 1] + 31) / 32), SL_I32));
            a_.push_back(d.buffer(batch * dims_[l + 1], SL_F32));
            gz_.push_back(d.buffer(batch * dims_[l + 1], SL_F32));
        }
        loss_tmp_[1] = d.buffer(batch, SL_F32);       // per-sample loss
        fused_batch_ = batch;
    }
    d.check(sl_clear(c, metrics_dev_, 16));
    // every activation / weight / gz buffer is read by two gemms of this step (forward + a gradient gemm): split each into its
    // TF32 planes once.  Nothing but gemms writes those buffers between here and the end of the backward pass.
    struct GemmScope {   // closes the scope on every exit path: a throwing check() must not leave stale planes valid for later calls
        sl_ctx* c;
        explicit GemmScope(sl_ctx* ctx) : c(ctx) { sl_gemm_scope_begin(c); }
        ~GemmScope() { sl_gemm_scope_end(c); }
    } gemm_scope(c);

    // ---- forward: Linear + relu fused; the 10-class head goes through the skinny kernel + add_row_mut + softmax
    const void* in = x->dptr;
    for (size_t l = 0; l < L; ++l) {
        // deferred data-parallel mode: the previous step's join + SGD update of this layer, right before its weights are used
        if (have_pending_ && pending_[l]) apply_pending_layer(l);
        // parameter gradients accumulate (bias: += column sums) -> zero that segment of the bucket (the weight gradients are SET by
        // their gemm; the padding between segments is never written); activation gradients are all SET
        d.check(sl_clear(c, (float*)bucket_->dptr + seg_off_[2 * l + 1], (seg_off_[2 * l + 2] - seg_off_[2 * l + 1]) * sizeof(float)));
        if (l + 1 == L)
            d.check(sl_linear_fwd(c, SL_F32, batch, dims_[l], dims_[l + 1], in, layers_[l].weights.data->dptr, layers_[l].bias.data->dptr,
                                  z_[l]->dptr, nullptr, -1));
        else
            d.check(sl_linear_fwd_bits(c, SL_F32, batch, dims_[l], dims_[l + 1], in, layers_[l].weights.data->dptr, layers_[l].bias.data->dptr,
                                       a_[l]->dptr, (uint32_t*)mbits_[l]->dptr, -1));
        in = a_[l]->dptr;
    }
    if (have_pending_) {   // every owed update has been applied: join the communication stream before new exchanges are issued
        d.check(sl_comm_wait(c));
        exchanged_ = false;
        have_pending_ = false;
    }
    Buf out = a_[L - 1];
    // softmax -> accuracy -> cce -> cce_grad -> softmax_grad (nn.rs:190-233: nine launches over batch x 10 on the tape path) as ONE
    // kernel, bit-identical to that chain (tests/test_gpu_parity.py::test_softmax_cce_fused_equals_chain); + the loss sum
    d.check(sl_softmax_cce(c, SL_F32, batch, oc, z_[L - 1]->dptr, y->dptr, labels ? (const int32_t*)labels->dptr : nullptr, grad_rows, out->dptr,
                           gz_[L - 1]->dptr, loss_tmp_[1]->dptr, (int32_t*)metrics_dev_ + 1));
    d.check(sl_sum(c, SL_F32, loss_tmp_[1]->dptr, batch, metrics_dev_));   // device.mean(&loss) * batch (nn.rs:224)

    // ---- backward: the tape of nn.rs in reverse, with gemm_grad(lhs) + relu grad fused
    // Data-parallel schedule knobs (read per call, so a sweep can change them at run time; all bit-identical in their results):
    //   SLICED_DP_CHUNKS   the step's LAST weight-gradient gemm is produced in this many row blocks, each handed to the exchange as it
    //                      completes: the one exchange with no later gemm to hide behind shrinks to its last block.  Default 1: with the
    //                      launch-control gemm schedule one exchange per layer measured fastest at N=8 (profiles/r2c_dp_sweep_n8.txt:
    //                      4.25 ms vs 4.38 with 4 blocks).  SLICED_DP_CHUNKS_ALL=1 chunks every layer's weight gradient.
    //   SLICED_DP_ORDER    0: per layer dW then dX (tape order); 1: the dX chain first, then the weight gradients from the input layer
    //                      up (the exchanges of the early ones hide behind the later gemms)
    const int dp_chunks = env_int("SLICED_DP_CHUNKS", 1);
    const bool chunk_all = env_int("SLICED_DP_CHUNKS_ALL", 0) != 0;
    const int order = env_int("SLICED_DP_ORDER", deferred_active() ? 1 : 0);
    layer_exchanges_.assign(L, 0);
    auto params_grad = [&](size_t li, bool last) {
        const size_t I = dims_[li], O = dims_[li + 1];
        const void* lin = li == 0 ? x->dptr : a_[li - 1]->dptr;
        // b.grad += colsum(gz) ; W.grad = Tgemm(k,n,m,lhs,og) SET — one entry point so that gz is read once for both.  This layer's
        // gradients are then final -> their sum all-reduce starts on the communication stream while the remaining backward gemms keep
        // the tensor cores busy (no-op for a world of one)
        const int chunks = (dp_chunks > 1 && (last || chunk_all)) ? dp_chunks : 1;
        if (chunks > 1) {
            d.check(sl_linear_bwd_params_exchange(c, SL_F32, batch, I, O, lin, gz_[li]->dptr, d.grad(layers_[li].weights.data)->dptr,
                                                  d.grad(layers_[li].bias.data)->dptr, chunks, -1));
        } else {   // one exchange for the layer's whole [W | b] segment of the bucket
            d.check(sl_linear_bwd_params(c, SL_F32, batch, I, O, lin, gz_[li]->dptr, d.grad(layers_[li].weights.data)->dptr,
                                         d.grad(layers_[li].bias.data)->dptr, -1));
            d.check(sl_allreduce_sum_async(c, SL_F32, (float*)bucket_->dptr + seg_off_[2 * li], seg_off_[2 * li + 2] - seg_off_[2 * li]));
        }
        layer_exchanges_[li] = sl_comm_issued(c);
        sgd_order_.push_back(li);
    };
    auto input_grad = [&](size_t li) {
        d.check(sl_linear_bwd_input_relu_bits(c, SL_F32, batch, dims_[li], dims_[li + 1], layers_[li].weights.data->dptr, gz_[li]->dptr,
                                              (const uint32_t*)mbits_[li - 1]->dptr, gz_[li - 1]->dptr, -1));
    };
    sgd_order_.clear();
    if (order == 1) {
        for (size_t li = L; li-- > 1;) input_grad(li);
        for (size_t li = 0; li < L; ++li) params_grad(li, li + 1 == L || (li + 2 == L && dims_.back() <= 16));
    } else {
        for (size_t li = L; li-- > 0;) {
            params_grad(li, li == 0);
            if (li > 0) input_grad(li);
        }
    }
    exchanged_ = true;
    return read_metrics(want_metrics);
}

// Join of the per-layer exchanges and the SGD step, layer by layer in the order the exchanges were issued: the update of a layer
// whose summed gradients have arrived overlaps the exchanges still in flight (element-wise: identical to one flat sgd()).
bool Mlp::deferred_active() const { return deferred_ && fused_active() && sl_comm_nranks(dev_.ctx()) > 1; }

void Mlp::apply_pending_layer(size_t li) {
    sl_ctx* c = dev_.ctx();
    dev_.check(sl_comm_wait_n(c, layer_exchanges_[li]));
    const size_t off = seg_off_[2 * li], n = seg_off_[2 * li + 2] - off;
    dev_.check(sl_sgd_step(c, SL_F32, (float*)params_->dptr + off, (float*)bucket_->dptr + off, pending_lr_, n));
    pending_[li] = 0;
}

void Mlp::flush() {
    if (!have_pending_) return;
    for (size_t li : sgd_order_)
        if (pending_[li]) apply_pending_layer(li);
    dev_.check(sl_comm_wait(dev_.ctx()));
    exchanged_ = false;
    have_pending_ = false;
}

void Mlp::exchange_and_sgd(double lr) {
    sl_ctx* c = dev_.ctx();
    if (exchanged_ && deferred_active()) {   // owed to the next forward pass (or flush())
        pending_.assign(layers_.size(), 0);
        for (size_t li : sgd_order_) pending_[li] = 1;
        pending_lr_ = lr;
        have_pending_ = true;
        return;
    }
    if (!exchanged_ || sl_comm_nranks(c) <= 1 || !env_int("SLICED_DP_LAYER_SGD", 1)) {
        allreduce_grads();
        sgd(lr);
        return;
    }
    for (size_t li : sgd_order_) {
        dev_.check(sl_comm_wait_n(c, layer_exchanges_[li]));
        const size_t off = seg_off_[2 * li], n = seg_off_[2 * li + 2] - off;
        dev_.check(sl_sgd_step(c, SL_F32, (float*)params_->dptr + off, (float*)bucket_->dptr + off, lr, n));
    }
    dev_.check(sl_comm_wait(c));
    exchanged_ = false;
}

void Mlp::allreduce_grads() {
    flush();
    if (exchanged_) {  // the fused backward already issued per-layer exchanges: just join the communication stream
        dev_.check(sl_comm_wait(dev_.ctx()));
        exchanged_ = false;
        return;
    }
    dev_.check(sl_allreduce_sum(dev_.ctx(), SL_F32, bucket_->dptr, n_params_));
}

void Mlp::sgd(double lr) {
    flush();
    // SGD::step over lin1..lin3 params (nn.rs:235-237): parameters and gradients are both flat -> one kernel
    dev_.check(sl_sgd_step(dev_.ctx(), SL_F32, params_->dptr, bucket_->dptr, lr, n_params_));
}

Matrix Mlp::predict(const Buf& x, size_t batch) {
    flush();
    dev_.set_tape_enabled(false);
    Matrix out(x, batch, dims_[0]);
    for (size_t l = 0; l < layers_.size(); ++l) {
        out = layers_[l].forward(out);
        if (l + 1 < layers_.size()) out = out.relu();
        else if (loss_ == LOSS_SOFTMAX_CCE) out = out.softmax();
    }
    dev_.set_tape_enabled(true);
    return out;
}

}  // namespace slh
