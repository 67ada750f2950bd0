// sliced_host.cpp — implementation of the host-side mirror declared in sliced_host.hpp.
// Forward ops call the C ABI; every *MayGrad op pushes the reference's grad closure on the tape.
#include "sliced_host.hpp"

#include "chain_builder.hpp"

#include <algorithm>
#include <cstring>

namespace slh {

BufferImpl::~BufferImpl() {
    if (dev) dev->drop_grad(id);   // custos drops a buffer's gradient with the buffer (OnDropBuffer †)
    if (owns && dptr && dev && dev->ctx()) sl_free(dev->ctx(), dptr);
}

Device::Device(int device_index, bool cached, void* cuda_stream, bool borrow_stream) : cached_(cached) {
    int rc = borrow_stream ? sl_ctx_create_on_stream(device_index, cuda_stream, &ctx_) : sl_ctx_create(device_index, &ctx_);
    if (rc != SL_OK) throw Error(rc, sl_last_error_string(nullptr));
    check(sl_malloc(ctx_, 16, &scalar_dev_));
}

void Device::drop_grad(uint64_t id) {
    if (tearing_down_) return;
    auto it = grads_.find(id);
    if (it == grads_.end()) return;
    Buf g = std::move(it->second);   // destroyed after the erase: its own destructor re-enters drop_grad with another id
    grads_.erase(it);
}

Device::~Device() {
    tearing_down_ = true;
    pending_.clear();
    tape_.clear();
    grads_.clear();
    cache_.clear();
    if (ctx_) {
        if (scalar_dev_) sl_free(ctx_, scalar_dev_);
        sl_ctx_destroy(ctx_);
        ctx_ = nullptr;
    }
}

void Device::check(int rc) const {
    if (rc != SL_OK) throw Error(rc, sl_last_error_string(ctx_));
}

Buf Device::new_buffer(size_t len, int dtype, bool zero) {
    auto b = std::make_shared<BufferImpl>();
    b->id = next_id_++;
    b->dtype = dtype;
    b->len = len;
    b->dev = this;
    check(sl_malloc(ctx_, b->bytes() ? b->bytes() : 4, &b->dptr));
    if (zero && len) check(sl_clear(ctx_, b->dptr, b->bytes()));
    return b;
}

Buf Device::buffer(size_t len, int dtype) { return new_buffer(len, dtype, true); }

Buf Device::from_host(const void* host, size_t len, int dtype) {
    Buf b = new_buffer(len, dtype, false);
    if (len) {
        check(sl_write(ctx_, b->dptr, host, b->bytes()));
        check(sl_sync(ctx_));  // the host source may be pageable / short-lived
    }
    return b;
}

Buf Device::wrap(void* dptr, size_t len, int dtype) {
    auto b = std::make_shared<BufferImpl>();
    b->id = next_id_++;
    b->dtype = dtype;
    b->len = len;
    b->dev = this;
    b->dptr = dptr;
    b->owns = false;
    return b;
}

void Device::read(const Buf& b, void* host) const {
    const_cast<Device*>(this)->flush_pending();
    if (b->len && !b->dptr) throw Error(SL_ERR_INVALID_ARG, "read: this buffer was an unobserved intermediate of a fused chain (never materialised)");
    check(sl_read(ctx_, host, b->dptr, b->bytes()));
}
void Device::write(const Buf& b, const void* host) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    check(sl_write(ctx_, b->dptr, host, b->bytes()));
    check(sl_sync(ctx_));
}
void Device::sync() const {
    const_cast<Device*>(this)->flush_pending();
    check(sl_sync(ctx_));
}

// custos `Retriever::retrieve` †.  Without `Cached`: a fresh zeroed buffer.  With `Cached`: the buffer handed out at the
// same cursor position in the previous iteration (contents STALE: every op below fully overwrites its output).
Buf Device::retrieve(size_t len, int dtype) {
    if (!cached_) return new_buffer(len, dtype, true);
    if (cursor_ < cache_.size() && cache_[cursor_]->len == len && cache_[cursor_]->dtype == dtype) return cache_[cursor_++];
    Buf b = new_buffer(len, dtype, true);
    if (cursor_ < cache_.size()) cache_[cursor_] = b;
    else cache_.push_back(b);
    ++cursor_;
    return b;
}

Buf Device::grad(const Buf& b) {
    auto it = grads_.find(b->id);
    if (it != grads_.end()) return it->second;
    Buf g = new_buffer(b->len, b->dtype, true);
    g->requires_grad = false;
    grads_[b->id] = g;
    return g;
}

void Device::zero_grad() {
    for (auto& kv : grads_)
        if (kv.second->len) check(sl_clear(ctx_, kv.second->dptr, kv.second->bytes()));
}

void Device::backward(const Buf& out) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    Buf g = grad(out);
    check(sl_fill(ctx_, out->dtype, g->dptr, 1.0, out->len));  // custos seeds the output gradient with ones †
    // reverse registration order; eager mode clears the grad fns afterwards †
    std::vector<std::function<void()>> fns;
    fns.swap(tape_);
    for (auto it = fns.rbegin(); it != fns.rend(); ++it) (*it)();
}

void Device::backward_with(const Buf& out, const Buf& seed) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    if (seed->len != out->len) throw Error(SL_ERR_INVALID_ARG, "backward_with: seed length mismatch");
    Buf g = grad(out);
    check(sl_copy(ctx_, g->dptr, seed->dptr, out->bytes()));
    std::vector<std::function<void()>> fns;
    fns.swap(tape_);
    for (auto it = fns.rbegin(); it != fns.rend(); ++it) (*it)();
}

static void same_kind(const Buf& a, const Buf& b, const char* what) {
    if (a->dtype != b->dtype) throw Error(SL_ERR_INVALID_ARG, std::string(what) + ": dtype mismatch");
}

// ------------------------------------------------------------------ binary ops (src/ops.rs:115-187)
// the launch-per-op form of one element-wise node: forward into n.dst (already allocated) + the reference's grad closure
void Device::run_node_unfused(const FNode& n) {
    if (n.kind == 1) {
        const Buf lhs = n.a, rhs = n.b, out = n.dst;
        const int op = n.op;
        const bool add2 = n.add2;
        check(sl_binary_ew(ctx_, lhs->dtype, op, lhs->dptr, rhs->dptr, out->dptr, lhs->len));
        if (n.with_grad)
            add_grad_fn([this, op, lhs, rhs, out, add2]() {
                Buf og = grad(out);
                Buf lg = grad(lhs), rg = grad(rhs);
                size_t len = std::min(std::min(lhs->len, rhs->len), og->len);  // binary_ew/grad/cpu_stack.rs:54
                if (add2) check(sl_add_ew_grad(ctx_, lhs->dtype, lg->dptr, rg->dptr, og->dptr, len));
                else check(sl_binary_ew_grad(ctx_, lhs->dtype, op, lhs->dptr, rhs->dptr, lg->dptr, rg->dptr, og->dptr, len));
            });
    } else {
        const Buf x = n.a, out = n.dst;
        const int unop = n.op;
        const double p0 = n.p0, p1 = n.p1;
        check(sl_unary(ctx_, x->dtype, unop, p0, p1, x->dptr, out->dptr, x->len));
        if (n.with_grad)
            add_grad_fn([this, x, out, unop, p0, p1]() {
                check(sl_unary_grad(ctx_, x->dtype, unop, p0, p1, x->dptr, grad(x)->dptr, grad(out)->dptr, x->len));
            });
    }
}

// Either runs the node now (fusion off) or appends it to the open chain and hands back a buffer whose value is still pending.
Buf Device::record(FNode n) {
    n.with_grad = n.with_grad && tape_enabled_;
    const bool fusable = fusion_ && n.a->len > 0 && (n.kind == 2 || n.b->len == n.a->len);
    if (!fusable) {
        flush_pending();
        n.dst = retrieve(n.a->len, n.a->dtype);  // binary_ew/cpu_stack.rs:32: len = lhs.len()
        run_node_unfused(n);
        return n.dst;
    }
    if (!pending_.empty() && (pending_.front().a->len != n.a->len || pending_.front().a->dtype != n.a->dtype || pending_.size() >= 8)) flush_pending();
    auto b = std::make_shared<BufferImpl>();   // virtual until the chain is flushed: only results somebody still holds get memory
    b->id = next_id_++;
    b->dtype = n.a->dtype;
    b->len = n.a->len;
    b->dev = this;
    b->pending = true;
    n.dst = b;
    pending_.push_back(std::move(n));
    return b;
}

namespace {
struct FusedGroup {   // what the single grad closure of a fused chain needs at backward time
    ChainBuilder cb;
    std::vector<Buf> leaves;
    std::vector<int> leaf_value;   // value id of leaves[i]
    std::vector<Buf> live;         // materialised node results
    std::vector<int> live_value;
    size_t len = 0;
    int dtype = SL_F32;
};
}  // namespace

void Device::flush_pending() {
    if (pending_.empty()) return;
    std::vector<FNode> nodes;
    nodes.swap(pending_);
    const size_t N = nodes.size();
    auto group = std::make_shared<FusedGroup>();
    group->len = nodes[0].a->len;
    group->dtype = nodes[0].a->dtype;
    std::unordered_map<uint64_t, int> value_of;   // buffer id -> chain value id
    auto value = [&](const Buf& b) {
        auto it = value_of.find(b->id);
        if (it != value_of.end()) return it->second;
        const int v = group->cb.input();
        value_of[b->id] = v;
        group->leaves.push_back(b);
        group->leaf_value.push_back(v);
        return v;
    };
    std::vector<int> node_value(N);
    bool any_grad = false;
    for (size_t i = 0; i < N; ++i) {
        const FNode& n = nodes[i];
        const int a = value(n.a);
        const int v = n.kind == 1 ? group->cb.binary(n.op, a, value(n.b)) : group->cb.unary(n.op, a, n.p0, n.p1);
        if (!n.with_grad) group->cb.no_grad(v);
        any_grad = any_grad || n.with_grad;
        value_of[n.dst->id] = v;
        node_value[i] = v;
    }
    // liveness: a result is materialised iff somebody outside this chain still holds it.  References we know of: the node itself
    // and every later node that reads it.
    std::vector<size_t> live_nodes;
    for (size_t i = 0; i < N; ++i) {
        long expected = 1;
        for (size_t j = i + 1; j < N; ++j) {
            if (nodes[j].a == nodes[i].dst) ++expected;
            if (nodes[j].kind == 1 && nodes[j].b == nodes[i].dst) ++expected;
        }
        if (nodes[i].dst.use_count() > expected) live_nodes.push_back(i);
    }
    if (live_nodes.empty()) return;   // nobody can observe any of it (and no gradient can reach it)
    // does it fit?  <= 3 leaves, <= 2 materialised results, forward and worst-case backward programs within the interpreter's limits
    sl_chain_prog fwd;
    bool fits = group->leaves.size() <= 3 && live_nodes.size() <= 2;
    std::vector<int> outs;
    for (size_t i : live_nodes) outs.push_back(node_value[i]);
    fits = fits && group->cb.build_forward(outs, &fwd, nullptr);
    if (fits && any_grad) {
        std::vector<int> wrt;
        for (size_t k = 0; k < group->leaves.size(); ++k)
            if (group->leaves[k]->requires_grad) wrt.push_back(group->leaf_value[k]);
        sl_chain_prog worst;
        if (!wrt.empty()) fits = group->cb.build_backward(outs, wrt, &worst, nullptr);
    }
    auto materialise = [&](const Buf& dst) {
        Buf mem = retrieve(dst->len, dst->dtype);
        dst->dptr = mem->dptr;
        dst->owns = false;
        dst->backing = mem;
        dst->pending = false;
    };
    if (!fits) {   // the launch-per-op path, exactly as without fusion
        ++unfused_groups_;
        for (auto& n : nodes) {
            materialise(n.dst);
            run_node_unfused(n);
        }
        return;
    }
    ++fused_groups_;
    const void* in_ptrs[SL_CHAIN_MAX_INPUTS];
    void* out_ptrs[SL_CHAIN_MAX_OUTPUTS];
    for (size_t k = 0; k < group->leaves.size(); ++k) in_ptrs[k] = group->leaves[k]->dptr;
    for (size_t k = 0; k < live_nodes.size(); ++k) {
        const Buf& dst = nodes[live_nodes[k]].dst;
        materialise(dst);
        out_ptrs[k] = dst->dptr;
        group->live.push_back(dst);
        group->live_value.push_back(node_value[live_nodes[k]]);
    }
    for (auto& n : nodes) n.dst->pending = false;
    check(sl_fused_chain(ctx_, group->dtype, &fwd, in_ptrs, out_ptrs, group->len));
    if (!any_grad) return;
    // ONE grad closure for the whole chain: the tape of its ops in reverse order as a single fused launch
    add_grad_fn([this, group]() {
        std::vector<int> seeds, wrt;
        std::vector<Buf> seed_grads, wrt_grads;
        bool any_seed = false;
        for (size_t k = 0; k < group->live.size(); ++k) any_seed = any_seed || has_grad(group->live[k]);
        if (!any_seed) return;   // no gradient reaches this chain
        // every materialised result is a seed: one that nothing outside contributed to starts from zeros, and still receives (and
        // shows through `.grad()`) what the chain's own later ops contribute — as its buffer would in the reference
        for (size_t k = 0; k < group->live.size(); ++k) {
            seeds.push_back(group->live_value[k]);
            seed_grads.push_back(grad(group->live[k]));
        }
        for (size_t k = 0; k < group->leaves.size(); ++k)
            if (group->leaves[k]->requires_grad) {
                wrt.push_back(group->leaf_value[k]);
                wrt_grads.push_back(grad(group->leaves[k]));
            }
        if (seeds.empty() || wrt.empty()) return;
        sl_chain_prog bwd;
        std::string why;
        std::vector<int> seed_totals;
        if (!group->cb.build_backward(seeds, wrt, &bwd, &why, &seed_totals)) throw Error(SL_ERR_UNSUPPORTED, "fused chain backward: " + why);
        const void* ins[SL_CHAIN_MAX_INPUTS];
        void* outs_[SL_CHAIN_MAX_OUTPUTS];
        size_t q = 0;
        for (auto& l : group->leaves) ins[q++] = l->dptr;
        for (auto& g : seed_grads) ins[q++] = g->dptr;
        for (auto& g : wrt_grads) ins[q++] = g->dptr;
        size_t o = 0;
        for (auto& g : wrt_grads) outs_[o++] = g->dptr;     // in place: old + contributions in tape order
        // extra outputs: seeds that were also consumed inside the chain take those contributions too (the builder says which)
        for (int si : seed_totals) outs_[o++] = seed_grads[si]->dptr;
        check(sl_fused_chain(ctx_, group->dtype, &bwd, ins, outs_, group->len));
    });
}

static Buf binary(Device& d, int op, const Buf& lhs, const Buf& rhs, bool with_grad, bool add2 = false) {
    same_kind(lhs, rhs, "binary_ew");
    return d.record_binary(op, lhs, rhs, with_grad, add2);
}
Buf Device::add(const Buf& l, const Buf& r) { return binary(*this, SL_ADD, l, r, true); }
Buf Device::add2(const Buf& l, const Buf& r) { return binary(*this, SL_ADD, l, r, true, true); }
Buf Device::sub(const Buf& l, const Buf& r) { return binary(*this, SL_SUB, l, r, true); }
Buf Device::mul(const Buf& l, const Buf& r) { return binary(*this, SL_MUL, l, r, true); }
Buf Device::div(const Buf& l, const Buf& r) { return binary(*this, SL_DIV, l, r, false); }
Buf Device::binary_ew(int op, const Buf& l, const Buf& r) { return binary(*this, op, l, r, false); }

// ------------------------------------------------------------------ unary (custos apply_fn / add_unary_grad †)
Buf Device::record_binary(int op, const Buf& lhs, const Buf& rhs, bool with_grad, bool add2) {
    return record(FNode{1, op, lhs, rhs, Buf(), 0, 0, with_grad, add2});
}
Buf Device::apply_fn(const Buf& x, int unop, double p0, double p1) { return record(FNode{2, unop, x, Buf(), Buf(), p0, p1, false, false}); }
Buf Device::unary_may_grad(const Buf& x, int unop, double p0, double p1) { return record(FNode{2, unop, x, Buf(), Buf(), p0, p1, true, false}); }
Buf Device::square(const Buf& x) { return unary_may_grad(x, SL_UN_SQUARE); }
Buf Device::pow(const Buf& x, double rhs) { return unary_may_grad(x, SL_UN_POW, rhs); }
Buf Device::clip(const Buf& x, double lo, double hi) { return apply_fn(x, SL_UN_CLIP, lo, hi); }
Buf Device::exp(const Buf& x) { return apply_fn(x, SL_UN_EXP); }

// ------------------------------------------------------------------ transpose (src/ops.rs:207-220)
Buf Device::transpose(size_t rows, size_t cols, const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    if (x->len != rows * cols) throw Error(SL_ERR_INVALID_ARG, "transpose: len != rows*cols");
    Buf out = retrieve(x->len, x->dtype);
    check(sl_transpose(ctx_, x->dtype, rows, cols, x->dptr, out->dptr, 0));
    add_grad_fn([this, rows, cols, x, out]() {
        // transpose_grad(cols, rows, x.grad_mut(), out.grad()) — effectively SET on the CPU reference (transpose/cpu.rs:22-23)
        check(sl_transpose(ctx_, x->dtype, cols, rows, grad(out)->dptr, grad(x)->dptr, 0));
    });
    return out;
}

// ------------------------------------------------------------------ gemm (src/ops.rs:250-289)
Buf Device::gemm(size_t m, size_t k, size_t n, const Buf& lhs, const Buf& rhs) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    same_kind(lhs, rhs, "gemm");
    if (lhs->len != m * k || rhs->len != k * n) throw Error(SL_ERR_INVALID_ARG, "gemm: operand length does not match (m,k,n)");
    Buf out = retrieve(m * n, lhs->dtype);
    check(sl_gemm(ctx_, lhs->dtype, m, k, n, lhs->dptr, rhs->dptr, out->dptr, -1));
    add_grad_fn([this, m, k, n, lhs, rhs, out]() {
        // gemm/grad/cpu_stack.rs:35-40: each side guarded by requires_grad(); SET (beta = 0)
        void* lg = lhs->requires_grad ? grad(lhs)->dptr : nullptr;
        void* rg = rhs->requires_grad ? grad(rhs)->dptr : nullptr;
        check(sl_gemm_grad(ctx_, lhs->dtype, m, k, n, lhs->dptr, rhs->dptr, lg, rg, grad(out)->dptr, 0, -1));
    });
    return out;
}

// ------------------------------------------------------------------ row_op (src/ops.rs:323-385)
Buf Device::add_row(size_t rows, size_t cols, const Buf& lhs, const Buf& rhs) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    if (rhs->len != cols) throw Error(SL_ERR_INVALID_ARG, "add_row: rhs.len() != cols");  // row_op/cpu.rs:60
    Buf out = retrieve(lhs->len, lhs->dtype);
    check(sl_add_row(ctx_, lhs->dtype, rows, cols, lhs->dptr, rhs->dptr, out->dptr));
    add_grad_fn([this, rows, cols, lhs, rhs, out]() {
        check(sl_add_row_grad(ctx_, lhs->dtype, rows, cols, grad(lhs)->dptr, grad(rhs)->dptr, grad(out)->dptr));
    });
    return out;
}
void Device::add_row_mut(size_t rows, size_t cols, const Buf& lhs, const Buf& rhs) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    if (rhs->len != cols || lhs->len != rows * cols) throw Error(SL_ERR_INVALID_ARG, "add_row_mut: shape mismatch");
    check(sl_add_row_mut(ctx_, lhs->dtype, rows, cols, lhs->dptr, rhs->dptr));
    add_grad_fn([this, rows, cols, lhs, rhs]() {
        // ops.rs:367-373: add_row_mut_grad(rows, cols, rhs.grad_mut(), lhs.grad())
        check(sl_add_row_mut_grad(ctx_, lhs->dtype, rows, cols, grad(rhs)->dptr, grad(lhs)->dptr));
    });
}

// ------------------------------------------------------------------ reductions (src/ops.rs:450-685)
Buf Device::max_cols(size_t rows, size_t cols, const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    Buf out = retrieve(rows, x->dtype);
    check(sl_max_cols(ctx_, x->dtype, rows, cols, x->dptr, out->dptr, nullptr));
    add_grad_fn([this, rows, cols, x, out]() {
        check(sl_max_cols_grad(ctx_, x->dtype, rows, cols, out->dptr, x->dptr, grad(x)->dptr, grad(out)->dptr));
    });
    return out;
}
Buf Device::max_rows(size_t cols, const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    const size_t rows = cols ? x->len / cols : 0;
    Buf out = retrieve(cols, x->dtype);
    check(sl_max_rows(ctx_, x->dtype, rows, cols, x->dptr, out->dptr, nullptr));
    add_grad_fn([this, rows, cols, x, out]() {
        check(sl_max_rows_grad(ctx_, x->dtype, rows, cols, out->dptr, x->dptr, grad(x)->dptr, grad(out)->dptr));
    });
    return out;
}
Buf Device::sum_rows(size_t cols, const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    const size_t rows = cols ? x->len / cols : 0;
    Buf out = retrieve(cols, x->dtype);
    check(sl_sum_rows(ctx_, x->dtype, rows, cols, x->dptr, out->dptr));
    add_grad_fn([this, rows, cols, x, out]() {  // the closure the reference left commented out (ops.rs:557-565)
        check(sl_sum_rows_grad(ctx_, x->dtype, rows, cols, grad(x)->dptr, grad(out)->dptr));
    });
    return out;
}
Buf Device::sum_cols(size_t cols, const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    const size_t rows = cols ? x->len / cols : 0;  // sum/cpu.rs:55
    Buf out = retrieve(rows, x->dtype);
    check(sl_sum_cols(ctx_, x->dtype, rows, cols, x->dptr, out->dptr));
    add_grad_fn([this, rows, cols, x, out]() {
        check(sl_sum_cols_grad(ctx_, x->dtype, rows, cols, grad(x)->dptr, grad(out)->dptr));
    });
    return out;
}
Buf Device::mean_cols(size_t cols, const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    const size_t rows = cols ? x->len / cols : 0;
    Buf out = retrieve(rows, x->dtype);
    check(sl_mean_cols(ctx_, x->dtype, rows, cols, x->dptr, out->dptr));
    add_grad_fn([this, rows, cols, x, out]() {
        check(sl_mean_cols_grad(ctx_, x->dtype, rows, cols, grad(x)->dptr, grad(out)->dptr));
    });
    return out;
}
Buf Device::mean_rows(size_t cols, const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    const size_t rows = cols ? x->len / cols : 0;
    Buf out = retrieve(cols, x->dtype);
    check(sl_mean_rows(ctx_, x->dtype, rows, cols, x->dptr, out->dptr));
    add_grad_fn([this, rows, cols, x, out]() {
        check(sl_mean_rows_grad(ctx_, x->dtype, rows, cols, grad(x)->dptr, grad(out)->dptr));
    });
    return out;
}

// ------------------------------------------------------------------ diagflat / softmax (src/ops.rs:707-778)
Buf Device::diagflat(const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    Buf out = retrieve(x->len * x->len, x->dtype);
    check(sl_clear(ctx_, out->dptr, out->bytes()));  // only the diagonal is written (diagflat/cpu.rs:42-46): never trust a cached buffer
    check(sl_diagflat(ctx_, x->dtype, x->len, x->dptr, out->dptr));
    add_grad_fn([this, x, out]() { check(sl_diagflat_grad(ctx_, x->dtype, x->len, grad(x)->dptr, grad(out)->dptr)); });
    return out;
}
Buf Device::softmax(size_t samples, size_t features, const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    if (x->len != samples * features) throw Error(SL_ERR_INVALID_ARG, "softmax: len != samples*features");
    Buf out = retrieve(x->len, x->dtype);
    check(sl_softmax(ctx_, x->dtype, samples, features, x->dptr, out->dptr));
    add_grad_fn([this, samples, features, x, out]() {
        check(sl_softmax_grad(ctx_, x->dtype, samples, features, grad(x)->dptr, out->dptr, grad(out)->dptr));
    });
    return out;
}

// ------------------------------------------------------------------ L2 ops the examples call directly
Buf Device::sub_cols(size_t cols, const Buf& lhs, const Buf& rhs) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    Buf out = retrieve(lhs->len, lhs->dtype);
    check(sl_col_op(ctx_, lhs->dtype, SL_SUB, cols ? lhs->len / cols : 0, cols, lhs->dptr, rhs->dptr, out->dptr));
    return out;
}
Buf Device::div_cols(size_t cols, const Buf& lhs, const Buf& rhs) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    Buf out = retrieve(lhs->len, lhs->dtype);
    check(sl_col_op(ctx_, lhs->dtype, SL_DIV, cols ? lhs->len / cols : 0, cols, lhs->dptr, rhs->dptr, out->dptr));
    return out;
}
Buf Device::onehot(const Buf& classes) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    const size_t hc = (size_t)max(classes) + 1;  // onehot/cpu.rs:8
    Buf out = retrieve(classes->len * hc, classes->dtype);
    check(sl_clear(ctx_, out->dptr, out->bytes()));
    check(sl_onehot(ctx_, classes->dtype, classes->len, hc, classes->dptr, out->dptr));
    return out;
}

double Device::scalar_out(int (*fn)(sl_ctx*, int, const void*, size_t, void*), const Buf& x) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    check(fn(ctx_, x->dtype, x->dptr, x->len, scalar_dev_));
    union { float f; double d; int32_t i; } u;
    check(sl_read(ctx_, &u, scalar_dev_, x->dtype == SL_F64 ? 8 : 4));
    return x->dtype == SL_F32 ? (double)u.f : (x->dtype == SL_F64 ? u.d : (double)u.i);
}
double Device::sum(const Buf& x) { return scalar_out(sl_sum, x); }
double Device::mean(const Buf& x) { return scalar_out(sl_mean, x); }
double Device::max(const Buf& x) { return scalar_out(sl_max, x); }

void Device::sgd_step(const Buf& param, double lr) {
    flush_pending();   // an open element-wise chain (fusion) runs before anything that reads or writes buffers
    check(sl_sgd_step(ctx_, param->dtype, param->dptr, grad(param)->dptr, lr, param->len));
}

}  // namespace slh
