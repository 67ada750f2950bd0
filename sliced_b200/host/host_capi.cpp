// host_capi.cpp — C bridge over the C++ host layer (slh::Device / Buf / Matrix / Mlp) so that Python (tests, bench)
// and C programs can drive it.  Handles are opaque; errors are returned as NULL / negative codes + slh_last_error().
#include <cstring>
#include <string>

#include "mlp.hpp"

using namespace slh;

namespace {
thread_local std::string g_err;
struct BufHandle {
    Buf b;
};
template <typename F>
auto guard(F&& f, decltype(f()) on_error) -> decltype(f()) {
    try {
        return f();
    } catch (const Error& e) {
        g_err = e.what();
    } catch (const std::exception& e) {
        g_err = e.what();
    }
    return on_error;
}
BufHandle* wrap_handle(const Buf& b) { return new BufHandle{b}; }
}  // namespace

extern "C" {

typedef struct slh_device slh_device;
typedef struct slh_buffer slh_buffer;
typedef struct slh_mlp slh_mlp;

const char* slh_last_error(void) { return g_err.c_str(); }

slh_device* slh_device_new(int device_index, int cached) {
    return guard([&]() { return (slh_device*)new Device(device_index, cached != 0); }, (slh_device*)nullptr);
}
slh_device* slh_device_new_on_stream(int device_index, int cached, void* stream) {
    return guard([&]() { return (slh_device*)new Device(device_index, cached != 0, stream, true); }, (slh_device*)nullptr);
}
void slh_device_free(slh_device* d) { delete (Device*)d; }
void* slh_device_ctx(slh_device* d) { return ((Device*)d)->ctx(); }
int slh_device_sync(slh_device* d) { return guard([&]() { ((Device*)d)->sync(); return 0; }, -1); }
void slh_range_begin(slh_device* d) { ((Device*)d)->range_begin(); }
int slh_zero_grad(slh_device* d) { return guard([&]() { ((Device*)d)->zero_grad(); return 0; }, -1); }
int slh_tape_len(slh_device* d) { return (int)((Device*)d)->tape_len(); }
int slh_n_grads(slh_device* d) { return (int)((Device*)d)->n_grads(); }
int slh_set_fusion(slh_device* d, int on) { return guard([&]() { ((Device*)d)->set_fusion(on != 0); return 0; }, -1); }
int slh_flush(slh_device* d) { return guard([&]() { ((Device*)d)->flush_pending(); return 0; }, -1); }
int slh_fused_groups(slh_device* d) { return (int)((Device*)d)->fused_groups(); }
int slh_unfused_groups(slh_device* d) { return (int)((Device*)d)->unfused_groups(); }
void slh_set_tape_enabled(slh_device* d, int on) { ((Device*)d)->set_tape_enabled(on != 0); }
int slh_set_gemm_mode(slh_device* d, int mode) { return guard([&]() { ((Device*)d)->set_gemm_mode(mode); return 0; }, -1); }

slh_buffer* slh_buffer_new(slh_device* d, size_t len, int dtype) {
    return guard([&]() { return (slh_buffer*)wrap_handle(((Device*)d)->buffer(len, dtype)); }, (slh_buffer*)nullptr);
}
slh_buffer* slh_buffer_from_host(slh_device* d, const void* host, size_t len, int dtype) {
    return guard([&]() { return (slh_buffer*)wrap_handle(((Device*)d)->from_host(host, len, dtype)); }, (slh_buffer*)nullptr);
}
slh_buffer* slh_buffer_wrap(slh_device* d, void* dptr, size_t len, int dtype) {
    return guard([&]() { return (slh_buffer*)wrap_handle(((Device*)d)->wrap(dptr, len, dtype)); }, (slh_buffer*)nullptr);
}
void slh_buffer_release(slh_buffer* b) { delete (BufHandle*)b; }
size_t slh_buffer_len(slh_buffer* b) { return ((BufHandle*)b)->b->len; }
int slh_buffer_dtype(slh_buffer* b) { return ((BufHandle*)b)->b->dtype; }
void* slh_buffer_dptr(slh_buffer* b) { return ((BufHandle*)b)->b->dptr; }
uint64_t slh_buffer_id(slh_buffer* b) { return ((BufHandle*)b)->b->id; }
void slh_buffer_set_requires_grad(slh_buffer* b, int on) { ((BufHandle*)b)->b->requires_grad = on != 0; }
int slh_buffer_requires_grad(slh_buffer* b) { return ((BufHandle*)b)->b->requires_grad ? 1 : 0; }
int slh_buffer_read(slh_buffer* b, void* host) {
    return guard([&]() { auto& x = ((BufHandle*)b)->b; x->dev->read(x, host); return 0; }, -1);
}
int slh_buffer_write(slh_buffer* b, const void* host) {
    return guard([&]() { auto& x = ((BufHandle*)b)->b; x->dev->write(x, host); return 0; }, -1);
}
slh_buffer* slh_grad(slh_buffer* b) {
    return guard([&]() { auto& x = ((BufHandle*)b)->b; return (slh_buffer*)wrap_handle(x->dev->grad(x)); }, (slh_buffer*)nullptr);
}
int slh_backward(slh_buffer* out) {
    return guard([&]() { auto& x = ((BufHandle*)out)->b; x->dev->backward(x); return 0; }, -1);
}
int slh_backward_with(slh_buffer* out, slh_buffer* seed) {
    return guard([&]() { auto& x = ((BufHandle*)out)->b; x->dev->backward_with(x, ((BufHandle*)seed)->b); return 0; }, -1);
}

// Generic op dispatcher: name selects the L3 / L4 method; bufs / dims / scalars are its arguments in declaration order.
slh_buffer* slh_op(slh_device* dh, const char* name, slh_buffer** bufs, int nbufs, const size_t* dims, int ndims, const double* sc, int nsc) {
    return guard([&]() -> slh_buffer* {
        Device& d = *(Device*)dh;
        auto B = [&](int i) -> const Buf& {
            if (i >= nbufs) throw Error(SL_ERR_INVALID_ARG, std::string(name) + ": missing buffer argument");
            return ((BufHandle*)bufs[i])->b;
        };
        auto D = [&](int i) -> size_t {
            if (i >= ndims) throw Error(SL_ERR_INVALID_ARG, std::string(name) + ": missing dimension argument");
            return dims[i];
        };
        auto S = [&](int i) -> double {
            if (i >= nsc) throw Error(SL_ERR_INVALID_ARG, std::string(name) + ": missing scalar argument");
            return sc[i];
        };
        const std::string n(name);
        Buf r;
        if (n == "add") r = d.add(B(0), B(1));
        else if (n == "add2") r = d.add2(B(0), B(1));
        else if (n == "sub") r = d.sub(B(0), B(1));
        else if (n == "mul") r = d.mul(B(0), B(1));
        else if (n == "div") r = d.div(B(0), B(1));
        else if (n == "square") r = d.square(B(0));
        else if (n == "pow") r = d.pow(B(0), S(0));
        else if (n == "transpose") r = d.transpose(D(0), D(1), B(0));
        else if (n == "gemm") r = d.gemm(D(0), D(1), D(2), B(0), B(1));
        else if (n == "add_row") r = d.add_row(D(0), D(1), B(0), B(1));
        else if (n == "add_row_mut") { d.add_row_mut(D(0), D(1), B(0), B(1)); r = B(0); }
        else if (n == "clip") r = d.clip(B(0), S(0), S(1));
        else if (n == "exp") r = d.exp(B(0));
        else if (n == "max_cols") r = d.max_cols(D(0), D(1), B(0));
        else if (n == "max_rows") r = d.max_rows(D(0), B(0));
        else if (n == "sum_rows") r = d.sum_rows(D(0), B(0));
        else if (n == "sum_cols") r = d.sum_cols(D(0), B(0));
        else if (n == "mean_cols") r = d.mean_cols(D(0), B(0));
        else if (n == "mean_rows") r = d.mean_rows(D(0), B(0));
        else if (n == "diagflat") r = d.diagflat(B(0));
        else if (n == "softmax") r = d.softmax(D(0), D(1), B(0));
        else if (n == "relu") r = d.unary_may_grad(B(0), SL_UN_RELU);
        else if (n == "tanh") r = d.unary_may_grad(B(0), SL_UN_TANH);
        else if (n == "sigmoid") r = d.unary_may_grad(B(0), SL_UN_SIGMOID);
        else if (n == "apply_fn") r = d.apply_fn(B(0), (int)D(0), nsc > 0 ? sc[0] : 0.0, nsc > 1 ? sc[1] : 0.0);
        else if (n == "unary_may_grad") r = d.unary_may_grad(B(0), (int)D(0), nsc > 0 ? sc[0] : 0.0, nsc > 1 ? sc[1] : 0.0);
        else if (n == "sub_cols") r = d.sub_cols(D(0), B(0), B(1));
        else if (n == "div_cols") r = d.div_cols(D(0), B(0), B(1));
        else if (n == "onehot") r = d.onehot(B(0));
        else throw Error(SL_ERR_UNSUPPORTED, "slh_op: unknown op '" + n + "'");
        return (slh_buffer*)wrap_handle(r);
    }, (slh_buffer*)nullptr);
}

int slh_scalar_op(slh_device* dh, const char* name, slh_buffer* b, double* out) {
    return guard([&]() {
        Device& d = *(Device*)dh;
        const std::string n(name);
        const Buf& x = ((BufHandle*)b)->b;
        if (n == "sum") *out = d.sum(x);
        else if (n == "mean") *out = d.mean(x);
        else if (n == "max") *out = d.max(x);
        else throw Error(SL_ERR_UNSUPPORTED, "slh_scalar_op: unknown op '" + n + "'");
        return 0;
    }, -1);
}

int slh_sgd_step(slh_buffer* param, double lr) {
    return guard([&]() { auto& x = ((BufHandle*)param)->b; x->dev->sgd_step(x, lr); return 0; }, -1);
}

// ---------------------------------------------------------------- Mlp (examples/nn.rs, examples/sine_net.rs)
slh_mlp* slh_mlp_new(slh_device* d, int n_dims, const size_t* dims, int loss_kind) {
    return guard([&]() { return (slh_mlp*)new Mlp(*(Device*)d, std::vector<size_t>(dims, dims + n_dims), (LossKind)loss_kind); }, (slh_mlp*)nullptr);
}
void slh_mlp_free(slh_mlp* m) { delete (Mlp*)m; }
size_t slh_mlp_n_params(slh_mlp* m) { return ((Mlp*)m)->n_params(); }
void* slh_mlp_metrics_dptr(slh_mlp* m) { return ((Mlp*)m)->metrics_dptr(); }
slh_buffer* slh_mlp_weights(slh_mlp* m, int layer) { return guard([&]() { ((Mlp*)m)->flush(); return (slh_buffer*)wrap_handle(((Mlp*)m)->layer(layer).weights.data); }, (slh_buffer*)nullptr); }
slh_buffer* slh_mlp_bias(slh_mlp* m, int layer) { return guard([&]() { ((Mlp*)m)->flush(); return (slh_buffer*)wrap_handle(((Mlp*)m)->layer(layer).bias.data); }, (slh_buffer*)nullptr); }
slh_buffer* slh_mlp_grad_bucket(slh_mlp* m) { return guard([&]() { ((Mlp*)m)->flush(); return (slh_buffer*)wrap_handle(((Mlp*)m)->grad_bucket()); }, (slh_buffer*)nullptr); }
slh_buffer* slh_mlp_params(slh_mlp* m) { return guard([&]() { ((Mlp*)m)->flush(); return (slh_buffer*)wrap_handle(((Mlp*)m)->params()); }, (slh_buffer*)nullptr); }

int slh_mlp_forward_backward(slh_mlp* m, slh_buffer* x, slh_buffer* y, slh_buffer* labels, size_t batch, size_t grad_rows, int want_metrics,
                             double* loss_sum, long long* correct) {
    return guard([&]() {
        Mlp& mlp = *(Mlp*)m;   // the fused step when set_fused(true) (softmax-cce loss only), else the op-by-op tape
        const Buf lb = labels ? ((BufHandle*)labels)->b : Buf();
        StepResult r = mlp.fused_active() ? mlp.forward_backward_fused(((BufHandle*)x)->b, ((BufHandle*)y)->b, lb, batch, grad_rows, want_metrics != 0)
                                          : mlp.forward_backward(((BufHandle*)x)->b, ((BufHandle*)y)->b, lb, batch, grad_rows, want_metrics != 0);
        if (loss_sum) *loss_sum = r.loss_sum;
        if (correct) *correct = r.correct;
        return 0;
    }, -1);
}
void slh_mlp_set_fused(slh_mlp* m, int on) { ((Mlp*)m)->set_fused(on != 0); }
int slh_mlp_set_deferred(slh_mlp* m, int on) { return guard([&]() { ((Mlp*)m)->set_deferred(on != 0); return 0; }, -1); }
int slh_mlp_flush(slh_mlp* m) { return guard([&]() { ((Mlp*)m)->flush(); return 0; }, -1); }
int slh_mlp_allreduce_grads(slh_mlp* m) { return guard([&]() { ((Mlp*)m)->allreduce_grads(); return 0; }, -1); }
int slh_mlp_sgd(slh_mlp* m, double lr) { return guard([&]() { ((Mlp*)m)->sgd(lr); return 0; }, -1); }
int slh_mlp_step(slh_mlp* m, slh_buffer* x, slh_buffer* y, slh_buffer* labels, size_t batch, size_t grad_rows, double lr, int want_metrics,
                 double* loss_sum, long long* correct) {
    return guard([&]() {
        StepResult r = ((Mlp*)m)->step(((BufHandle*)x)->b, ((BufHandle*)y)->b, labels ? ((BufHandle*)labels)->b : Buf(), batch, grad_rows, lr,
                                       want_metrics != 0);
        if (loss_sum) *loss_sum = r.loss_sum;
        if (correct) *correct = r.correct;
        return 0;
    }, -1);
}
int slh_mlp_step_replay(slh_mlp* m, slh_buffer* x, slh_buffer* y, slh_buffer* labels, size_t batch, size_t grad_rows, double lr, int want_metrics,
                        double* loss_sum, long long* correct) {
    return guard([&]() {
        StepResult r = ((Mlp*)m)->step_replay(((BufHandle*)x)->b, ((BufHandle*)y)->b, labels ? ((BufHandle*)labels)->b : Buf(), batch, grad_rows, lr,
                                              want_metrics != 0);
        if (loss_sum) *loss_sum = r.loss_sum;
        if (correct) *correct = r.correct;
        return 0;
    }, -1);
}
slh_buffer* slh_mlp_predict(slh_mlp* m, slh_buffer* x, size_t batch) {
    return guard([&]() { return (slh_buffer*)wrap_handle(((Mlp*)m)->predict(((BufHandle*)x)->b, batch).data); }, (slh_buffer*)nullptr);
}

}  // extern "C"
