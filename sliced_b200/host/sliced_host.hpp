// sliced_host.hpp — host-side mirror (C++) of the reference's device-facing layers, sitting ABOVE the C ABI:
//
//   reference (Rust)                                   here
//   ------------------------------------------------   -----------------------------------------------------------
//   custos `CUDA<Autograd<Cached<Base>>>` device †      slh::Device  (ctx + tape + gradient map + Cached/Cursor module)
//   custos `Buffer<T, D>` †                             slh::Buf     (shared handle: id, dtype, len, device pointer)
//   sliced L3 `*MayGrad` traits     src/ops.rs          Device::add/sub/mul/square/pow/transpose/gemm/add_row/... :
//                                                       forward through the C ABI + the SAME grad closure pushed on the tape
//   sliced L4 `Matrix`              src/matrix.rs       slh::Matrix  (shape-carrying wrapper, relu/tanh/sigmoid/softmax/...)
//
// The reference's toolchain (rustc) is absent from this image, so this layer is C++ instead of a Rust `cuda.rs`
// per op; INTEGRATION.md shows the Rust binding a maintainer would add.  Every method cites the reference lines it
// mirrors.  Nothing here computes on the host: all arithmetic is a C-ABI call into the CUDA library.
#pragma once

#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/sliced_b200.h"

namespace slh {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class Device;

struct BufferImpl {
    uint64_t id = 0;
    int dtype = SL_F32;
    size_t len = 0;
    void* dptr = nullptr;
    bool requires_grad = true;  // custos buffers take part in autograd unless `.no_grad()` (examples/nn.rs:170,177)
    bool owns = true;
    bool pending = false;       // fusion: recorded in an open element-wise chain, not computed yet (dptr may still be null)
    std::shared_ptr<BufferImpl> backing;   // fusion: the (cached) allocation a materialised chain output lives in
    Device* dev = nullptr;
    ~BufferImpl();
    size_t bytes() const { return len * (dtype == SL_F64 ? 8 : 4); }
};
using Buf = std::shared_ptr<BufferImpl>;

class Device {
  public:
    // cached = custos `Cached` module: `retrieve` hands back the same buffer for the same call-site position in every
    // iteration of a `range()` loop (examples/nn.rs:156,184) -> zero allocations after the first iteration.
    explicit Device(int device_index = 0, bool cached = false, void* cuda_stream = nullptr, bool borrow_stream = false);
    ~Device();
    Device(const Device&) = delete;

    sl_ctx* ctx() const { return ctx_; }
    void check(int rc) const;

    // ---- custos Alloc / Buffer::from / Read / WriteBuf †
    Buf buffer(size_t len, int dtype);                              // zero-initialised, like custos `Buffer::new`
    Buf from_host(const void* host, size_t len, int dtype);         // Buffer::from((&device, slice))
    Buf wrap(void* dptr, size_t len, int dtype);                    // borrow an existing device allocation
    void read(const Buf& b, void* host) const;                      // Buffer::read
    void write(const Buf& b, const void* host);                     // WriteBuf::write
    void sync() const;

    // ---- custos Cached + Cursor †: `for _ in device.range(..)` rewinds the cursor (examples/nn.rs:184)
    void range_begin() { cursor_ = 0; }
    Buf retrieve(size_t len, int dtype);

    // ---- custos Autograd †: tape + gradients
    Buf grad(const Buf& b);                                         // Buffer::grad / grad_mut: lazily zero-allocated
    bool has_grad(const Buf& b) const { return grads_.count(b->id) != 0; }
    size_t n_grads() const { return grads_.size(); }
    void drop_grad(uint64_t id);                                    // OnDropBuffer †: a buffer's gradient dies with the buffer
    bool cached() const { return cached_; }
    void bind_grad(const Buf& b, const Buf& g) { grads_[b->id] = g; }  // place b's gradient in caller-owned memory (DP bucket)
    void zero_grad();                                               // gradients_mut().zero_grad() (examples/nn.rs:186-188)
    void backward(const Buf& out);                                  // seeds ones, runs the tape in reverse, clears it
    void backward_with(const Buf& out, const Buf& seed);            // examples/nn.rs:233
    void add_grad_fn(std::function<void()> f) { if (tape_enabled_) tape_.push_back(std::move(f)); }
    size_t tape_len() const { return tape_.size(); }
    void set_tape_enabled(bool on) { tape_enabled_ = on; }
    void set_gemm_mode(int mode) { check(sl_ctx_set_gemm_mode(ctx_, mode)); }

    // ---- custos `Lazy` + `optimize()` analogue for element-wise chains (examples/chained_perf.rs:114, sine_net.rs:178-233): with
    // fusion on, binary / unary element-wise ops are RECORDED instead of launched; the chain runs as one sl_fused_chain launch when
    // something needs its values (any other op, read, backward), materialising only the results somebody still holds a handle to, and
    // its grad closures become one fused backward launch.  Chains that do not fit the interpreter run op by op as before.
    void set_fusion(bool on) { flush_pending(); fusion_ = on; }
    bool fusion() const { return fusion_; }
    void flush_pending();
    size_t fused_groups() const { return fused_groups_; }       // chains executed as one launch so far
    size_t unfused_groups() const { return unfused_groups_; }   // recorded chains that had to run op by op

    // ---- L3: src/ops.rs
    Buf add(const Buf& lhs, const Buf& rhs);                        // BinaryOpsMayGrad::add   ops.rs:115-132
    Buf add2(const Buf& lhs, const Buf& rhs);                       // BinaryOpsMayGrad::add2  ops.rs:173-186
    Buf sub(const Buf& lhs, const Buf& rhs);                        // ops.rs:134-151
    Buf mul(const Buf& lhs, const Buf& rhs);                        // ops.rs:153-171
    Buf div(const Buf& lhs, const Buf& rhs);                        // BinaryElementWise::div (no grad in the reference; binary_ew/mod.rs:85-90)
    Buf binary_ew(int binop, const Buf& lhs, const Buf& rhs);       // L2 BinaryElementWise::{add,sub,mul,div}: forward only, nothing on the tape
    Buf record_binary(int binop, const Buf& lhs, const Buf& rhs, bool with_grad, bool add2);  // (shared body of the five above)
    Buf square(const Buf& x);                                       // SquareMayGrad ops.rs:25-45
    Buf pow(const Buf& x, double rhs);                              // PowMayGrad    ops.rs:64-77
    Buf transpose(size_t rows, size_t cols, const Buf& x);          // TransposeMayGrad ops.rs:207-220
    Buf gemm(size_t m, size_t k, size_t n, const Buf& lhs, const Buf& rhs);  // GemmMayGrad ops.rs:250-289
    Buf add_row(size_t rows, size_t cols, const Buf& lhs, const Buf& rhs);   // RowOpMayGrad ops.rs:323-356
    void add_row_mut(size_t rows, size_t cols, const Buf& lhs, const Buf& rhs);  // ops.rs:358-385
    Buf clip(const Buf& x, double lo, double hi);                   // Clip ops.rs:419-425
    Buf exp(const Buf& x);                                          // Exp  ops.rs:435-440
    Buf max_cols(size_t rows, size_t cols, const Buf& x);           // MaxColsMayGrad ops.rs:471-491
    Buf max_rows(size_t cols, const Buf& x);                        // MaxRowsMayGrad ops.rs:515-534
    Buf sum_rows(size_t cols, const Buf& x);                        // SumRowsMayGrad ops.rs:553-567 (reference: unimplemented!(); spec = tests/test_sum_rows.rs)
    Buf sum_cols(size_t cols, const Buf& x);                        // SumColsMayGrad ops.rs:592-611
    Buf mean_cols(size_t cols, const Buf& x);                       // MeanColsMayGrad ops.rs:635-651
    Buf mean_rows(size_t cols, const Buf& x);                       // MeanRowsMayGrad ops.rs:675-684
    Buf diagflat(const Buf& x);                                     // DiagflatMayGrad ops.rs:707-724
    Buf softmax(size_t samples, size_t features, const Buf& x);     // SoftmaxMayGrad ops.rs:751-778
    // custos ApplyFunction::apply_fn / UnaryGrad::add_unary_grad † for the closures sliced uses
    Buf apply_fn(const Buf& x, int unop, double p0 = 0, double p1 = 0);                       // no grad registered
    Buf unary_may_grad(const Buf& x, int unop, double p0 = 0, double p1 = 0);                 // + add_unary_grad closure
    // L2 ops without a MayGrad wrapper that the examples call directly
    Buf sub_cols(size_t cols, const Buf& lhs, const Buf& rhs);      // ColOp::sub_cols col_op/mod.rs:33-43
    Buf div_cols(size_t cols, const Buf& lhs, const Buf& rhs);      // ColOp::div_cols col_op/mod.rs:47-57
    Buf onehot(const Buf& classes);                                 // Onehot::onehot onehot/cpu.rs:7-16
    double sum(const Buf& x);                                       // Sum::sum  sum/cpu.rs:18-20
    double mean(const Buf& x);                                      // Mean::mean mean/cpu.rs:7-9
    double max(const Buf& x);                                       // Max::max  max/cpu.rs:23-25
    void sgd_step(const Buf& param, double lr);                     // SGD::step examples/nn.rs:108-119 (param -= grad * lr)

  private:
    struct FNode {   // one recorded element-wise op
        int kind;    // 1 binary (sl_binop), 2 unary (sl_unop)
        int op;
        Buf a, b, dst;
        double p0, p1;
        bool with_grad, add2;
    };
    Buf record(FNode n);
    void run_node_unfused(const FNode& n);
    bool fusion_ = false;
    std::vector<FNode> pending_;
    size_t fused_groups_ = 0, unfused_groups_ = 0;
    Buf new_buffer(size_t len, int dtype, bool zero);
    double scalar_out(int (*fn)(sl_ctx*, int, const void*, size_t, void*), const Buf& x);

    sl_ctx* ctx_ = nullptr;
    bool cached_ = false;
    bool tape_enabled_ = true;
    uint64_t next_id_ = 1;
    std::vector<std::function<void()>> tape_;
    std::unordered_map<uint64_t, Buf> grads_;
    std::vector<Buf> cache_;
    size_t cursor_ = 0;
    void* scalar_dev_ = nullptr;
    bool tearing_down_ = false;
};

// sliced L4: src/matrix.rs:21-25 — (Buffer, rows, cols)
struct Matrix {
    Buf data;
    size_t rows = 0, cols = 0;
    Matrix() = default;
    Matrix(Buf d, size_t r, size_t c) : data(std::move(d)), rows(r), cols(c) {
        if (data->len != r * c) throw Error(SL_ERR_INVALID_ARG, "Matrix: data.len() != rows * cols");  // matrix/impl_from.rs:8
    }
    Device& device() const { return *data->dev; }
    Matrix T() const { return {device().transpose(rows, cols, data), cols, rows}; }                 // matrix.rs:98-107
    Matrix gemm(const Matrix& rhs) const { return {device().gemm(rows, cols, rhs.cols, data, rhs.data), rows, rhs.cols}; }  // :111-122
    Matrix add(const Matrix& rhs) const { return {device().add(data, rhs.data), rows, cols}; }      // :127-132
    Matrix sub(const Matrix& rhs) const { return {device().sub(data, rhs.data), rows, cols}; }      // `&out - &y` sine_net.rs:150
    Matrix mul(const Matrix& rhs) const { return {device().mul(data, rhs.data), rows, cols}; }      // :137-142
    Matrix add_row(const Matrix& rhs) const { return {device().add_row(rows, cols, data, rhs.data), rows, cols}; }  // :146-156
    void add_row_mut(const Matrix& rhs) { device().add_row_mut(rows, cols, data, rhs.data); }       // :160-165
    Matrix relu() const { return {device().unary_may_grad(data, SL_UN_RELU), rows, cols}; }         // :169-190
    Matrix tanh() const { return {device().unary_may_grad(data, SL_UN_TANH), rows, cols}; }         // :206-230
    Matrix sigmoid() const { return {device().unary_may_grad(data, SL_UN_SIGMOID), rows, cols}; }   // :234-262
    Matrix squared() const { return {device().square(data), rows, cols}; }                          // :275-288
    Matrix pow(double rhs) const { return {device().pow(data, rhs), rows, cols}; }                  // :292-297
    Matrix sum_cols() const { return {device().sum_cols(cols, data), rows, 1}; }                    // :329-334
    Matrix l2_norm_cols() const { return squared().sum_cols().pow(0.5); }                           // :337-355
    Matrix diagflat() const { return {device().diagflat(data), rows, rows}; }                       // :359-364
    Matrix softmax() const { return {device().softmax(rows, cols, data), rows, cols}; }             // :368-378
};

}  // namespace slh
