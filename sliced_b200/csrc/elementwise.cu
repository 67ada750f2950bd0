// elementwise.cu — family E: element-wise binary / unary ops and their gradients, SGD step, the fused chained graph.
//
// All kernels are HBM-bound: 128-bit coalesced accesses (LDG.128 / STG.128 with L1::no_allocate on read-once
// streams), UNROLL independent packs in flight per thread, grid sized to a multiple of the SM count with a
// block-stride loop.  Compiled with -fmad=false: the reference (Rust) never contracts a*b+c, so neither do we.
//
// Algorithmic bytes per element (f32): binary fwd 12, add/sub grad 20, mul grad 28, unary fwd 8, unary grad 16,
// sgd 12, chained fwd 12, chained bwd 28.
#include "common.cuh"
#include "ew_math.cuh"

namespace {

constexpr int EW_THREADS = 256;

template <typename T, int NR, int NW>
struct EwPtrs {
    const T* in[NR > 0 ? NR : 1];
    T* out[NW];
};

// Generic vectorised map.  F::apply(const T (&in)[NR], T (&out)[NW]) handles ONE element.
// LOAD_OUT: outputs are read-modify-write (ACC kernels) and are preloaded.
template <typename T, int NR, int NW, bool LOAD_OUT, int UNROLL, typename F>
__global__ void __launch_bounds__(EW_THREADS) ew_vec_kernel(EwPtrs<T, NR, NW> p, size_t npacks, F f) {
    constexpr int V = Pack<T>::N;
    const size_t chunk_packs = (size_t)EW_THREADS * UNROLL;
    const size_t nchunks = (npacks + chunk_packs - 1) / chunk_packs;
    for (size_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const size_t base = chunk * chunk_packs + threadIdx.x;
        Pack<T> in[NR > 0 ? NR : 1][UNROLL];
        Pack<T> out[NW][UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const size_t idx = base + (size_t)j * EW_THREADS;
            if (idx < npacks) {
#pragma unroll
                for (int r = 0; r < NR; ++r) in[r][j] = ld_stream(p.in[r] + idx * V);
                if (LOAD_OUT) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) out[w][j] = ld_pack(p.out[w] + idx * V);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const size_t idx = base + (size_t)j * EW_THREADS;
            if (idx < npacks) {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    T a[NR > 0 ? NR : 1];
                    T o[NW];
#pragma unroll
                    for (int r = 0; r < NR; ++r) a[r] = in[r][j].v[e];
#pragma unroll
                    for (int w = 0; w < NW; ++w) o[w] = LOAD_OUT ? out[w][j].v[e] : T(0);
                    f.apply(a, o);
#pragma unroll
                    for (int w = 0; w < NW; ++w) out[w][j].v[e] = o[w];
                }
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    if (LOAD_OUT) st_pack(p.out[w] + idx * V, out[w][j]);
                    else st_stream(p.out[w] + idx * V, out[w][j]);
                }
            }
        }
    }
}

// Scalar variant: tails (n % VEC) and unaligned pointers.
template <typename T, int NR, int NW, bool LOAD_OUT, typename F>
__global__ void __launch_bounds__(EW_THREADS) ew_scalar_kernel(EwPtrs<T, NR, NW> p, size_t begin, size_t n, F f) {
    for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T a[NR > 0 ? NR : 1];
        T o[NW];
#pragma unroll
        for (int r = 0; r < NR; ++r) a[r] = p.in[r][i];
#pragma unroll
        for (int w = 0; w < NW; ++w) o[w] = LOAD_OUT ? p.out[w][i] : T(0);
        f.apply(a, o);
#pragma unroll
        for (int w = 0; w < NW; ++w) p.out[w][i] = o[w];
    }
}

template <typename T, int NR, int NW, bool LOAD_OUT, typename F>
int ew_launch(sl_ctx* ctx, EwPtrs<T, NR, NW> p, size_t n, F f) {
    if (n == 0) return SL_OK;
    constexpr int V = Pack<T>::N;
    constexpr int UNROLL = (NR + NW * (LOAD_OUT ? 2 : 1)) <= 3 ? 4 : 2;
    bool aligned = true;
    for (int r = 0; r < NR; ++r) aligned = aligned && sl_aligned16(p.in[r]);
    for (int w = 0; w < NW; ++w) aligned = aligned && sl_aligned16(p.out[w]);
    size_t npacks = aligned ? n / V : 0;
    if (npacks) {
        const size_t chunk_packs = (size_t)EW_THREADS * UNROLL;
        size_t nchunks = (npacks + chunk_packs - 1) / chunk_packs;
        size_t cap = (size_t)ctx->num_sms * 8;  // 8 resident CTAs of 256 threads per SM
        unsigned grid = (unsigned)(nchunks < cap ? nchunks : cap);
        SL_LAUNCH(ctx, (ew_vec_kernel<T, NR, NW, LOAD_OUT, UNROLL, F>), grid, EW_THREADS, 0, p, npacks, f);
    }
    size_t done = npacks * V;
    if (done < n) {
        size_t rem = n - done;
        size_t blocks = (rem + EW_THREADS - 1) / EW_THREADS;
        size_t cap = (size_t)ctx->num_sms * 8;
        unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
        SL_LAUNCH(ctx, (ew_scalar_kernel<T, NR, NW, LOAD_OUT, F>), grid, EW_THREADS, 0, p, done, n, f);
    }
    return SL_OK;
}

// ------------------------------------------------------------------ functors
template <int OP, typename T>
struct BinaryF {
    __device__ __forceinline__ void apply(const T (&a)[2], T (&o)[1]) const { o[0] = binop<OP>(a[0], a[1]); }
};
// inputs: lhs, rhs, og ; outputs: lg, rg (ACC)
template <int OP, typename T>
struct BinaryGradBothF {
    __device__ __forceinline__ void apply(const T (&a)[3], T (&o)[2]) const {
        o[0] += binop_dl<OP>(a[0], a[1]) * a[2];
        o[1] += binop_dr<OP>(a[0], a[1]) * a[2];
    }
};
template <int OP, typename T, bool LHS>
struct BinaryGradOneF {
    __device__ __forceinline__ void apply(const T (&a)[3], T (&o)[1]) const {
        o[0] += (LHS ? binop_dl<OP>(a[0], a[1]) : binop_dr<OP>(a[0], a[1])) * a[2];
    }
};
// add / sub need only og: 20 B/elem instead of 28
template <int OP, typename T>
struct AddSubGradBothF {
    __device__ __forceinline__ void apply(const T (&a)[1], T (&o)[2]) const {
        o[0] += T(1) * a[0];
        o[1] += (OP == SL_SUB ? -T(1) : T(1)) * a[0];
    }
};
template <typename T, bool NEGATE>
struct AddSubGradOneF {
    __device__ __forceinline__ void apply(const T (&a)[1], T (&o)[1]) const { o[0] += (NEGATE ? -T(1) : T(1)) * a[0]; }
};
template <typename T>
struct AddEwGradOneF {
    __device__ __forceinline__ void apply(const T (&a)[1], T (&o)[1]) const { o[0] += a[0]; }
};
template <typename T>
struct AddEwGradF {
    __device__ __forceinline__ void apply(const T (&a)[1], T (&o)[2]) const {
        o[0] += a[0];
        o[1] += a[0];
    }
};
template <int OP, typename T>
struct UnaryF {
    T p0, p1;
    __device__ __forceinline__ void apply(const T (&a)[1], T (&o)[1]) const { o[0] = unary_f<OP>(a[0], p0, p1); }
};
template <int OP, typename T>
struct UnaryGradF {
    T p0, p1;
    __device__ __forceinline__ void apply(const T (&a)[2], T (&o)[1]) const { o[0] += unary_d<OP>(a[0], p0, p1) * a[1]; }
};
template <typename T>
struct SgdF {
    T lr;
    __device__ __forceinline__ void apply(const T (&a)[1], T (&o)[1]) const { o[0] -= a[0] * lr; }
};
template <typename T>
struct FillF {
    T v;
    __device__ __forceinline__ void apply(const T (&)[1], T (&o)[1]) const { o[0] = v; }
};
// examples/chained_perf.rs:86-90 in the reference's operation order
template <typename T>
struct ChainedFwdF {
    __device__ __forceinline__ void apply(const T (&a)[2], T (&o)[1]) const {
        const T x = a[0], b = a[1];
        const T squared = x * x;
        const T add = b + x;
        const T mul_b = add * b;
        const T mul = squared * x;
        o[0] = mul + mul_b;
    }
};
// the five grad closures of that graph run in reverse registration order, intermediate grads starting at zero
template <typename T>
struct ChainedBwdF {
    __device__ __forceinline__ void apply(const T (&a)[3], T (&o)[2]) const {
        const T x = a[0], b = a[1], og = a[2];
        const T squared = x * x;
        const T add = b + x;
        const T g_mul = T(1) * og;    // out = add(mul, mul_b)
        const T g_mul_b = T(1) * og;
        const T g_squared = x * g_mul;          // mul = mul(squared, x): lhs_grad = rhs*og
        T xg = o[0] + squared * g_mul;          //                         rhs_grad = lhs*og
        const T g_add = b * g_mul_b;            // mul_b = mul(add, b)
        T bg = o[1] + add * g_mul_b;
        bg = bg + T(1) * g_add;                 // add = add(b, x)
        xg = xg + T(1) * g_add;
        xg = xg + (x * T(2)) * g_squared;       // squared = square(x)
        o[0] = xg;
        o[1] = bg;
    }
};

template <typename T>
bool is_float_only_unop(int op) {
    return op == SL_UN_POW || op == SL_UN_TANH || op == SL_UN_SIGMOID || op == SL_UN_EXP || op == SL_UN_LN || op == SL_UN_NEG_LN;
}

#define SL_UNOP_SWITCH(op, OPC, ...)                                         \
    switch (op) {                                                            \
    case SL_UN_SQUARE: { constexpr int OPC = SL_UN_SQUARE; __VA_ARGS__; } break; \
    case SL_UN_POW: { constexpr int OPC = SL_UN_POW; __VA_ARGS__; } break;   \
    case SL_UN_RELU: { constexpr int OPC = SL_UN_RELU; __VA_ARGS__; } break; \
    case SL_UN_TANH: { constexpr int OPC = SL_UN_TANH; __VA_ARGS__; } break; \
    case SL_UN_SIGMOID: { constexpr int OPC = SL_UN_SIGMOID; __VA_ARGS__; } break; \
    case SL_UN_EXP: { constexpr int OPC = SL_UN_EXP; __VA_ARGS__; } break;   \
    case SL_UN_LN: { constexpr int OPC = SL_UN_LN; __VA_ARGS__; } break;     \
    case SL_UN_NEG_LN: { constexpr int OPC = SL_UN_NEG_LN; __VA_ARGS__; } break; \
    case SL_UN_CLIP: { constexpr int OPC = SL_UN_CLIP; __VA_ARGS__; } break; \
    case SL_UN_NEG: { constexpr int OPC = SL_UN_NEG; __VA_ARGS__; } break;   \
    case SL_UN_MUL_SCALAR: { constexpr int OPC = SL_UN_MUL_SCALAR; __VA_ARGS__; } break; \
    case SL_UN_NEG_DIV_SCALAR: { constexpr int OPC = SL_UN_NEG_DIV_SCALAR; __VA_ARGS__; } break; \
    case SL_UN_ADD_SCALAR: { constexpr int OPC = SL_UN_ADD_SCALAR; __VA_ARGS__; } break; \
    default: return sl_set_error(ctx, SL_ERR_INVALID_ARG, "%s: bad unop %d", __func__, (int)(op)); \
    }

#define SL_BINOP_SWITCH(op, OPC, ...)                                        \
    switch (op) {                                                            \
    case SL_ADD: { constexpr int OPC = SL_ADD; __VA_ARGS__; } break;         \
    case SL_SUB: { constexpr int OPC = SL_SUB; __VA_ARGS__; } break;         \
    case SL_MUL: { constexpr int OPC = SL_MUL; __VA_ARGS__; } break;         \
    case SL_DIV: { constexpr int OPC = SL_DIV; __VA_ARGS__; } break;         \
    default: return sl_set_error(ctx, SL_ERR_INVALID_ARG, "%s: bad binop %d", __func__, (int)(op)); \
    }

template <typename T>
int binary_ew_t(sl_ctx* ctx, int op, const void* lhs, const void* rhs, void* out, size_t n) {
    EwPtrs<T, 2, 1> p{{(const T*)lhs, (const T*)rhs}, {(T*)out}};
    SL_BINOP_SWITCH(op, OPC, return (ew_launch<T, 2, 1, false>(ctx, p, n, BinaryF<OPC, T>{})));
    return SL_OK;
}

template <typename T>
int binary_ew_grad_t(sl_ctx* ctx, int op, const void* lhs, const void* rhs, void* lg, void* rg, const void* og, size_t n) {
    if (!lg && !rg) return SL_OK;
    if (lg && lg == rg) {
        // x.mul(x), add(x, x): both gradients are the SAME buffer.  The reference's loop does `lg[i] += ..; rg[i] += ..` one after the
        // other (binary_ew/grad/cpu_stack.rs:54-59), so both contributions land; a two-output kernel would preload the element twice
        // and lose one.  Two one-sided passes give the reference's per-element order: (g + dl*og) + dr*og.
        int rc = binary_ew_grad_t<T>(ctx, op, lhs, rhs, lg, nullptr, og, n);
        return rc != SL_OK ? rc : binary_ew_grad_t<T>(ctx, op, lhs, rhs, nullptr, rg, og, n);
    }
    if ((op == SL_ADD || op == SL_SUB) && lg && rg) {
        EwPtrs<T, 1, 2> p{{(const T*)og}, {(T*)lg, (T*)rg}};
        if (op == SL_ADD) return ew_launch<T, 1, 2, true>(ctx, p, n, AddSubGradBothF<SL_ADD, T>{});
        return ew_launch<T, 1, 2, true>(ctx, p, n, AddSubGradBothF<SL_SUB, T>{});
    }
    if (lg && rg) {
        EwPtrs<T, 3, 2> p{{(const T*)lhs, (const T*)rhs, (const T*)og}, {(T*)lg, (T*)rg}};
        SL_BINOP_SWITCH(op, OPC, return (ew_launch<T, 3, 2, true>(ctx, p, n, BinaryGradBothF<OPC, T>{})));
    } else if (op == SL_ADD || op == SL_SUB) {   // one-sided add / sub: the operands are not needed (and may be NULL)
        EwPtrs<T, 1, 1> p{{(const T*)og}, {(T*)(lg ? lg : rg)}};
        if (op == SL_SUB && !lg) return ew_launch<T, 1, 1, true>(ctx, p, n, AddSubGradOneF<T, true>{});
        return ew_launch<T, 1, 1, true>(ctx, p, n, AddSubGradOneF<T, false>{});
    } else if (lg) {
        EwPtrs<T, 3, 1> p{{(const T*)lhs, (const T*)rhs, (const T*)og}, {(T*)lg}};
        SL_BINOP_SWITCH(op, OPC, return (ew_launch<T, 3, 1, true>(ctx, p, n, BinaryGradOneF<OPC, T, true>{})));
    } else {
        EwPtrs<T, 3, 1> p{{(const T*)lhs, (const T*)rhs, (const T*)og}, {(T*)rg}};
        SL_BINOP_SWITCH(op, OPC, return (ew_launch<T, 3, 1, true>(ctx, p, n, BinaryGradOneF<OPC, T, false>{})));
    }
    return SL_OK;
}

template <typename T>
int unary_t(sl_ctx* ctx, int op, double p0, double p1, const void* x, void* out, size_t n) {
    if (std::is_integral<T>::value && is_float_only_unop<T>(op))
        return sl_set_error(ctx, SL_ERR_UNSUPPORTED, "sl_unary: unop %d needs a float dtype", op);
    EwPtrs<T, 1, 1> p{{(const T*)x}, {(T*)out}};
    SL_UNOP_SWITCH(op, OPC, return (ew_launch<T, 1, 1, false>(ctx, p, n, UnaryF<OPC, T>{(T)p0, (T)p1})));
    return SL_OK;
}

template <typename T>
int unary_grad_t(sl_ctx* ctx, int op, double p0, double p1, const void* x, void* xg, const void* og, size_t n) {
    if (std::is_integral<T>::value && is_float_only_unop<T>(op))
        return sl_set_error(ctx, SL_ERR_UNSUPPORTED, "sl_unary_grad: unop %d needs a float dtype", op);
    EwPtrs<T, 2, 1> p{{(const T*)x, (const T*)og}, {(T*)xg}};
    SL_UNOP_SWITCH(op, OPC, return (ew_launch<T, 2, 1, true>(ctx, p, n, UnaryGradF<OPC, T>{(T)p0, (T)p1})));
    return SL_OK;
}

}  // namespace

extern "C" {

int sl_binary_ew(sl_ctx* ctx, int dtype, int binop, const void* lhs, const void* rhs, void* out, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    SL_REQUIRE(ctx, n == 0 || (lhs && rhs && out), "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return binary_ew_t<T>(ctx, binop, lhs, rhs, out, n));
    return SL_OK;
}

int sl_binary_ew_grad(sl_ctx* ctx, int dtype, int binop, const void* lhs, const void* rhs, void* lhs_grad, void* rhs_grad,
                      const void* out_grad, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, lhs_grad, rhs_grad);
    SL_REQUIRE(ctx, n == 0 || out_grad, "NULL out_grad");
    SL_REQUIRE(ctx, n == 0 || binop == SL_ADD || binop == SL_SUB || (lhs && rhs), "NULL operand");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return binary_ew_grad_t<T>(ctx, binop, lhs, rhs, lhs_grad, rhs_grad, out_grad, n));
    return SL_OK;
}

int sl_add_ew_grad(sl_ctx* ctx, int dtype, void* lhs_grad, void* rhs_grad, const void* out_grad, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, lhs_grad, rhs_grad);
    SL_REQUIRE(ctx, n == 0 || (lhs_grad && rhs_grad && out_grad), "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        if (lhs_grad == rhs_grad) {   // add2(x, x): the same buffer takes both `+= og` in turn (grad/cpu_stack.rs:80-88)
            EwPtrs<T, 1, 1> q{{(const T*)out_grad}, {(T*)lhs_grad}};
            int rc = ew_launch<T, 1, 1, true>(ctx, q, n, AddEwGradOneF<T>{});
            return rc != SL_OK ? rc : ew_launch<T, 1, 1, true>(ctx, q, n, AddEwGradOneF<T>{});
        }
        EwPtrs<T, 1, 2> p{{(const T*)out_grad}, {(T*)lhs_grad, (T*)rhs_grad}};
        return ew_launch<T, 1, 2, true>(ctx, p, n, AddEwGradF<T>{});
    });
    return SL_OK;
}

int sl_unary(sl_ctx* ctx, int dtype, int unop, double p0, double p1, const void* x, void* out, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    SL_REQUIRE(ctx, n == 0 || (x && out), "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return unary_t<T>(ctx, unop, p0, p1, x, out, n));
    return SL_OK;
}

int sl_unary_grad(sl_ctx* ctx, int dtype, int unop, double p0, double p1, const void* x, void* x_grad, const void* out_grad,
                  size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, x_grad);
    SL_REQUIRE(ctx, n == 0 || (x && x_grad && out_grad), "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return unary_grad_t<T>(ctx, unop, p0, p1, x, x_grad, out_grad, n));
    return SL_OK;
}

int sl_sgd_step(sl_ctx* ctx, int dtype, void* w, const void* g, double lr, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, w);
    SL_REQUIRE(ctx, n == 0 || (w && g), "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        EwPtrs<T, 1, 1> p{{(const T*)g}, {(T*)w}};
        return ew_launch<T, 1, 1, true>(ctx, p, n, SgdF<T>{(T)lr});
    });
    return SL_OK;
}

int sl_fill(sl_ctx* ctx, int dtype, void* dst_dev, double value, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, dst_dev);
    SL_REQUIRE(ctx, n == 0 || dst_dev, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        EwPtrs<T, 0, 1> p{{nullptr}, {(T*)dst_dev}};
        return ew_launch<T, 0, 1, false>(ctx, p, n, FillF<T>{(T)value});
    });
    return SL_OK;
}

int sl_chained_fwd(sl_ctx* ctx, int dtype, const void* x, const void* b, void* out, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    SL_REQUIRE(ctx, n == 0 || (x && b && out), "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        EwPtrs<T, 2, 1> p{{(const T*)x, (const T*)b}, {(T*)out}};
        return ew_launch<T, 2, 1, false>(ctx, p, n, ChainedFwdF<T>{});
    });
    return SL_OK;
}

int sl_chained_bwd(sl_ctx* ctx, int dtype, const void* x, const void* b, void* x_grad, void* b_grad, const void* out_grad,
                   size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, x_grad, b_grad);
    SL_REQUIRE(ctx, n == 0 || (x && b && x_grad && b_grad && out_grad), "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        EwPtrs<T, 3, 2> p{{(const T*)x, (const T*)b, (const T*)out_grad}, {(T*)x_grad, (T*)b_grad}};
        return ew_launch<T, 3, 2, true>(ctx, p, n, ChainedBwdF<T>{});
    });
    return SL_OK;
}

}  // extern "C"
