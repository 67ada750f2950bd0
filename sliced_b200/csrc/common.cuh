// common.cuh — shared plumbing of libsliced_b200.so: context object, error handling, dtype dispatch,
// 128-bit streaming load/store helpers.  B200 / sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <type_traits>

#include "../../include/sliced_b200.h"

struct sl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int num_sms = 148;
    int gemm_mode = SL_GEMM_3XF16;
    uint64_t launches = 0;
    std::string last_error;
    // grow-only scratch (reduction partials, gemm operand planes); owned by the ctx, never user visible
    void* ws = nullptr;
    size_t ws_bytes = 0;
    void* ws2 = nullptr;  // second scratch region (gemm planes live here so reductions can run concurrently in-order)
    size_t ws2_bytes = 0;
    // optional per-launch event timing of the gemm MMA kernel (sl_ctx_profile_begin / _end)
    bool profiling = false;
    struct ProfRec { cudaEvent_t a, b; double flops; const char* name; };
    std::vector<ProfRec> prof;
    bool profiling_all = false;  // SLICED_PROFILE_ALL / sl_ctx_profile_report: time every kernel launch, not only the MMA kernels
    // operand-plane reuse (sl_gemm_scope_begin / _end): hi/lo TF32 planes of gemm operands kept for later gemms on the same buffer
    struct PlaneEntry { const void* src; size_t elems; float* hi; float* lo; size_t cap_bytes; bool valid; };
    std::vector<PlaneEntry> plane_cache;
    size_t plane_cursor = 0;
    bool plane_scope = false;
    // 3xFP16 planes of a K-contiguous operand [rows x cols] (scaled per ROW), kept inside a scope: a later gemm of the scope that
    // contracts over the ROWS of the same buffer consumes them as its MN-major operand and folds the row scales into the other
    // operand's split instead of splitting this buffer a second time per column
    struct RowPlanes { const void* src; size_t rows, cols; void* hi; void* lo; float* inv; size_t cap_bytes, cap_rows; bool valid; };
    std::vector<RowPlanes> rowplane_cache;
    size_t rowplane_cursor = 0;
    // sl_mlp_small_step: function attributes are per device, so their one-time setup is remembered per context
    bool mlp_small_attr = false;
    int mlp_small_can16 = -1;   // -1 unknown | 0 | 1: a 16-CTA cluster of the kernel can be placed on this device
    // second stream for host->device prefetch (sl_write_prefetch), created lazily
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done = nullptr, compute_done = nullptr;
    // NCCL (dlopen'ed lazily)
    void* nccl_comm = nullptr;
    int nranks = 1;
    int rank = 0;
    cudaStream_t comm_stream = nullptr;  // overlapped exchange (sl_allreduce_sum_async)
    cudaEvent_t comm_ready = nullptr, comm_done = nullptr;
    bool comm_pending = false;
    std::vector<cudaEvent_t> comm_events;   // one per overlapped exchange issued since the last sl_comm_wait (sl_comm_wait_n)
    size_t comm_issued = 0;
};

int sl_set_error(sl_ctx* ctx, int code, const char* fmt, ...);

// Plane-cache coherence (sl_gemm_scope_begin / _end): EVERY entry point of the library that writes device memory reports its output
// pointers here, so cached operand planes / column scales derived from a buffer die as soon as anything overwrites (part of) it.
static inline void sl_note_write(sl_ctx* ctx, const void* p) {
    if (!ctx || !ctx->plane_scope || !p) return;
    const char* q = (const char*)p;
    for (auto& e : ctx->plane_cache)
        if (e.valid && q >= (const char*)e.src && q < (const char*)e.src + e.elems * 4) e.valid = false;
    for (auto& e : ctx->rowplane_cache)
        if (e.valid && q >= (const char*)e.src && q < (const char*)e.src + e.rows * e.cols * 4) e.valid = false;
}
template <typename... P>
static inline void sl_note_writes(sl_ctx* ctx, P... ptrs) {
    const void* a[] = {(const void*)ptrs...};
    for (const void* p : a) sl_note_write(ctx, p);
}
int sl_ws_reserve(sl_ctx* ctx, size_t bytes, void** out);
int sl_ws2_reserve(sl_ctx* ctx, size_t bytes, void** out);

#define SL_REQUIRE(ctx, cond, msg)                                                   \
    do {                                                                             \
        if (!(cond)) return sl_set_error((ctx), SL_ERR_INVALID_ARG, "%s: %s", __func__, (msg)); \
    } while (0)

#define SL_CUDA(ctx, expr)                                                                           \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return sl_set_error((ctx), SL_ERR_CUDA, "%s: %s -> %s", __func__, #expr, cudaGetErrorString(e__)); \
    } while (0)

// Launch + count + check.  Every kernel this library runs goes through here (sl_ctx_launch_count).
#define SL_LAUNCH(ctx, kernel, grid, block, smem, ...)                                               \
    do {                                                                                             \
        sl_ctx::ProfRec pr__{};                                                                      \
        const bool prof__ = (ctx)->profiling && (ctx)->profiling_all;                                \
        if (prof__) {                                                                                \
            cudaEventCreate(&pr__.a);                                                                \
            cudaEventCreate(&pr__.b);                                                                \
            pr__.name = #kernel;                                                                     \
            cudaEventRecord(pr__.a, (ctx)->stream);                                                  \
        }                                                                                            \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                             \
        if (prof__) {                                                                                \
            cudaEventRecord(pr__.b, (ctx)->stream);                                                  \
            (ctx)->prof.push_back(pr__);                                                             \
        }                                                                                            \
        (ctx)->launches++;                                                                           \
        cudaError_t e__ = cudaGetLastError();                                                        \
        if (e__ != cudaSuccess)                                                                      \
            return sl_set_error((ctx), SL_ERR_CUDA, "%s: launch %s -> %s", __func__, #kernel, cudaGetErrorString(e__)); \
    } while (0)

#define SL_DISPATCH_DTYPE(ctx, dtype, T, ...)                                         \
    switch (dtype) {                                                                  \
    case SL_F32: { using T = float; __VA_ARGS__; } break;                             \
    case SL_F64: { using T = double; __VA_ARGS__; } break;                            \
    case SL_I32: { using T = int32_t; __VA_ARGS__; } break;                           \
    default: return sl_set_error((ctx), SL_ERR_INVALID_ARG, "%s: bad dtype %d", __func__, (int)(dtype)); \
    }

#define SL_DISPATCH_FLOAT(ctx, dtype, T, ...)                                         \
    switch (dtype) {                                                                  \
    case SL_F32: { using T = float; __VA_ARGS__; } break;                             \
    case SL_F64: { using T = double; __VA_ARGS__; } break;                            \
    default: return sl_set_error((ctx), SL_ERR_UNSUPPORTED, "%s: dtype %d unsupported (f32/f64 only)", __func__, (int)(dtype)); \
    }

static inline size_t sl_dtype_size(int dtype) { return dtype == SL_F64 ? 8 : 4; }

// ---------------------------------------------------------------------------------------------
// 128-bit packs.  VEC = 16 / sizeof(T) elements travel in one LDG.128 / STG.128.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct alignas(16) Pack {
    static constexpr int N = 16 / sizeof(T);
    T v[N];
};

// streaming (read-once) 128-bit load: bypass L1 allocation, read-only path
template <typename T>
__device__ __forceinline__ Pack<T> ld_stream(const T* p) {
    Pack<T> r;
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    uint4 u = make_uint4(a, b, c, d);
    r = *reinterpret_cast<Pack<T>*>(&u);
    return r;
}
// plain 128-bit load (for buffers that are read-modify-written in the same kernel)
template <typename T>
__device__ __forceinline__ Pack<T> ld_pack(const T* p) {
    return *reinterpret_cast<const Pack<T>*>(p);
}
template <typename T>
__device__ __forceinline__ void st_pack(T* p, const Pack<T>& v) {
    *reinterpret_cast<Pack<T>*>(p) = v;
}
// streaming store: written once, not re-read by this kernel
template <typename T>
__device__ __forceinline__ void st_stream(T* p, const Pack<T>& v) {
    const uint4 u = *reinterpret_cast<const uint4*>(&v);
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

static inline bool sl_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
__device__ __forceinline__ T sl_max2(T a, T b) { return a > b ? a : b; }
template <>
__device__ __forceinline__ float sl_max2<float>(float a, float b) { return fmaxf(a, b); }
template <>
__device__ __forceinline__ double sl_max2<double>(double a, double b) { return fmax(a, b); }

static inline unsigned sl_pow2_ceil(unsigned x) {
    unsigned p = 1;
    while (p < x) p <<= 1;
    return p;
}
