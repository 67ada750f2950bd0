// softmax.cu — family S: row softmax forward (single pass: the row is read from HBM once, kept in registers,
// written once = 8 B/elem) and its gradient in closed form  dx = s * (g - <s,g>)  (12 B/elem, SET).
//
// The reference computes the forward as 5 separate passes (max_cols, sub_cols, exp, sum_cols, div_cols —
// src/ops2/softmax/cpu.rs:11-16) and the backward by materialising the F x F Jacobian per row
// (src/ops2/softmax/grad/cpu.rs:40-60); the values agree up to fp rounding (tests pin the tolerance).
//
// Row ownership by feature count:  F <= 32: one thread per row;  F <= 1024: one warp per row;
// otherwise one 256..1024-thread block per row with the row cached in registers (up to 8 packs per thread),
// falling back to a re-reading loop for rows that do not fit or are not 16-byte aligned.
#include "common.cuh"

namespace {

__device__ __forceinline__ float m_exp(float a) { return expf(a); }
__device__ __forceinline__ double m_exp(double a) { return exp(a); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(0xffffffffu, v, o);
        v = ov > v ? ov : v;
    }
    return v;
}

// block-wide reductions over NW warps, fixed order, result broadcast to every thread
template <typename T, bool IS_MAX>
__device__ __forceinline__ T block_reduce(T v, T* s_buf) {
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    const int w = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[w] = v;
    __syncthreads();
    T r = s_buf[0];
    for (int k = 1; k < nw; ++k) r = IS_MAX ? (s_buf[k] > r ? s_buf[k] : r) : (r + s_buf[k]);
    return r;
}

// ---------------------------------------------------------------- forward
// thread per row (tiny F, e.g. the 10-class head of the nn.rs MLP): the row stays in L1 between the passes
template <typename T>
__global__ void __launch_bounds__(256) softmax_thread_kernel(size_t samples, size_t features, const T* __restrict__ x, T* __restrict__ out) {
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < samples; r += (size_t)gridDim.x * blockDim.x) {
        const T* row = x + r * features;
        T* orow = out + r * features;
        T mx = row[0];
        for (size_t c = 1; c < features; ++c) mx = row[c] > mx ? row[c] : mx;
        T sum = T(0);
        for (size_t c = 0; c < features; ++c) {
            const T e = m_exp(row[c] - mx);
            orow[c] = e;
            sum += e;
        }
        for (size_t c = 0; c < features; ++c) orow[c] = orow[c] / sum;
    }
}

// warp per row, F <= 1024: up to 32 values per lane in registers
template <typename T>
__global__ void __launch_bounds__(256) softmax_warp_kernel(size_t samples, size_t features, const T* __restrict__ x, T* __restrict__ out) {
    constexpr int MAXV = 32;
    const int lane = threadIdx.x & 31;
    const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = wid; r < samples; r += nwarps) {
        const T* row = x + r * features;
        T v[MAXV];
        T mx = __ldg(row);
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const size_t c = (size_t)k * 32 + lane;
            if (c < features) {
                v[k] = __ldg(row + c);
                mx = v[k] > mx ? v[k] : mx;
            }
        }
        mx = warp_max(mx);
        T sum = T(0);
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const size_t c = (size_t)k * 32 + lane;
            if (c < features) {
                v[k] = m_exp(v[k] - mx);
                sum += v[k];
            }
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const size_t c = (size_t)k * 32 + lane;
            if (c < features) out[r * features + c] = v[k] / sum;
        }
    }
}

// block per row, row cached in registers as NP 128-bit packs per thread (features <= NP * VEC * blockDim.x)
template <typename T, int NP>
__global__ void __launch_bounds__(1024) softmax_block_kernel(size_t samples, size_t features, const T* __restrict__ x, T* __restrict__ out) {
    constexpr int V = Pack<T>::N;
    __shared__ T s_buf[32];
    const size_t packs = features / V;
    for (size_t r = blockIdx.x; r < samples; r += gridDim.x) {
        const T* row = x + r * features;
        Pack<T> p[NP];
        T mx = __ldg(row);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const size_t i = (size_t)k * blockDim.x + threadIdx.x;
            if (i < packs) {
                p[k] = ld_stream(row + i * V);
#pragma unroll
                for (int e = 0; e < V; ++e) mx = p[k].v[e] > mx ? p[k].v[e] : mx;
            }
        }
        mx = block_reduce<T, true>(mx, s_buf);
        T sum = T(0);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const size_t i = (size_t)k * blockDim.x + threadIdx.x;
            if (i < packs) {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    p[k].v[e] = m_exp(p[k].v[e] - mx);
                    sum += p[k].v[e];
                }
            }
        }
        sum = block_reduce<T, false>(sum, s_buf);
        const T inv = T(1) / sum;  // one IEEE division per row; e * (1/s) is within 1 ulp of e / s (tolerance: K-scaled)
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const size_t i = (size_t)k * blockDim.x + threadIdx.x;
            if (i < packs) {
#pragma unroll
                for (int e = 0; e < V; ++e) p[k].v[e] = p[k].v[e] * inv;
                st_stream(out + r * features + i * V, p[k]);
            }
        }
    }
}

// ---------------------------------------------------------------- forward, long rows: bulk-async (TMA engine) row pipeline
// One persistent 512-thread CTA per SM; rows travel HBM -> shared memory -> HBM with cp.async.bulk (1-D TMA) through a ring of
// 3 row buffers, so the load of row i+2 and the store of row i-1 are in flight while row i is being reduced in registers.
// The register-only variant above stops issuing memory traffic during its two block-wide reductions and its expf phase
// (measured 54-69 % of HBM peak at 16384 features); here HBM never idles.
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load_row(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_row(void* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bar_wait_parity(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

template <typename T, int NP>
__global__ void __launch_bounds__(512, 1) softmax_pipe_kernel(size_t samples, size_t features, const T* __restrict__ x, T* __restrict__ out) {
    constexpr int V = Pack<T>::N;
    constexpr int NBUF = 3;
    extern __shared__ __align__(128) unsigned char sm_raw[];
    __shared__ T s_buf[32];
    __shared__ __align__(8) unsigned long long bars[NBUF];
    const uint32_t row_bytes = (uint32_t)(features * sizeof(T));
    const size_t packs = features / V;
    const size_t nrows = samples > blockIdx.x ? (samples - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;  // rows of this CTA
    auto row_of = [&](size_t i) { return (size_t)blockIdx.x + i * gridDim.x; };
    auto buf_ptr = [&](int b) { return reinterpret_cast<T*>(sm_raw + (size_t)b * row_bytes); };
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_u32(&bars[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (size_t i = 0; i < 2 && i < nrows; ++i)
            bulk_load_row(sm_u32(buf_ptr((int)i)), x + row_of(i) * features, row_bytes, sm_u32(&bars[i]));
    }
    __syncthreads();
    for (size_t i = 0; i < nrows; ++i) {
        const int b = (int)(i % NBUF);
        bar_wait_parity(sm_u32(&bars[b]), (uint32_t)((i / NBUF) & 1));
        T* row = buf_ptr(b);
        Pack<T> p[NP];
        T mx = row[0];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const size_t j = (size_t)k * blockDim.x + threadIdx.x;
            if (j < packs) {
                p[k] = *reinterpret_cast<const Pack<T>*>(row + j * V);
#pragma unroll
                for (int e = 0; e < V; ++e) mx = p[k].v[e] > mx ? p[k].v[e] : mx;
            }
        }
        mx = block_reduce<T, true>(mx, s_buf);
        T sum = T(0);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const size_t j = (size_t)k * blockDim.x + threadIdx.x;
            if (j < packs) {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    p[k].v[e] = m_exp(p[k].v[e] - mx);
                    sum += p[k].v[e];
                }
            }
        }
        sum = block_reduce<T, false>(sum, s_buf);
        const T inv = T(1) / sum;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const size_t j = (size_t)k * blockDim.x + threadIdx.x;
            if (j < packs) {
#pragma unroll
                for (int e = 0; e < V; ++e) p[k].v[e] = p[k].v[e] * inv;
                *reinterpret_cast<Pack<T>*>(row + j * V) = p[k];
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk store
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store_row(out + row_of(i) * features, sm_u32(row), row_bytes);
            if (i + 2 < nrows) {
                // buffer (i+2)%3 was last stored from at iteration i-1: at most the store just committed may still be reading
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                const int nb = (int)((i + 2) % NBUF);
                bulk_load_row(sm_u32(buf_ptr(nb)), x + row_of(i + 2) * features, row_bytes, sm_u32(&bars[nb]));
            }
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA's smem is released
    __syncthreads();
}

// generic fallback: block per row, re-reads the row (any length / alignment)
template <typename T>
__global__ void __launch_bounds__(256) softmax_loop_kernel(size_t samples, size_t features, const T* __restrict__ x, T* __restrict__ out) {
    __shared__ T s_buf[32];
    for (size_t r = blockIdx.x; r < samples; r += gridDim.x) {
        const T* row = x + r * features;
        T* orow = out + r * features;
        T mx = __ldg(row);
        for (size_t c = threadIdx.x; c < features; c += blockDim.x) {
            const T v = __ldg(row + c);
            mx = v > mx ? v : mx;
        }
        mx = block_reduce<T, true>(mx, s_buf);
        T sum = T(0);
        for (size_t c = threadIdx.x; c < features; c += blockDim.x) {
            const T e = m_exp(__ldg(row + c) - mx);
            orow[c] = e;
            sum += e;
        }
        sum = block_reduce<T, false>(sum, s_buf);
        for (size_t c = threadIdx.x; c < features; c += blockDim.x) orow[c] = orow[c] / sum;
    }
}

// ---------------------------------------------------------------- backward (closed form, SET)
template <typename T>
__global__ void __launch_bounds__(256) softmax_grad_thread_kernel(size_t samples, size_t features, T* __restrict__ xg, const T* __restrict__ s,
                                                                  const T* __restrict__ g) {
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < samples; r += (size_t)gridDim.x * blockDim.x) {
        const T* sr = s + r * features;
        const T* gr = g + r * features;
        T dot = T(0);
        for (size_t c = 0; c < features; ++c) dot += sr[c] * gr[c];
        for (size_t c = 0; c < features; ++c) xg[r * features + c] = sr[c] * (gr[c] - dot);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) softmax_grad_warp_kernel(size_t samples, size_t features, T* __restrict__ xg, const T* __restrict__ s,
                                                                const T* __restrict__ g) {
    constexpr int MAXV = 32;
    const int lane = threadIdx.x & 31;
    const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = wid; r < samples; r += nwarps) {
        T sv[MAXV], gv[MAXV];
        T dot = T(0);
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const size_t c = (size_t)k * 32 + lane;
            if (c < features) {
                sv[k] = __ldg(s + r * features + c);
                gv[k] = __ldg(g + r * features + c);
                dot += sv[k] * gv[k];
            }
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const size_t c = (size_t)k * 32 + lane;
            if (c < features) xg[r * features + c] = sv[k] * (gv[k] - dot);
        }
    }
}

template <typename T, int NP>
__global__ void __launch_bounds__(1024) softmax_grad_block_kernel(size_t samples, size_t features, T* __restrict__ xg, const T* __restrict__ s,
                                                                  const T* __restrict__ g) {
    constexpr int V = Pack<T>::N;
    __shared__ T s_buf[32];
    const size_t packs = features / V;
    for (size_t r = blockIdx.x; r < samples; r += gridDim.x) {
        Pack<T> ps[NP], pg[NP];
        T dot = T(0);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const size_t i = (size_t)k * blockDim.x + threadIdx.x;
            if (i < packs) {
                ps[k] = ld_stream(s + r * features + i * V);
                pg[k] = ld_stream(g + r * features + i * V);
#pragma unroll
                for (int e = 0; e < V; ++e) dot += ps[k].v[e] * pg[k].v[e];
            }
        }
        dot = block_reduce<T, false>(dot, s_buf);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const size_t i = (size_t)k * blockDim.x + threadIdx.x;
            if (i < packs) {
#pragma unroll
                for (int e = 0; e < V; ++e) ps[k].v[e] = ps[k].v[e] * (pg[k].v[e] - dot);
                st_stream(xg + r * features + i * V, ps[k]);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) softmax_grad_loop_kernel(size_t samples, size_t features, T* __restrict__ xg, const T* __restrict__ s,
                                                                const T* __restrict__ g) {
    __shared__ T s_buf[32];
    for (size_t r = blockIdx.x; r < samples; r += gridDim.x) {
        T dot = T(0);
        for (size_t c = threadIdx.x; c < features; c += blockDim.x) dot += __ldg(s + r * features + c) * __ldg(g + r * features + c);
        dot = block_reduce<T, false>(dot, s_buf);
        for (size_t c = threadIdx.x; c < features; c += blockDim.x)
            xg[r * features + c] = __ldg(s + r * features + c) * (__ldg(g + r * features + c) - dot);
    }
}

// pick (threads, NP) so that NP * VEC * threads >= features with threads in {256, 512, 1024}, NP in {1,2,4,8}
struct BlockPlan {
    int threads;
    int np;
};
template <typename T>
static BlockPlan plan_block(size_t features, int max_np) {
    const size_t packs = features / Pack<T>::N;
    // smallest block first: several resident blocks per SM sit in different phases (load / reduce / exp / store), which keeps HBM
    // busy while one block is inside its reductions (one 1024-thread block per SM measured 54 % of peak at 16384 features)
    for (int th = 256; th <= 1024; th *= 2)
        for (int np = 1; np <= max_np; np *= 2)
            if ((size_t)np * th >= packs) return {th, np};
    return {0, 0};
}

template <typename T>
int softmax_t(sl_ctx* ctx, size_t samples, size_t features, const void* x, void* out) {
    const size_t cap = (size_t)ctx->num_sms * 8;
    if (features <= 32) {
        size_t blocks = (samples + 255) / 256;
        SL_LAUNCH(ctx, (softmax_thread_kernel<T>), (unsigned)(blocks < cap ? blocks : cap), 256, 0, samples, features, (const T*)x, (T*)out);
        return SL_OK;
    }
    if (features <= 1024) {
        size_t blocks = (samples + 7) / 8;
        SL_LAUNCH(ctx, (softmax_warp_kernel<T>), (unsigned)(blocks < cap ? blocks : cap), 256, 0, samples, features, (const T*)x, (T*)out);
        return SL_OK;
    }
    const bool vec = (features % Pack<T>::N == 0) && sl_aligned16(x) && sl_aligned16(out);
    const size_t row_bytes = features * sizeof(T);
    if (vec && row_bytes > 8192 && row_bytes <= 65536 && samples >= (size_t)ctx->num_sms && !getenv("SLICED_SOFTMAX_NO_PIPE")) {
        const size_t packs = features / Pack<T>::N;
        const int np = packs <= 1024 ? 2 : (packs <= 2048 ? 4 : 8);
        const size_t smem = 3 * row_bytes;
        const unsigned pgrid = (unsigned)(samples < (size_t)ctx->num_sms ? samples : (size_t)ctx->num_sms);
#define SL_PIPE(NPV)                                                                                                   \
    {                                                                                                                  \
        auto kern = softmax_pipe_kernel<T, NPV>;                                                                       \
        SL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        SL_LAUNCH(ctx, kern, pgrid, 512, smem, samples, features, (const T*)x, (T*)out);                               \
    }
        if (np == 2) SL_PIPE(2) else if (np == 4) SL_PIPE(4) else SL_PIPE(8)
#undef SL_PIPE
        return SL_OK;
    }
    BlockPlan bp = vec ? plan_block<T>(features, 8) : BlockPlan{0, 0};
    const unsigned grid = (unsigned)(samples < cap * 4 ? samples : cap * 4);
    if (bp.np == 1) SL_LAUNCH(ctx, (softmax_block_kernel<T, 1>), grid, bp.threads, 0, samples, features, (const T*)x, (T*)out);
    else if (bp.np == 2) SL_LAUNCH(ctx, (softmax_block_kernel<T, 2>), grid, bp.threads, 0, samples, features, (const T*)x, (T*)out);
    else if (bp.np == 4) SL_LAUNCH(ctx, (softmax_block_kernel<T, 4>), grid, bp.threads, 0, samples, features, (const T*)x, (T*)out);
    else if (bp.np == 8) SL_LAUNCH(ctx, (softmax_block_kernel<T, 8>), grid, bp.threads, 0, samples, features, (const T*)x, (T*)out);
    else SL_LAUNCH(ctx, (softmax_loop_kernel<T>), grid, 256, 0, samples, features, (const T*)x, (T*)out);
    return SL_OK;
}

template <typename T>
int softmax_grad_t(sl_ctx* ctx, size_t samples, size_t features, void* xg, const void* s, const void* g) {
    const size_t cap = (size_t)ctx->num_sms * 8;
    if (features <= 32) {
        size_t blocks = (samples + 255) / 256;
        SL_LAUNCH(ctx, (softmax_grad_thread_kernel<T>), (unsigned)(blocks < cap ? blocks : cap), 256, 0, samples, features, (T*)xg, (const T*)s,
                  (const T*)g);
        return SL_OK;
    }
    if (features <= 1024) {
        size_t blocks = (samples + 7) / 8;
        SL_LAUNCH(ctx, (softmax_grad_warp_kernel<T>), (unsigned)(blocks < cap ? blocks : cap), 256, 0, samples, features, (T*)xg, (const T*)s,
                  (const T*)g);
        return SL_OK;
    }
    const bool vec = (features % Pack<T>::N == 0) && sl_aligned16(xg) && sl_aligned16(s) && sl_aligned16(g);
    BlockPlan bp = vec ? plan_block<T>(features, 4) : BlockPlan{0, 0};
    const unsigned grid = (unsigned)(samples < cap * 4 ? samples : cap * 4);
    if (bp.np == 1) SL_LAUNCH(ctx, (softmax_grad_block_kernel<T, 1>), grid, bp.threads, 0, samples, features, (T*)xg, (const T*)s, (const T*)g);
    else if (bp.np == 2) SL_LAUNCH(ctx, (softmax_grad_block_kernel<T, 2>), grid, bp.threads, 0, samples, features, (T*)xg, (const T*)s, (const T*)g);
    else if (bp.np == 4) SL_LAUNCH(ctx, (softmax_grad_block_kernel<T, 4>), grid, bp.threads, 0, samples, features, (T*)xg, (const T*)s, (const T*)g);
    else SL_LAUNCH(ctx, (softmax_grad_loop_kernel<T>), grid, 256, 0, samples, features, (T*)xg, (const T*)s, (const T*)g);
    return SL_OK;
}


// ---------------------------------------------------------------- fused softmax + categorical cross-entropy
// examples/nn.rs:190-233 after the last Linear, in ONE pass over the logits:
//   s      = softmax(z)                                   (softmax/cpu.rs:11-16)
//   loss_r = -ln( sum_c clip(s, 1e-7, 1 - 1e-7)[r,c] * y[r,c] )            cce,      nn.rs:124-138
//   g      = (-(y / s)) / rows                                            cce_grad, nn.rs:140-152 (unclipped division)
//   dz     = softmax_grad(s, g) = s * (g - <s, g>)  (SET)                  softmax/grad/cpu.rs:14-62, closed form
//   correct += (argmax_c s[r,c] == labels[r])                             nn.rs:195-211
// The thread-per-row form (features <= 32: the 10-class head) performs exactly the per-element operations of the separate
// kernels (softmax_thread_kernel, unary CLIP, binary MUL, rowreduce sum, unary NEG_LN, binary DIV, unary NEG_DIV_SCALAR,
// softmax_grad_thread_kernel, count_correct_kernel) in the same order, so every output is bit-identical to that 9-launch chain.
__device__ __forceinline__ float m_log(float a) { return logf(a); }
__device__ __forceinline__ double m_log(double a) { return log(a); }

template <typename T>
__device__ __forceinline__ T cce_clip(T x, T lo, T hi) {
    const T m = x > lo ? x : lo;
    return m < hi ? m : hi;
}

template <typename T, int MAXF>
__global__ void __launch_bounds__(256) softmax_cce_thread_kernel(size_t samples, size_t features, const T* __restrict__ z, const T* __restrict__ y,
                                                                 const int32_t* __restrict__ labels, T rows_div, T clip_lo, T clip_hi,
                                                                 T* __restrict__ probs, T* __restrict__ dz, T* __restrict__ loss_out,
                                                                 int32_t* correct) {
    int local = 0;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < samples; r += (size_t)gridDim.x * blockDim.x) {
        const T* zr = z + r * features;
        const T* yr = y + r * features;
        T sv[MAXF];
        T mx = zr[0];
#pragma unroll
        for (int c = 1; c < MAXF; ++c)
            if ((size_t)c < features) mx = zr[c] > mx ? zr[c] : mx;
        T sum = T(0);
#pragma unroll
        for (int c = 0; c < MAXF; ++c)
            if ((size_t)c < features) {
                sv[c] = m_exp(zr[c] - mx);
                sum += sv[c];
            }
        T loss_acc = T(0), dot = T(0);
        T smax = T(0);
        int smi = 0;
#pragma unroll
        for (int c = 0; c < MAXF; ++c)
            if ((size_t)c < features) {
                const T sc = sv[c] / sum;
                sv[c] = sc;
                probs[r * features + c] = sc;
                loss_acc += cce_clip(sc, clip_lo, clip_hi) * yr[c];
                const T g = (-(yr[c] / sc)) / rows_div;
                dot += sc * g;
                if (c == 0 || sc > smax) {
                    smax = sc;
                    smi = c;
                }
            }
#pragma unroll
        for (int c = 0; c < MAXF; ++c)
            if ((size_t)c < features) {
                const T g = (-(yr[c] / sv[c])) / rows_div;
                dz[r * features + c] = sv[c] * (g - dot);
            }
        loss_out[r] = -m_log(loss_acc);
        if (labels) local += (labels[r] == smi) ? 1 : 0;
    }
    if (correct) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(correct, local);  // integer atomics: order-independent
    }
}

// any feature count: one block per row, the row is re-read from L1/L2 between the passes (fixed-order block reductions)
template <typename T>
__global__ void __launch_bounds__(256) softmax_cce_block_kernel(size_t samples, size_t features, const T* __restrict__ z, const T* __restrict__ y,
                                                                const int32_t* __restrict__ labels, T rows_div, T clip_lo, T clip_hi,
                                                                T* __restrict__ probs, T* __restrict__ dz, T* __restrict__ loss_out, int32_t* correct) {
    __shared__ T s_buf[32];
    __shared__ int s_idx[8];
    for (size_t r = blockIdx.x; r < samples; r += gridDim.x) {
        const T* zr = z + r * features;
        const T* yr = y + r * features;
        T* pr = probs + r * features;
        T mx = zr[0];
        for (size_t c = threadIdx.x; c < features; c += blockDim.x) mx = zr[c] > mx ? zr[c] : mx;
        mx = block_reduce<T, true>(mx, s_buf);
        T sum = T(0);
        for (size_t c = threadIdx.x; c < features; c += blockDim.x) {
            const T e = m_exp(zr[c] - mx);
            pr[c] = e;
            sum += e;
        }
        sum = block_reduce<T, false>(sum, s_buf);
        T loss_acc = T(0), dot = T(0);
        T smax = T(-1);
        int smi = 0x7fffffff;
        for (size_t c = threadIdx.x; c < features; c += blockDim.x) {   // every thread re-reads only what it wrote itself
            const T sc = pr[c] / sum;
            pr[c] = sc;
            loss_acc += cce_clip(sc, clip_lo, clip_hi) * yr[c];
            dot += sc * ((-(yr[c] / sc)) / rows_div);
            if (sc > smax) {
                smax = sc;
                smi = (int)c;
            }
        }
        loss_acc = block_reduce<T, false>(loss_acc, s_buf);
        dot = block_reduce<T, false>(dot, s_buf);
        for (size_t c = threadIdx.x; c < features; c += blockDim.x) dz[r * features + c] = pr[c] * (((-(yr[c] / pr[c])) / rows_div) - dot);
        if (labels && correct) {   // first index attaining the maximum: (value, index) folded lane -> warp -> block in a fixed order
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const T ov = __shfl_xor_sync(0xffffffffu, smax, o);
                const int oi = __shfl_xor_sync(0xffffffffu, smi, o);
                if (ov > smax || (ov == smax && oi < smi)) {
                    smax = ov;
                    smi = oi;
                }
            }
            __syncthreads();
            if ((threadIdx.x & 31) == 0) {
                s_buf[threadIdx.x >> 5] = smax;
                s_idx[threadIdx.x >> 5] = smi;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
                    if (s_buf[k] > smax || (s_buf[k] == smax && s_idx[k] < smi)) {
                        smax = s_buf[k];
                        smi = s_idx[k];
                    }
                if (labels[r] == smi) atomicAdd(correct, 1);
            }
        }
        if (threadIdx.x == 0) loss_out[r] = -m_log(loss_acc);
        __syncthreads();
    }
}

template <typename T>
int softmax_cce_t(sl_ctx* ctx, size_t samples, size_t features, const void* z, const void* y, const int32_t* labels, size_t grad_rows, void* probs,
                  void* dz, void* loss, int32_t* correct) {
    const size_t cap = (size_t)ctx->num_sms * 8;
    const T rows_div = (T)(double)grad_rows, lo = (T)1E-7, hi = (T)(1. - 1E-7);   // the scalars as custos casts them (f64 literal -> T)
#define SL_CCE_THREAD(MAXF)                                                                                                                \
    SL_LAUNCH(ctx, (softmax_cce_thread_kernel<T, MAXF>), (unsigned)(blocks < cap ? blocks : cap), 256, 0, samples, features, (const T*)z,  \
              (const T*)y, labels, rows_div, lo, hi, (T*)probs, (T*)dz, (T*)loss, correct)
    if (features <= 32) {
        const size_t blocks = (samples + 255) / 256;
        if (features <= 10) SL_CCE_THREAD(10);
        else if (features <= 16) SL_CCE_THREAD(16);
        else SL_CCE_THREAD(32);
        return SL_OK;
    }
#undef SL_CCE_THREAD
    SL_LAUNCH(ctx, (softmax_cce_block_kernel<T>), (unsigned)(samples < cap ? samples : cap), 256, 0, samples, features, (const T*)z, (const T*)y, labels,
              rows_div, lo, hi, (T*)probs, (T*)dz, (T*)loss, correct);
    return SL_OK;
}

}  // namespace

extern "C" {

int sl_softmax(sl_ctx* ctx, int dtype, size_t samples, size_t features, const void* x, void* out) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    if (samples == 0 || features == 0) return SL_OK;
    SL_REQUIRE(ctx, x && out, "NULL pointer");
    SL_DISPATCH_FLOAT(ctx, dtype, T, return softmax_t<T>(ctx, samples, features, x, out));
    return SL_OK;
}

int sl_softmax_grad(sl_ctx* ctx, int dtype, size_t samples, size_t features, void* x_grad, const void* out, const void* out_grad) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, x_grad);
    if (samples == 0 || features == 0) return SL_OK;
    SL_REQUIRE(ctx, x_grad && out && out_grad, "NULL pointer");
    SL_DISPATCH_FLOAT(ctx, dtype, T, return softmax_grad_t<T>(ctx, samples, features, x_grad, out, out_grad));
    return SL_OK;
}

int sl_softmax_cce(sl_ctx* ctx, int dtype, size_t samples, size_t features, const void* logits, const void* targets, const int32_t* labels,
                   size_t grad_rows, void* probs_out, void* logits_grad, void* loss_per_sample, int32_t* correct_dev) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, probs_out, logits_grad, loss_per_sample, correct_dev);
    if (samples == 0 || features == 0) return SL_OK;
    SL_REQUIRE(ctx, logits && targets && probs_out && logits_grad && loss_per_sample, "NULL pointer");
    SL_REQUIRE(ctx, grad_rows > 0, "grad_rows == 0");
    SL_DISPATCH_FLOAT(ctx, dtype, T,
                      return softmax_cce_t<T>(ctx, samples, features, logits, targets, labels, grad_rows, probs_out, logits_grad, loss_per_sample,
                                              labels ? correct_dev : nullptr));
    return SL_OK;
}

}  // extern "C"
