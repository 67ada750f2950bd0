// gemm_skinny.cu — HBM-roofline kernels for gemms with one extent <= 16: the 10-class head of the nn.rs MLP
// (reference examples/nn.rs:199-201 Linear<4096,10>: gemm at src/ops2/gemm/cpu.rs, gemm_grad at src/ops2/gemm/grad.rs).
//
//   nn:  C[m x n]  = A[m x k]   * B[k x n],    n <= 16   z3  = a2 * W3            reads the big A once
//   tn:  C[m x n]  = A[k x m]^T * B[k x n],    n <= 16   dW3 = a2^T * dz3         reads the big A once
//   nt:  C[m x n]  = A[m x k]   * B[n x k]^T,  k <= 16   da2 = dz3 * W3^T (* relu mask)   writes the big C once
//
// Each moves ~1 GB at the MLP shapes (0.16 ms at the measured 6.5 TB/s) but also needs 10 FMAs per 4 bytes moved, i.e. 0.09 ms
// of pure FMA issue on 148 SMs: the kernels are therefore written around INSTRUCTION count, not just bytes — the small extent
// is a template parameter (no padding to 16: 37 % fewer FMAs), the small operand lives in shared memory in the exact order
// the FMAs consume it (128-bit broadcast or conflict-free loads), and the big operand is streamed with 128-bit loads kept 8
// rows deep in flight per thread.  v1 of these kernels (gemm_simt.cu) ran at 25 % of the HBM roofline; it remains the fallback
// for shapes these do not take.
//
// fp32 FMA accumulation; the summation order differs from the sequential-k oracle (K-scaled tolerance in the tests);
// SL_GEMM_SIMT mode never comes here.
#include "common.cuh"

int sl_gemm_skinny_f32_v1(sl_ctx* ctx, int trans_a, int trans_b, size_t m, size_t n, size_t k, const float* a, const float* b, float* c,
                          int accumulate);

namespace {

// ------------------------------------------------------------------------------------------------ nn
// One CTA per SM, all of B resident in shared memory as [k/4][NP + 1][4] floats: group g holds, for each column j, the float4
// (B[4g][j], B[4g+1][j], B[4g+2][j], B[4g+3][j]); the +1 pad makes the 8 lanes of a quarter-warp hit distinct banks.
// A warp owns RW = 8 rows; lane l of step s multiplies its float4 A[r][128 s + 4 l ..] with group 32 s + l.
template <int NP, bool ACC>
__global__ void __launch_bounds__(384, 1) skinny_nn_kernel(size_t m, size_t n, size_t k, const float* __restrict__ A, const float* __restrict__ B,
                                                           float* C) {
    constexpr int RW = 8;
    extern __shared__ __align__(16) float Bs[];
    const size_t groups = (k + 3) / 4;
    for (size_t e = threadIdx.x; e < groups * NP; e += blockDim.x) {
        const size_t g = e / NP;
        const int j = (int)(e % NP);
        float4 v;
        v.x = ((size_t)j < n && 4 * g + 0 < k) ? __ldg(B + (4 * g + 0) * n + j) : 0.f;
        v.y = ((size_t)j < n && 4 * g + 1 < k) ? __ldg(B + (4 * g + 1) * n + j) : 0.f;
        v.z = ((size_t)j < n && 4 * g + 2 < k) ? __ldg(B + (4 * g + 2) * n + j) : 0.f;
        v.w = ((size_t)j < n && 4 * g + 3 < k) ? __ldg(B + (4 * g + 3) * n + j) : 0.f;
        *reinterpret_cast<float4*>(Bs + (g * (NP + 1) + j) * 4) = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
    const size_t row_groups = (m + RW - 1) / RW;
    const size_t steps = (groups + 31) / 32;
    for (size_t rg = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rg < row_groups; rg += warps) {
        const size_t r0 = rg * RW;
        float acc[RW][NP];
#pragma unroll
        for (int i = 0; i < RW; ++i)
#pragma unroll
            for (int j = 0; j < NP; ++j) acc[i][j] = 0.f;
        for (size_t s = 0; s < steps; ++s) {
            const size_t g = s * 32 + lane;
            if (g < groups) {   // k % 4 == 0 (host check): a group is either whole or absent
                float4 a[RW];
#pragma unroll
                for (int i = 0; i < RW; ++i) {
                    const size_t r = r0 + i < m ? r0 + i : m - 1;   // clamp: tail rows recompute the last row, never stored
                    const Pack<float> p = ld_stream(A + r * k + 4 * g);
                    a[i] = make_float4(p.v[0], p.v[1], p.v[2], p.v[3]);
                }
                const float4* bg = reinterpret_cast<const float4*>(Bs + g * (NP + 1) * 4);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float4 b = bg[j];
#pragma unroll
                    for (int i = 0; i < RW; ++i)
                        acc[i][j] = fmaf(a[i].w, b.w, fmaf(a[i].z, b.z, fmaf(a[i].y, b.y, fmaf(a[i].x, b.x, acc[i][j]))));
                }
            }
        }
        // butterfly over the lanes' k-subsets; every lane ends with all sums, lane i*NP+j (mod 32 rounds) stores one
#pragma unroll
        for (int i = 0; i < RW; ++i)
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                float v = acc[i][j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == ((i * NP + j) & 31) && r0 + i < m && (size_t)j < n) {
                    float* dst = C + (r0 + i) * n + j;
                    *dst = ACC ? *dst + v : v;
                }
            }
    }
}

// ------------------------------------------------------------------------------------------------ tn
// Thread owns 4 consecutive columns of A (= 4 rows of C) and walks a slice of k; the slice of B ([rows x NP], padded to
// NP + 2 floats per row so rows stay 8-byte aligned) sits in shared memory and is read as warp-uniform broadcasts.
// A rows are prefetched 8 deep (double-buffered registers).  partial[split][m][n]; folded in split order afterwards.
template <int NP>
__global__ void __launch_bounds__(128) skinny_tn_kernel(size_t m, size_t n, size_t k, size_t k_per_split, const float* __restrict__ A,
                                                        const float* __restrict__ B, float* __restrict__ partial) {
    constexpr int D = 8;
    constexpr int LDB = NP + 2;
    extern __shared__ __align__(16) float Bs[];
    const size_t kbeg = (size_t)blockIdx.y * k_per_split;
    const size_t kend = kbeg + k_per_split < k ? kbeg + k_per_split : k;
    const size_t rows = kend - kbeg;
    for (size_t e = threadIdx.x; e < rows * NP; e += blockDim.x) {
        const size_t r = e / NP;
        const int j = (int)(e % NP);
        Bs[r * LDB + j] = (size_t)j < n ? __ldg(B + (kbeg + r) * n + j) : 0.f;
    }
    __syncthreads();
    const size_t c0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c0 >= m) return;   // m % 4 == 0 (host check)
    float acc[4][NP];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) acc[i][j] = 0.f;
    const float* ap = A + kbeg * m + c0;
    float4 cur[D], nxt[D];
    auto load = [&](float4 (&dst)[D], size_t r) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const size_t rr = r + d < rows ? r + d : rows - 1;
            const Pack<float> p = ld_stream(ap + rr * m);
            dst[d] = make_float4(p.v[0], p.v[1], p.v[2], p.v[3]);
        }
    };
    load(cur, 0);
    for (size_t r = 0; r < rows; r += D) {
        if (r + D < rows) load(nxt, r + D);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (r + d < rows) {
                const float2* bp = reinterpret_cast<const float2*>(Bs + (r + d) * LDB);
#pragma unroll
                for (int j = 0; j < NP; j += 2) {
                    const float2 b = bp[j / 2];
                    acc[0][j] = fmaf(cur[d].x, b.x, acc[0][j]); acc[1][j] = fmaf(cur[d].y, b.x, acc[1][j]);
                    acc[2][j] = fmaf(cur[d].z, b.x, acc[2][j]); acc[3][j] = fmaf(cur[d].w, b.x, acc[3][j]);
                    acc[0][j + 1] = fmaf(cur[d].x, b.y, acc[0][j + 1]); acc[1][j + 1] = fmaf(cur[d].y, b.y, acc[1][j + 1]);
                    acc[2][j + 1] = fmaf(cur[d].z, b.y, acc[2][j + 1]); acc[3][j + 1] = fmaf(cur[d].w, b.y, acc[3][j + 1]);
                }
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) cur[d] = nxt[d];
    }
    float* p = partial + (size_t)blockIdx.y * m * n;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j)
            if ((size_t)j < n) p[(c0 + i) * n + j] = acc[i][j];
}

template <bool ACC>
__global__ void __launch_bounds__(256) skinny_fold_kernel(size_t total, size_t nsplit, const float* __restrict__ partial, float* C) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float s = partial[i];
        for (size_t p = 1; p < nsplit; ++p) s += partial[p * total + i];
        C[i] = ACC ? C[i] + s : s;
    }
}

// ------------------------------------------------------------------------------------------------ tn, small output, deep K
// C[m x n] (+)= A[k x m]^T * B[k x n] with few 64 x 64 output tiles and a long contraction: the weight gradients of small layers
// (examples/sine_net.rs: 64 x 64, 1 x 64 and 64 x 1 outputs over K = 1000 samples).  The generic CUDA-core kernel gives such a
// problem ONE block walking all of K (71 us per launch measured: 53 % of a sine_net step); here K is split over
// gridDim.z blocks (deterministic partials + fold), both operands are read coalesced (k is the slow index of both), and each
// thread keeps a 4 x 4 block of outputs in registers.
template <bool DIRECT_ACC>
__global__ void __launch_bounds__(256) smallout_tn_kernel(size_t m, size_t n, size_t k, size_t kps, const float* __restrict__ A,
                                                          const float* __restrict__ B, float* out, int direct) {
    constexpr int KT = 32;
    __shared__ float As[KT][64 + 4], Bs[KT][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t m0 = (size_t)blockIdx.y * 64, n0 = (size_t)blockIdx.x * 64;
    const size_t k0 = (size_t)blockIdx.z * kps, k1 = k0 + kps < k ? k0 + kps : k;
    float acc[4][4] = {};
    for (size_t kb = k0; kb < k1; kb += KT) {
#pragma unroll
        for (int i = 0; i < (KT * 64) / 256; ++i) {
            const int e = threadIdx.x + i * 256;
            const int kk = e >> 6, c = e & 63;
            const size_t kr = kb + kk;
            As[kk][c] = (kr < k1 && m0 + c < m) ? __ldg(A + kr * m + m0 + c) : 0.f;
            Bs[kk][c] = (kr < k1 && n0 + c < n) ? __ldg(B + kr * n + n0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < KT; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* dst = direct ? out : out + (size_t)blockIdx.z * m * n;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const size_t r = m0 + ty * 4 + i;
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const size_t c = n0 + tx * 4 + j;
            if (c < n) dst[r * n + c] = (DIRECT_ACC ? dst[r * n + c] : 0.f) + acc[i][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------ nt
// Thread owns 4 consecutive output columns: their B rows (4 x KP) live in registers for the whole kernel.  A CTA covers
// 1024 columns and a slab of rows whose A values ([rows x KP], padded to KP + 2) are staged in shared memory and read as
// broadcasts.  Optional fused epilogue: out *= (mask_src >= 0) (the relu gradient that follows this gemm in the MLP).
// MASK 1: the mask source is the pre-activation matrix (4 bytes per element); MASK 2: one bit per element ([m x n/32] words, n % 32 == 0).
template <int KP, bool ACC, int MASK>
__global__ void __launch_bounds__(256) skinny_nt_kernel(size_t m, size_t n, size_t k, size_t rows_per_block, const float* __restrict__ A,
                                                        const float* __restrict__ B, float* C, const float* __restrict__ mask_src,
                                                        const uint32_t* __restrict__ mask_bits) {
    constexpr int LDA = KP + 2;
    extern __shared__ __align__(16) float As[];
    const size_t r0 = (size_t)blockIdx.y * rows_per_block;
    const size_t r1 = r0 + rows_per_block < m ? r0 + rows_per_block : m;
    const size_t rows = r1 - r0;
    for (size_t e = threadIdx.x; e < rows * KP; e += blockDim.x) {
        const size_t r = e / KP;
        const int q = (int)(e % KP);
        As[r * LDA + q] = (size_t)q < k ? __ldg(A + (r0 + r) * k + q) : 0.f;
    }
    __syncthreads();
    const size_t j0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= n) return;   // n % 4 == 0 (host check)
    float b[4][KP];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int q = 0; q < KP; ++q) b[e][q] = (size_t)q < k ? __ldg(B + (j0 + e) * k + q) : 0.f;
    float* cp = C + r0 * n + j0;
    const float* mp = MASK == 1 ? mask_src + r0 * n + j0 : nullptr;
    const uint32_t* bp = MASK == 2 ? mask_bits + r0 * (n >> 5) + (j0 >> 5) : nullptr;
    const int bshift = (int)(j0 & 31);
#pragma unroll 4
    for (size_t r = 0; r < rows; ++r) {
        const float2* ap = reinterpret_cast<const float2*>(As + r * LDA);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < KP; q += 2) {
            const float2 a = ap[q / 2];
            o.x = fmaf(a.x, b[0][q], o.x); o.y = fmaf(a.x, b[1][q], o.y); o.z = fmaf(a.x, b[2][q], o.z); o.w = fmaf(a.x, b[3][q], o.w);
            o.x = fmaf(a.y, b[0][q + 1], o.x); o.y = fmaf(a.y, b[1][q + 1], o.y); o.z = fmaf(a.y, b[2][q + 1], o.z); o.w = fmaf(a.y, b[3][q + 1], o.w);
        }
        if (ACC) {
            const float4 c = *reinterpret_cast<const float4*>(cp + r * n);
            o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
        }
        if (MASK == 2) {
            const uint32_t w = __ldg(bp + r * (n >> 5)) >> bshift;
            o.x = (w & 1u ? 1.f : 0.f) * o.x; o.y = (w & 2u ? 1.f : 0.f) * o.y; o.z = (w & 4u ? 1.f : 0.f) * o.z; o.w = (w & 8u ? 1.f : 0.f) * o.w;
        }
        if (MASK == 1) {
            const Pack<float> mk = ld_stream(mp + r * n);
            o.x = (mk.v[0] >= 0.f ? 1.f : 0.f) * o.x; o.y = (mk.v[1] >= 0.f ? 1.f : 0.f) * o.y;
            o.z = (mk.v[2] >= 0.f ? 1.f : 0.f) * o.z; o.w = (mk.v[3] >= 0.f ? 1.f : 0.f) * o.w;
        }
        *reinterpret_cast<float4*>(cp + r * n) = o;
    }
}

// even extents 2..16 are instantiated; an odd extent runs the next even one with a zero column
#define SK_DISPATCH(NPV, ...)                    \
    switch (NPV) {                               \
    case 2: { constexpr int NP = 2; __VA_ARGS__; } break;   \
    case 4: { constexpr int NP = 4; __VA_ARGS__; } break;   \
    case 6: { constexpr int NP = 6; __VA_ARGS__; } break;   \
    case 8: { constexpr int NP = 8; __VA_ARGS__; } break;   \
    case 10: { constexpr int NP = 10; __VA_ARGS__; } break; \
    case 12: { constexpr int NP = 12; __VA_ARGS__; } break; \
    case 14: { constexpr int NP = 14; __VA_ARGS__; } break; \
    default: { constexpr int NP = 16; __VA_ARGS__; } break; \
    }

}  // namespace

// returns SL_OK when a skinny kernel handled the call, 1 when the shape is not skinny (caller falls through).
// mask_src (nt shape only): fused `C *= (mask_src >= 0)`; *mask_done tells the caller whether it was applied.
int sl_gemm_skinny_f32(sl_ctx* ctx, int trans_a, int trans_b, size_t m, size_t n, size_t k, const float* a, const float* b, float* c,
                       int accumulate, const float* mask_src, int* mask_done, const uint32_t* mask_bits) {
    if (mask_done) *mask_done = 0;
    if (mask_bits && (n % 32 != 0 || mask_src)) mask_bits = nullptr;   // (the caller applies the bits itself when *mask_done stays 0)
    const bool al = sl_aligned16(a) && sl_aligned16(b) && sl_aligned16(c);
    if (!trans_a && !trans_b && n >= 1 && n <= 16 && k >= 256 && m >= 64 && k % 4 == 0 && al) {
        const int np = (int)((n + 1) & ~size_t(1));
        const size_t smem = ((k + 3) / 4) * (size_t)(np + 1) * 16;
        if (smem <= 200 * 1024) {
            const size_t row_groups = (m + 7) / 8;
            const size_t want = (row_groups + 11) / 12;
            const unsigned grid = (unsigned)(want < (size_t)ctx->num_sms ? want : (size_t)ctx->num_sms);
            SK_DISPATCH(np, {
                if (accumulate) {
                    SL_CUDA(ctx, cudaFuncSetAttribute(skinny_nn_kernel<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    SL_LAUNCH(ctx, (skinny_nn_kernel<NP, true>), grid, 384, smem, m, n, k, a, b, c);
                } else {
                    SL_CUDA(ctx, cudaFuncSetAttribute(skinny_nn_kernel<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    SL_LAUNCH(ctx, (skinny_nn_kernel<NP, false>), grid, 384, smem, m, n, k, a, b, c);
                }
            });
            return SL_OK;
        }
    }
    if (trans_a && !trans_b && n >= 1 && n <= 16 && k >= 1024 && m >= 64 && m % 4 == 0 && al) {
        const int np = (int)((n + 1) & ~size_t(1));
        const unsigned gx = (unsigned)((m / 4 + 127) / 128);
        // about two waves of CTAs (3 resident per SM), slices of at least 256 and at most 896 rows (<= 64 KB of B in shared memory)
        size_t nsplit = ((size_t)ctx->num_sms * 6 + gx - 1) / gx;
        size_t kps = (k + nsplit - 1) / nsplit;
        kps = kps < 256 ? 256 : (kps > 896 ? 896 : kps);
        kps = (kps + 7) & ~size_t(7);
        nsplit = (k + kps - 1) / kps;
        void* ws = nullptr;
        int rc = sl_ws_reserve(ctx, nsplit * m * n * sizeof(float), &ws);
        if (rc != SL_OK) return rc;
        const size_t smem = kps * (size_t)(np + 2) * 4;
        SK_DISPATCH(np, {
            SL_CUDA(ctx, cudaFuncSetAttribute(skinny_tn_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
            SL_LAUNCH(ctx, (skinny_tn_kernel<NP>), dim3(gx, (unsigned)nsplit, 1), 128, smem, m, n, k, kps, a, b, (float*)ws);
        });
        const size_t total = m * n;
        const size_t cap = (size_t)ctx->num_sms * 8;
        const size_t fb = (total + 255) / 256;
        const unsigned fgrid = (unsigned)(fb < cap ? fb : cap);
        if (accumulate) SL_LAUNCH(ctx, (skinny_fold_kernel<true>), fgrid, 256, 0, total, nsplit, (const float*)ws, c);
        else SL_LAUNCH(ctx, (skinny_fold_kernel<false>), fgrid, 256, 0, total, nsplit, (const float*)ws, c);
        return SL_OK;
    }
    if (trans_a && !trans_b && k >= 256) {   // few output tiles, deep K: split K over the SMs (see smallout_tn_kernel)
        const size_t gx = (n + 63) / 64, gy = (m + 63) / 64, tiles = gx * gy;
        if (tiles <= (size_t)ctx->num_sms / 2 && gx <= 65535 && gy <= 65535) {
            size_t nsplit = ((size_t)ctx->num_sms * 2) / tiles;
            const size_t max_split = (k + 63) / 64;
            nsplit = nsplit > max_split ? max_split : (nsplit < 1 ? 1 : nsplit);
            size_t kps = (k + nsplit - 1) / nsplit;
            kps = (kps + 31) & ~size_t(31);
            nsplit = (k + kps - 1) / kps;
            const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)nsplit);
            if (nsplit == 1) {
                if (accumulate) SL_LAUNCH(ctx, (smallout_tn_kernel<true>), grid, 256, 0, m, n, k, kps, a, b, c, 1);
                else SL_LAUNCH(ctx, (smallout_tn_kernel<false>), grid, 256, 0, m, n, k, kps, a, b, c, 1);
                return SL_OK;
            }
            void* ws = nullptr;
            int rc = sl_ws_reserve(ctx, nsplit * m * n * sizeof(float), &ws);
            if (rc != SL_OK) return rc;
            SL_LAUNCH(ctx, (smallout_tn_kernel<false>), grid, 256, 0, m, n, k, kps, a, b, (float*)ws, 0);
            const size_t total = m * n, cap = (size_t)ctx->num_sms * 8, fb = (total + 255) / 256;
            const unsigned fgrid = (unsigned)(fb < cap ? fb : cap);
            if (accumulate) SL_LAUNCH(ctx, (skinny_fold_kernel<true>), fgrid, 256, 0, total, nsplit, (const float*)ws, c);
            else SL_LAUNCH(ctx, (skinny_fold_kernel<false>), fgrid, 256, 0, total, nsplit, (const float*)ws, c);
            return SL_OK;
        }
    }
    if (!trans_a && trans_b && k >= 1 && k <= 16 && n >= 64 && m >= 64 && n % 4 == 0 && al && (!mask_src || sl_aligned16(mask_src))) {
        const int kp = (int)((k + 1) & ~size_t(1));
        const unsigned gx = (unsigned)((n / 4 + 255) / 256);
        // about four waves of CTAs, slabs of 64 .. 1024 rows
        size_t gy = ((size_t)ctx->num_sms * 8 * 4 + gx - 1) / gx;
        size_t rpb = (m + gy - 1) / gy;
        rpb = rpb < 64 ? 64 : (rpb > 1024 ? 1024 : rpb);
        gy = (m + rpb - 1) / rpb;
        if (gy > 65535) { rpb = (m + 65534) / 65535; gy = (m + rpb - 1) / rpb; }
        const size_t smem = rpb * (size_t)(kp + 2) * 4;
        if (smem <= 96 * 1024) {
            const dim3 grid(gx, (unsigned)gy, 1);
#define SK_NT(ACCV, MASKV)                                                                                                              \
    SK_DISPATCH(kp, {                                                                                                                   \
        SL_CUDA(ctx, cudaFuncSetAttribute(skinny_nt_kernel<NP, ACCV, MASKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
        SL_LAUNCH(ctx, (skinny_nt_kernel<NP, ACCV, MASKV>), grid, 256, smem, m, n, k, rpb, a, b, c, mask_src, mask_bits);               \
    })
            if (accumulate && mask_src) SK_NT(true, 1)
            else if (accumulate && mask_bits) SK_NT(true, 2)
            else if (accumulate) SK_NT(true, 0)
            else if (mask_src) SK_NT(false, 1)
            else if (mask_bits) SK_NT(false, 2)
            else SK_NT(false, 0)
#undef SK_NT
            if ((mask_src || mask_bits) && mask_done) *mask_done = 1;
            return SL_OK;
        }
    }
    return sl_gemm_skinny_f32_v1(ctx, trans_a, trans_b, m, n, k, a, b, c, accumulate);
}
