// gemm_simt.cu — CUDA-core gemm: c[m x n] (=|+=) op(a) * op(b) for f32 / f64 / i32, any shape.
//
// Role: (1) the gemm of the dtypes tensor cores do not serve (the reference's unit tests run gemm in f64 and i32),
// (2) small / skinny problems where a 128-wide tcgen05 tile would be mostly padding (sine_net's 1000x64x64, the
// 10-class head of the nn.rs MLP), (3) SL_GEMM_SIMT mode: a per-element sequential-k `acc = acc + a*b`
// (no FMA contraction: built with -fmad=false), i.e. bit-identical to the CPU oracle's restatement, used by the
// parity tests to separate "layout / indexing" errors from "TF32 rounding" in the tcgen05 path.
//
// 64x64 output tile per 256-thread CTA, 4x4 micro-tile per thread, BK = 16, operands staged through shared memory.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T, bool TA, bool TB, bool ACC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(size_t m, size_t n, size_t k, const T* __restrict__ A, const T* __restrict__ B, T* C) {
    __shared__ T As[BK][BM + 4];
    __shared__ T Bs[BK][BN + 4];
    const int t = threadIdx.x;
    const int tx = t % 16, ty = t / 16;
    const size_t m0 = (size_t)blockIdx.y * BM, n0 = (size_t)blockIdx.x * BN;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

    for (size_t k0 = 0; k0 < k; k0 += BK) {
#pragma unroll
        for (int l = 0; l < (BM * BK) / 256; ++l) {
            const int e = t + l * 256;
            int i, p;
            if (TA) { i = e % BM; p = e / BM; } else { p = e % BK; i = e / BK; }
            const size_t gi = m0 + i, gp = k0 + p;
            T v = T(0);
            if (gi < m && gp < k) v = TA ? A[gp * m + gi] : A[gi * k + gp];
            As[p][i] = v;
        }
#pragma unroll
        for (int l = 0; l < (BN * BK) / 256; ++l) {
            const int e = t + l * 256;
            int j, p;
            if (TB) { p = e % BK; j = e / BK; } else { j = e % BN; p = e / BN; }
            const size_t gj = n0 + j, gp = k0 + p;
            T v = T(0);
            if (gj < n && gp < k) v = TB ? B[gj * k + gp] : B[gp * n + gj];
            Bs[p][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < BK; ++p) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[p][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[p][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = acc[i][j] + a[i] * b[j];  // zero-padded k tail adds +0 exactly
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const size_t gi = m0 + ty * 4 + i;
        if (gi >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const size_t gj = n0 + tx * 4 + j;
            if (gj < n) {
                if (ACC) C[gi * n + gj] += acc[i][j];
                else C[gi * n + gj] = acc[i][j];
            }
        }
    }
}

}  // namespace

template <typename T>
int sl_gemm_simt_t(sl_ctx* ctx, int trans_a, int trans_b, size_t m, size_t n, size_t k, const T* a, const T* b, T* c, int accumulate) {
    dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), 1);
    if (grid.y > 65535) return sl_set_error(ctx, SL_ERR_UNSUPPORTED, "gemm_simt: m too large (%zu)", m);
#define GO(TA, TB, AC) SL_LAUNCH(ctx, (gemm_simt_kernel<T, TA, TB, AC>), grid, 256, 0, m, n, k, a, b, c)
    if (!trans_a && !trans_b) { if (accumulate) GO(false, false, true); else GO(false, false, false); }
    else if (!trans_a && trans_b) { if (accumulate) GO(false, true, true); else GO(false, true, false); }
    else if (trans_a && !trans_b) { if (accumulate) GO(true, false, true); else GO(true, false, false); }
    else { if (accumulate) GO(true, true, true); else GO(true, true, false); }
#undef GO
    return SL_OK;
}

int sl_gemm_simt(sl_ctx* ctx, int dtype, int trans_a, int trans_b, size_t m, size_t n, size_t k, const void* a, const void* b, void* c,
                 int accumulate) {
    SL_DISPATCH_DTYPE(ctx, dtype, T, return sl_gemm_simt_t<T>(ctx, trans_a, trans_b, m, n, k, (const T*)a, (const T*)b, (T*)c, accumulate));
    return SL_OK;
}

// =====================================================================================================================
// Skinny shapes (one dimension <= 16): the 10-class head of the nn.rs MLP.  These are HBM-bound (they read or write one
// big matrix once), so they get bandwidth-shaped CUDA-core kernels instead of a 64x64 tile that is 84 % padding.
//   nn_skinny:  C[m x n] = A[m x k] * B[k x n],    n <= 16  (z3 = a2 * W3)          reads A once
//   tn_skinny:  C[m x n] = A[k x m]^T * B[k x n],  n <= 16  (dW3 = a2^T * dz3)      reads A once, split over k + ordered fold
//   nt_skinny:  C[m x n] = A[m x k] * B[n x k]^T,  k <= 16  (da2 = dz3 * W3^T)      writes C once
// Summation order differs from the sequential-k oracle (tolerance: K-scaled); SL_GEMM_SIMT keeps the bit-exact kernel.
// =====================================================================================================================
namespace {

constexpr int SK_N = 16;  // padded skinny extent

// ---- nn_skinny: warp handles 4 rows, lanes split k in float4 steps, B chunk [KC x 16] in shared memory
template <bool ACC>
__global__ void __launch_bounds__(256) gemm_nn_skinny_kernel(size_t m, size_t n, size_t k, const float* __restrict__ A, const float* __restrict__ B,
                                                             float* C) {
    constexpr int KC = 512;     // k per shared chunk (32 KB of B)
    constexpr int RW = 4;       // rows per warp
    __shared__ __align__(16) float Bs[KC][SK_N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t rows_per_block = 8 * RW;
    for (size_t r0 = (size_t)blockIdx.x * rows_per_block; r0 < m; r0 += (size_t)gridDim.x * rows_per_block) {
        const size_t rbase = r0 + (size_t)warp * RW;
        float acc[RW][SK_N];
#pragma unroll
        for (int i = 0; i < RW; ++i)
#pragma unroll
            for (int j = 0; j < SK_N; ++j) acc[i][j] = 0.f;
        for (size_t k0 = 0; k0 < k; k0 += KC) {
            __syncthreads();
            for (int e = threadIdx.x; e < KC * SK_N; e += 256) {
                const int kk = e / SK_N, j = e % SK_N;
                Bs[kk][j] = (k0 + kk < k && (size_t)j < n) ? B[(k0 + kk) * n + j] : 0.f;
            }
            __syncthreads();
            const int kc = (int)((k - k0) < (size_t)KC ? (k - k0) : KC);
            // lanes take consecutive k (coalesced 128-byte row segments of A, conflict-free 64-byte rows of Bs); 4 k-steps in flight
            for (int kk = lane; kk < kc; kk += 128) {
                float a[RW][4];
#pragma unroll
                for (int i = 0; i < RW; ++i) {
                    const size_t r = rbase + i;
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[i][q] = (r < m && kk + q * 32 < kc) ? __ldg(A + r * k + k0 + kk + q * 32) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (kk + q * 32 < kc) {
                        const float4* bp = reinterpret_cast<const float4*>(&Bs[kk + q * 32][0]);
                        const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
                        const float bv[SK_N] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
                        for (int i = 0; i < RW; ++i)
#pragma unroll
                            for (int j = 0; j < SK_N; ++j) acc[i][j] = fmaf(a[i][q], bv[j], acc[i][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < RW; ++i)
#pragma unroll
            for (int j = 0; j < SK_N; ++j) {
                float v = acc[i][j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[i][j] = v;
            }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                const size_t r = rbase + i;
                if (r < m)
#pragma unroll
                    for (int j = 0; j < SK_N; ++j)
                        if ((size_t)j < n) {
                            if (ACC) C[r * n + j] += acc[i][j];
                            else C[r * n + j] = acc[i][j];
                        }
            }
        }
    }
}

// ---- tn_skinny: thread owns 4 columns of A (= 4 rows of C); block walks a slice of k; partial[split][m][n]
__global__ void __launch_bounds__(128) gemm_tn_skinny_partial_kernel(size_t m, size_t n, size_t k, size_t k_per_split, const float* __restrict__ A,
                                                                     const float* __restrict__ B, float* __restrict__ partial) {
    constexpr int RC = 64;  // rows of B staged per chunk
    __shared__ __align__(16) float Bs[RC][SK_N];
    const size_t c0 = ((size_t)blockIdx.x * 128 + threadIdx.x) * 4;
    const size_t kbeg = (size_t)blockIdx.y * k_per_split;
    const size_t kend = kbeg + k_per_split < k ? kbeg + k_per_split : k;
    float acc[4][SK_N];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < SK_N; ++j) acc[i][j] = 0.f;
    const bool vec = (m % 4 == 0) && c0 + 4 <= m;
    for (size_t r0 = kbeg; r0 < kend; r0 += RC) {
        __syncthreads();
        for (int e = threadIdx.x; e < RC * SK_N; e += 128) {
            const int rr = e / SK_N, j = e % SK_N;
            Bs[rr][j] = (r0 + rr < kend && (size_t)j < n) ? B[(r0 + rr) * n + j] : 0.f;
        }
        __syncthreads();
        const int rc = (int)((kend - r0) < (size_t)RC ? (kend - r0) : RC);
        if (c0 < m) {
#pragma unroll 4
            for (int rr = 0; rr < rc; ++rr) {
                float a[4];
                if (vec) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(A + (r0 + rr) * m + c0));
                    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = c0 + i < m ? __ldg(A + (r0 + rr) * m + c0 + i) : 0.f;
                }
                const float4* bp = reinterpret_cast<const float4*>(&Bs[rr][0]);
                const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
                const float bv[SK_N] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < SK_N; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
            }
        }
    }
    if (c0 < m) {
        float* p = partial + (size_t)blockIdx.y * m * n;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (c0 + i < m)
#pragma unroll
                for (int j = 0; j < SK_N; ++j)
                    if ((size_t)j < n) p[(c0 + i) * n + j] = acc[i][j];
    }
}

template <bool ACC>
__global__ void __launch_bounds__(256) gemm_skinny_fold_kernel(size_t total, size_t nsplit, const float* __restrict__ partial, float* C) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float s = partial[i];
        for (size_t p = 1; p < nsplit; ++p) s += partial[p * total + i];
        if (ACC) C[i] += s;
        else C[i] = s;
    }
}

// ---- nt_skinny: thread owns 4 output columns (their B rows live in registers) and walks down the rows
template <bool ACC>
__global__ void __launch_bounds__(256) gemm_nt_skinny_kernel(size_t m, size_t n, size_t k, const float* __restrict__ A, const float* __restrict__ B,
                                                             float* C) {
    const size_t j0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= n) return;
    float b[4][SK_N];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int q = 0; q < SK_N; ++q) b[e][q] = (j0 + e < n && (size_t)q < k) ? __ldg(B + (j0 + e) * k + q) : 0.f;
    const bool vec = (n % 4 == 0) && j0 + 4 <= n && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
    for (size_t r = blockIdx.y; r < m; r += gridDim.y) {
        float a[SK_N];
#pragma unroll
        for (int q = 0; q < SK_N; ++q) a[q] = (size_t)q < k ? __ldg(A + r * k + q) : 0.f;  // warp-uniform broadcast
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < SK_N; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fmaf(a[q], b[e][q], o[e]);
        float* crow = C + r * n + j0;
        if (vec) {
            float4 v = make_float4(o[0], o[1], o[2], o[3]);
            if (ACC) {
                const float4 c = *reinterpret_cast<const float4*>(crow);
                v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
            }
            *reinterpret_cast<float4*>(crow) = v;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (j0 + e < n) {
                    if (ACC) crow[e] += o[e];
                    else crow[e] = o[e];
                }
        }
    }
}

}  // namespace

// returns SL_OK when a skinny kernel handled the call, 1 when the shape is not skinny (caller falls through)
// v1 kernels: fallback of gemm_skinny.cu for the shapes its kernels do not take (unaligned, k % 4 != 0, B too large for shared memory)
int sl_gemm_skinny_f32_v1(sl_ctx* ctx, int trans_a, int trans_b, size_t m, size_t n, size_t k, const float* a, const float* b, float* c,
                          int accumulate) {
    const size_t cap = (size_t)ctx->num_sms * 8;
    if (!trans_a && !trans_b && n <= SK_N && k >= 256 && m >= 64) {
        size_t blocks = (m + 31) / 32;
        unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
        if (accumulate) SL_LAUNCH(ctx, (gemm_nn_skinny_kernel<true>), grid, 256, 0, m, n, k, a, b, c);
        else SL_LAUNCH(ctx, (gemm_nn_skinny_kernel<false>), grid, 256, 0, m, n, k, a, b, c);
        return SL_OK;
    }
    if (trans_a && !trans_b && n <= SK_N && k >= 1024 && m >= 64) {
        const unsigned gx = (unsigned)((m + 511) / 512);
        size_t nsplit = (cap * 2 + gx - 1) / gx;
        const size_t max_split = (k + 255) / 256;
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
        const size_t k_per_split = ((k + nsplit - 1) / nsplit + 63) / 64 * 64;
        nsplit = (k + k_per_split - 1) / k_per_split;
        void* ws = nullptr;
        int rc = sl_ws_reserve(ctx, nsplit * m * n * sizeof(float), &ws);
        if (rc != SL_OK) return rc;
        SL_LAUNCH(ctx, gemm_tn_skinny_partial_kernel, dim3(gx, (unsigned)nsplit, 1), 128, 0, m, n, k, k_per_split, a, b, (float*)ws);
        const size_t total = m * n;
        size_t fb = (total + 255) / 256;
        unsigned fgrid = (unsigned)(fb < cap ? fb : cap);
        if (accumulate) SL_LAUNCH(ctx, (gemm_skinny_fold_kernel<true>), fgrid, 256, 0, total, nsplit, (const float*)ws, c);
        else SL_LAUNCH(ctx, (gemm_skinny_fold_kernel<false>), fgrid, 256, 0, total, nsplit, (const float*)ws, c);
        return SL_OK;
    }
    if (!trans_a && trans_b && k <= SK_N && n >= 64 && m >= 64) {
        const unsigned gx = (unsigned)((n + 1023) / 1024);
        size_t gy = (cap + gx - 1) / gx;
        if (gy > m) gy = m;
        if (gy > 65535) gy = 65535;
        if (accumulate) SL_LAUNCH(ctx, (gemm_nt_skinny_kernel<true>), dim3(gx, (unsigned)gy, 1), 256, 0, m, n, k, a, b, c);
        else SL_LAUNCH(ctx, (gemm_nt_skinny_kernel<false>), dim3(gx, (unsigned)gy, 1), 256, 0, m, n, k, a, b, c);
        return SL_OK;
    }
    return 1;
}
