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
