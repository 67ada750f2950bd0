// transpose.cu — family T: out[c*rows + r] (=|+=) x[r*cols + c], bit-exact for every dtype (pure data movement),
// plus the small "next-row" ops diagflat / onehot (+grads) and the accuracy counter of examples/nn.rs:195-211.
//
// 64x64 element tiles staged through padded shared memory: global reads and writes are both 128-byte coalesced
// row segments (the reference's CPU loop and its OpenCL kernel both write with stride `rows`).  8 B/elem (SET),
// 12 B/elem (ACC).
#include "common.cuh"

namespace {

constexpr int TT = 64;      // tile edge
constexpr int TROWS = 8;    // blockDim.y ; blockDim.x = 32, each thread moves 2 columns x 8 rows per phase

template <typename T, bool ACC>
__global__ void __launch_bounds__(256) transpose_kernel(size_t rows, size_t cols, const T* __restrict__ x, T* out) {
    __shared__ T tile[TT][TT + 1];
    const size_t tiles_c = (cols + TT - 1) / TT;
    const size_t tiles_r = (rows + TT - 1) / TT;
    const size_t ntiles = tiles_c * tiles_r;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t tr = t / tiles_c, tc = t % tiles_c;
        const size_t r0 = tr * TT, c0 = tc * TT;
        // load: rows r0.., contiguous along c
#pragma unroll
        for (int i = 0; i < TT; i += TROWS) {
            const size_t r = r0 + threadIdx.y + i;
#pragma unroll
            for (int j = 0; j < TT; j += 32) {
                const size_t c = c0 + threadIdx.x + j;
                if (r < rows && c < cols) tile[threadIdx.y + i][threadIdx.x + j] = __ldg(x + r * cols + c);
            }
        }
        __syncthreads();
        // store: out rows are the input columns c0.., contiguous along r
#pragma unroll
        for (int i = 0; i < TT; i += TROWS) {
            const size_t c = c0 + threadIdx.y + i;
#pragma unroll
            for (int j = 0; j < TT; j += 32) {
                const size_t r = r0 + threadIdx.x + j;
                if (r < rows && c < cols) {
                    const T v = tile[threadIdx.x + j][threadIdx.y + i];
                    if (ACC) out[c * rows + r] += v;
                    else out[c * rows + r] = v;
                }
            }
        }
        __syncthreads();
    }
}

// 128-bit variant: rows and cols multiples of VEC, 16-byte aligned pointers.  Loads and stores are both LDG/STG.128 on
// contiguous row segments; the 64x64 tile lives in padded shared memory and is read back column-wise (4 scalar LDS per
// 128-bit store).  Same arithmetic-free data movement => bit-exact.
template <typename T, bool ACC>
__global__ void __launch_bounds__(256) transpose_vec_kernel(size_t rows, size_t cols, const T* __restrict__ x, T* out) {
    constexpr int V = Pack<T>::N;
    constexpr int PR = TT / V;              // packs per tile row
    constexpr int RP = 256 / PR;            // tile rows covered per pass
    __shared__ T tile[TT][TT + 1];
    const size_t tiles_c = (cols + TT - 1) / TT;
    const size_t tiles_r = (rows + TT - 1) / TT;
    const size_t ntiles = tiles_c * tiles_r;
    const int t = threadIdx.y * 32 + threadIdx.x;
    const int pk = t % PR, rr = t / PR;
    for (size_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
        const size_t r0 = (tix / tiles_c) * TT, c0 = (tix % tiles_c) * TT;
#pragma unroll
        for (int i = 0; i < TT; i += RP) {
            const size_t r = r0 + rr + i, c = c0 + (size_t)pk * V;
            if (r < rows && c < cols) {
                const Pack<T> v = ld_stream(x + r * cols + c);
#pragma unroll
                for (int e = 0; e < V; ++e) tile[rr + i][pk * V + e] = v.v[e];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < TT; i += RP) {
            const size_t c = c0 + rr + i, r = r0 + (size_t)pk * V;   // output row = input column c; V consecutive input rows
            if (c < cols && r < rows) {
                Pack<T> v;
#pragma unroll
                for (int e = 0; e < V; ++e) v.v[e] = tile[pk * V + e][rr + i];
                T* dst = out + c * rows + r;
                if (ACC) {
                    const Pack<T> o = ld_pack(dst);
#pragma unroll
                    for (int e = 0; e < V; ++e) v.v[e] += o.v[e];
                    st_pack(dst, v);
                } else {
                    st_stream(dst, v);
                }
            }
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void diagflat_kernel(size_t n, const T* __restrict__ x, T* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i * n + i] = x[i];
}
template <typename T>
__global__ void diagflat_grad_kernel(size_t n, T* xg, const T* __restrict__ og) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) xg[i] += og[i * n + i];
}
template <typename T>
__global__ void onehot_kernel(size_t n, size_t hc, const T* __restrict__ classes, T* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i * hc + (size_t)classes[i]] = T(1);
}
template <typename T>
__global__ void onehot_grad_kernel(size_t n, size_t hc, const T* __restrict__ classes, T* cg, const T* __restrict__ og) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        cg[i] += og[i * hc + (size_t)classes[i]];
}

// examples/nn.rs:195-211: argmax with a strict `>` scan from column 0, compared with the label
template <typename T>
__global__ void __launch_bounds__(256) count_correct_kernel(size_t rows, size_t cols, const T* __restrict__ preds,
                                                            const int32_t* __restrict__ labels, int32_t* count) {
    int local = 0;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
        const T* row = preds + r * cols;
        T mx = row[0];
        size_t mi = 0;
        for (size_t c = 1; c < cols; ++c)
            if (row[c] > mx) {
                mx = row[c];
                mi = c;
            }
        local += ((size_t)labels[r] == mi) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);  // integer atomics: order-independent
}

static unsigned grid1d(sl_ctx* ctx, size_t n) {
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)ctx->num_sms * 8;
    return (unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

extern "C" {

int sl_transpose(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out, int accumulate) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    if (rows == 0 || cols == 0) return SL_OK;
    SL_REQUIRE(ctx, x && out && x != out, "NULL or aliased pointer");
    const size_t ntiles = ((rows + TT - 1) / TT) * ((cols + TT - 1) / TT);
    const size_t cap = (size_t)ctx->num_sms * 8;
    const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
    const dim3 block(32, TROWS, 1);
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        const bool vec = rows % Pack<T>::N == 0 && cols % Pack<T>::N == 0 && sl_aligned16(x) && sl_aligned16(out);
        if (vec) {
            if (accumulate) SL_LAUNCH(ctx, (transpose_vec_kernel<T, true>), grid, block, 0, rows, cols, (const T*)x, (T*)out);
            else SL_LAUNCH(ctx, (transpose_vec_kernel<T, false>), grid, block, 0, rows, cols, (const T*)x, (T*)out);
        } else {
            if (accumulate) SL_LAUNCH(ctx, (transpose_kernel<T, true>), grid, block, 0, rows, cols, (const T*)x, (T*)out);
            else SL_LAUNCH(ctx, (transpose_kernel<T, false>), grid, block, 0, rows, cols, (const T*)x, (T*)out);
        }
    });
    return SL_OK;
}

int sl_diagflat(sl_ctx* ctx, int dtype, size_t n, const void* x, void* out) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    if (n == 0) return SL_OK;
    SL_REQUIRE(ctx, x && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, SL_LAUNCH(ctx, (diagflat_kernel<T>), grid1d(ctx, n), 256, 0, n, (const T*)x, (T*)out));
    return SL_OK;
}

int sl_diagflat_grad(sl_ctx* ctx, int dtype, size_t n, void* x_grad, const void* out_grad) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, x_grad);
    if (n == 0) return SL_OK;
    SL_REQUIRE(ctx, x_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, SL_LAUNCH(ctx, (diagflat_grad_kernel<T>), grid1d(ctx, n), 256, 0, n, (T*)x_grad, (const T*)out_grad));
    return SL_OK;
}

int sl_onehot(sl_ctx* ctx, int dtype, size_t n, size_t highest_class, const void* classes, void* out) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    if (n == 0) return SL_OK;
    SL_REQUIRE(ctx, classes && out && highest_class > 0, "bad argument");
    SL_DISPATCH_DTYPE(ctx, dtype, T,
                      SL_LAUNCH(ctx, (onehot_kernel<T>), grid1d(ctx, n), 256, 0, n, highest_class, (const T*)classes, (T*)out));
    return SL_OK;
}

int sl_onehot_grad(sl_ctx* ctx, int dtype, size_t n, size_t highest_class, const void* classes, void* classes_grad, const void* out_grad) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, classes_grad);
    if (n == 0) return SL_OK;
    SL_REQUIRE(ctx, classes && classes_grad && out_grad && highest_class > 0, "bad argument");
    SL_DISPATCH_DTYPE(ctx, dtype, T,
                      SL_LAUNCH(ctx, (onehot_grad_kernel<T>), grid1d(ctx, n), 256, 0, n, highest_class, (const T*)classes, (T*)classes_grad,
                                (const T*)out_grad));
    return SL_OK;
}

int sl_count_correct(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* preds, const int32_t* labels, int32_t* count_dev) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, count_dev);
    SL_REQUIRE(ctx, count_dev != nullptr, "NULL count");
    int rc = sl_clear(ctx, count_dev, sizeof(int32_t));
    if (rc != SL_OK) return rc;
    if (rows == 0 || cols == 0) return SL_OK;
    SL_REQUIRE(ctx, preds && labels, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T,
                      SL_LAUNCH(ctx, (count_correct_kernel<T>), grid1d(ctx, rows), 256, 0, rows, cols, (const T*)preds, labels, count_dev));
    return SL_OK;
}

}  // extern "C"
