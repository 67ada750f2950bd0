// rowcol.cu — broadcast ops over rows / columns (row_op, col_op and their gradients), and the reductions
// over rows / columns (sum, mean, max + argmax) with their gradients.  Families E (broadcast part) and R.
//
// Naming follows the reference: "*_rows" reduces OVER rows (output length = cols); "*_cols" reduces OVER columns
// (output length = rows).  All matrices are row-major [rows x cols].
//
// Design (HBM-bound, 16384 x 16384 f32 is the sized case):
//  * 2-D thread blocks: x runs along columns in 128-bit packs (a warp covers 512 contiguous bytes of one row),
//    y runs along rows; per-column operands (row_op's rhs, out_grad of sum_rows_grad ...) are loaded ONCE per thread
//    into registers and reused down the rows; per-row operands are a warp-uniform broadcast load.
//  * Column-direction reductions never go uncoalesced: each thread accumulates its own column pack down a slice
//    of the rows (4 independent row loads in flight), the block folds its y-threads through shared memory in a fixed
//    order, every block-row writes one partial row, and a second tiny kernel folds the partials in order.
//    No float atomics anywhere => bit-reproducible run to run.
//  * Row-direction reductions use 1 / 32 / 256 threads per row depending on cols, warp shuffles + a fixed smem tree.
//
// Algorithmic bytes per element (f32): row_op/add_row 8, add_row_mut 8, add_row_grad 8, add_row_mut_grad 4,
// sum/mean/max fwd 4, sum/mean grads 8, max_rows_grad 12.
#include <limits.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------ V-wide values (V = 1 or a full 128-bit pack)
template <typename T, int V>
struct PV {
    T v[V];
};
template <typename T, int V>
__device__ __forceinline__ PV<T, V> pv_load_stream(const T* p) {
    PV<T, V> r;
    if constexpr (V == 1) {
        r.v[0] = __ldg(p);
    } else {
        Pack<T> k = ld_stream(p);
#pragma unroll
        for (int e = 0; e < V; ++e) r.v[e] = k.v[e];
    }
    return r;
}
template <typename T, int V>
__device__ __forceinline__ PV<T, V> pv_load(const T* p) {
    PV<T, V> r;
    if constexpr (V == 1) {
        r.v[0] = *p;
    } else {
        Pack<T> k = ld_pack(p);
#pragma unroll
        for (int e = 0; e < V; ++e) r.v[e] = k.v[e];
    }
    return r;
}
template <typename T, int V>
__device__ __forceinline__ void pv_store(T* p, const PV<T, V>& x) {
    if constexpr (V == 1) {
        *p = x.v[0];
    } else {
        Pack<T> k;
#pragma unroll
        for (int e = 0; e < V; ++e) k.v[e] = x.v[e];
        st_pack(p, k);
    }
}

template <int OP, typename T>
__device__ __forceinline__ T binop(T l, T r) {
    if (OP == SL_ADD) return l + r;
    if (OP == SL_SUB) return l - r;
    if (OP == SL_MUL) return l * r;
    return l / r;
}
template <int OP, typename T>
__device__ __forceinline__ T binop_dl(T l, T r) {
    if (OP == SL_ADD || OP == SL_SUB) return T(1);
    if (OP == SL_MUL) return r;
    return T(1) / r;
}
template <int OP, typename T>
__device__ __forceinline__ T binop_dr(T l, T r) {
    if (OP == SL_ADD) return T(1);
    if (OP == SL_SUB) return -T(1);
    if (OP == SL_MUL) return l;
    return l / (-(r * r));
}

// ------------------------------------------------------------------ 2-D geometry
struct Geo2D {
    dim3 grid, block;
    int V;
    size_t colpacks;
};

// can the matrix be walked in 128-bit packs?  (cols multiple of VEC and every base pointer 16-byte aligned)
template <typename T>
static Geo2D geo2d(sl_ctx* ctx, size_t rows, size_t cols, bool vec_ok) {
    Geo2D g;
    g.V = vec_ok ? Pack<T>::N : 1;
    g.colpacks = cols / g.V;
    unsigned tx = sl_pow2_ceil((unsigned)(g.colpacks < 256 ? g.colpacks : 256));
    if (tx < 1) tx = 1;
    unsigned ty = 256 / tx;
    unsigned gx = (unsigned)((g.colpacks + tx - 1) / tx);
    size_t max_gy = (rows + ty - 1) / ty;
    size_t want = ((size_t)ctx->num_sms * 8 + gx - 1) / gx;
    if (want < 1) want = 1;
    size_t gy = max_gy < want ? max_gy : want;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    g.block = dim3(tx, ty, 1);
    g.grid = dim3(gx, (unsigned)gy, 1);
    return g;
}

template <typename T>
static bool vec_ok_cols(size_t cols, std::initializer_list<const void*> ptrs) {
    if (cols % Pack<T>::N) return false;
    for (const void* p : ptrs)
        if (p && !sl_aligned16(p)) return false;
    return true;
}

constexpr int RB = 4;  // rows in flight per thread

// ------------------------------------------------------------------ row_op: out[r,c] = lhs[r,c] op rhs[c]
template <typename T, int V, int OP>
__global__ void __launch_bounds__(256) row_op_kernel(size_t rows, size_t cols, size_t colpacks, const T* __restrict__ lhs,
                                                     const T* __restrict__ rhs, T* __restrict__ out) {
    const size_t cp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= colpacks) return;
    const PV<T, V> b = pv_load<T, V>(rhs + cp * V);
    const size_t rstride = (size_t)gridDim.y * blockDim.y;
    for (size_t r0 = (size_t)blockIdx.y * blockDim.y + threadIdx.y; r0 < rows; r0 += rstride * RB) {
        PV<T, V> a[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) a[j] = pv_load_stream<T, V>(lhs + r * cols + cp * V);
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) {
#pragma unroll
                for (int e = 0; e < V; ++e) a[j].v[e] = binop<OP>(a[j].v[e], b.v[e]);
                pv_store<T, V>(out + r * cols + cp * V, a[j]);
            }
        }
    }
}

// in place: lhs[r,c] += rhs[c]   (out aliases lhs: plain loads, not the read-only path)
template <typename T, int V>
__global__ void __launch_bounds__(256) add_row_mut_kernel(size_t rows, size_t cols, size_t colpacks, T* lhs, const T* __restrict__ rhs) {
    const size_t cp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= colpacks) return;
    const PV<T, V> b = pv_load<T, V>(rhs + cp * V);
    const size_t rstride = (size_t)gridDim.y * blockDim.y;
    for (size_t r0 = (size_t)blockIdx.y * blockDim.y + threadIdx.y; r0 < rows; r0 += rstride * RB) {
        PV<T, V> a[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) a[j] = pv_load<T, V>(lhs + r * cols + cp * V);
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) {
#pragma unroll
                for (int e = 0; e < V; ++e) a[j].v[e] += b.v[e];
                pv_store<T, V>(lhs + r * cols + cp * V, a[j]);
            }
        }
    }
}

// ------------------------------------------------------------------ col_op: out[r,c] = lhs[r,c] op rhs[r]
template <typename T, int V, int OP>
__global__ void __launch_bounds__(256) col_op_kernel(size_t rows, size_t cols, size_t colpacks, const T* __restrict__ lhs,
                                                     const T* __restrict__ rhs, T* __restrict__ out) {
    const size_t cp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= colpacks) return;
    const size_t rstride = (size_t)gridDim.y * blockDim.y;
    for (size_t r0 = (size_t)blockIdx.y * blockDim.y + threadIdx.y; r0 < rows; r0 += rstride * RB) {
        PV<T, V> a[RB];
        T b[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) {
                a[j] = pv_load_stream<T, V>(lhs + r * cols + cp * V);
                b[j] = __ldg(rhs + r);
            }
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) {
#pragma unroll
                for (int e = 0; e < V; ++e) a[j].v[e] = binop<OP>(a[j].v[e], b[j]);
                pv_store<T, V>(out + r * cols + cp * V, a[j]);
            }
        }
    }
}

// ------------------------------------------------------------------ broadcast-accumulate gradients (RMW on x_grad)
// MODE 0: xg[r,c] += og[c]                    (sum_rows_grad)
// MODE 1: xg[r,c] += factor * og[c]           (mean_rows_grad, factor = T(cols)/T(len))
// MODE 2: xg[r,c] += og[r]                    (sum_cols_grad)
// MODE 3: xg[r,c] += og[r] / T(cols)          (mean_cols_grad)
// MODE 4: xg[r,c] += dl(rhs[c]) * og2[r,c]    (row_op_grad lhs; aux = rhs, og = og2 full matrix) -- OPG selects dl
template <typename T, int V, int MODE, int OPG>
__global__ void __launch_bounds__(256) bcast_acc_kernel(size_t rows, size_t cols, size_t colpacks, T* xg, const T* __restrict__ og,
                                                        const T* __restrict__ aux, T factor) {
    const size_t cp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= colpacks) return;
    PV<T, V> colv;
    if (MODE == 0 || MODE == 1) colv = pv_load<T, V>(og + cp * V);
    if (MODE == 4) colv = pv_load<T, V>(aux + cp * V);
    const size_t rstride = (size_t)gridDim.y * blockDim.y;
    for (size_t r0 = (size_t)blockIdx.y * blockDim.y + threadIdx.y; r0 < rows; r0 += rstride * RB) {
        PV<T, V> a[RB];
        PV<T, V> g[RB];
        T s[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) {
                a[j] = pv_load<T, V>(xg + r * cols + cp * V);
                if (MODE == 2 || MODE == 3) s[j] = __ldg(og + r);
                if (MODE == 4) g[j] = pv_load_stream<T, V>(og + r * cols + cp * V);
            }
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    if (MODE == 0) a[j].v[e] += colv.v[e];
                    if (MODE == 1) a[j].v[e] += factor * colv.v[e];
                    if (MODE == 2) a[j].v[e] += s[j];
                    if (MODE == 3) a[j].v[e] += s[j] / factor;
                    if (MODE == 4) a[j].v[e] += (OPG == SL_MUL ? colv.v[e] : T(1)) * g[j].v[e];
                }
                pv_store<T, V>(xg + r * cols + cp * V, a[j]);
            }
        }
    }
}

// max_rows_grad: every (r,c) with x[r,c] == out[c] gets += og[c]; others keep their exact bits
template <typename T, int V>
__global__ void __launch_bounds__(256) max_rows_grad_kernel(size_t rows, size_t cols, size_t colpacks, const T* __restrict__ out,
                                                            const T* __restrict__ x, T* xg, const T* __restrict__ og) {
    const size_t cp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= colpacks) return;
    const PV<T, V> mx = pv_load<T, V>(out + cp * V);
    const PV<T, V> g = pv_load<T, V>(og + cp * V);
    const size_t rstride = (size_t)gridDim.y * blockDim.y;
    for (size_t r0 = (size_t)blockIdx.y * blockDim.y + threadIdx.y; r0 < rows; r0 += rstride * RB) {
        PV<T, V> a[RB];
        PV<T, V> xv[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) {
                xv[j] = pv_load_stream<T, V>(x + r * cols + cp * V);
                a[j] = pv_load<T, V>(xg + r * cols + cp * V);
            }
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const size_t r = r0 + j * rstride;
            if (r < rows) {
#pragma unroll
                for (int e = 0; e < V; ++e)
                    if (xv[j].v[e] == mx.v[e]) a[j].v[e] += g.v[e];
                pv_store<T, V>(xg + r * cols + cp * V, a[j]);
            }
        }
    }
}

// col_op_grad lhs: lg[r,c] += dl(l, rhs[r]) * og[r,c]
template <typename T, int V, int OP>
__global__ void __launch_bounds__(256) col_op_grad_lhs_kernel(size_t rows, size_t cols, size_t colpacks, const T* __restrict__ lhs,
                                                              const T* __restrict__ rhs, T* lg, const T* __restrict__ og) {
    const size_t cp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= colpacks) return;
    const size_t rstride = (size_t)gridDim.y * blockDim.y;
    for (size_t r = (size_t)blockIdx.y * blockDim.y + threadIdx.y; r < rows; r += rstride) {
        PV<T, V> a = pv_load<T, V>(lg + r * cols + cp * V);
        PV<T, V> l = pv_load_stream<T, V>(lhs + r * cols + cp * V);
        PV<T, V> g = pv_load_stream<T, V>(og + r * cols + cp * V);
        const T b = __ldg(rhs + r);
#pragma unroll
        for (int e = 0; e < V; ++e) a.v[e] += binop_dl<OP>(l.v[e], b) * g.v[e];
        pv_store<T, V>(lg + r * cols + cp * V, a);
    }
}

// ------------------------------------------------------------------ column-direction reductions (output length = cols)
// KIND 0: sum_r x[r,c]
// KIND 1: sum_r x[r,c], and copy: copy_dst[r,c] = x[r,c]           (add_row_grad: lhs_grad = out_grad fused with colsum)
// KIND 2: sum_r f(aux[r,c]) * x[r,c]   (row_op_grad rhs; x = out_grad, aux = lhs; OPG selects f: MUL -> aux, SUB -> -1, ADD -> 1)
// KIND 3: max_r x[r,c] with first row index
template <typename T, int V, int KIND, int OPG>
__global__ void __launch_bounds__(256) colreduce_partial_kernel(size_t rows, size_t cols, size_t colpacks, const T* __restrict__ x,
                                                                const T* __restrict__ aux, T* __restrict__ copy_dst,
                                                                T* __restrict__ partial, int32_t* __restrict__ partial_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PV<T, V>* sv = reinterpret_cast<PV<T, V>*>(smem_raw);
    const size_t cp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = cp < colpacks;
    PV<T, V> acc;
    int32_t idx[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        acc.v[e] = T(0);
        idx[e] = INT_MAX;
    }
    if (active) {
        if (KIND == 3) acc = pv_load<T, V>(x + cp * V);  // row 0: a valid lower bound for the running max (max/cpu.rs:41)
        const size_t rstride = (size_t)gridDim.y * blockDim.y;
        for (size_t r0 = (size_t)blockIdx.y * blockDim.y + threadIdx.y; r0 < rows; r0 += rstride * RB) {
            PV<T, V> a[RB];
            PV<T, V> b[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const size_t r = r0 + j * rstride;
                if (r < rows) {
                    a[j] = pv_load_stream<T, V>(x + r * cols + cp * V);
                    if (KIND == 2 && OPG == SL_MUL) b[j] = pv_load_stream<T, V>(aux + r * cols + cp * V);
                }
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const size_t r = r0 + j * rstride;
                if (r < rows) {
                    if (KIND == 1) pv_store<T, V>(copy_dst + r * cols + cp * V, a[j]);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        if (KIND == 0 || KIND == 1) acc.v[e] += a[j].v[e];
                        if (KIND == 2) acc.v[e] += (OPG == SL_MUL ? b[j].v[e] : (OPG == SL_SUB ? -T(1) : T(1))) * a[j].v[e];
                        if (KIND == 3) {
                            const T v = a[j].v[e];
                            if (v > acc.v[e] || (v == acc.v[e] && (int32_t)r < idx[e])) {
                                acc.v[e] = v;
                                idx[e] = (int32_t)r;
                            }
                        }
                    }
                }
            }
        }
    }
    // fold the y-threads of this block in a fixed order
    int32_t* si = reinterpret_cast<int32_t*>(smem_raw + sizeof(PV<T, V>) * blockDim.x * blockDim.y);
    const unsigned slot = threadIdx.y * blockDim.x + threadIdx.x;
    sv[slot] = acc;
    if (KIND == 3) {
#pragma unroll
        for (int e = 0; e < V; ++e) si[slot * V + e] = idx[e];
    }
    __syncthreads();
    if (threadIdx.y == 0 && active) {
        for (unsigned y = 1; y < blockDim.y; ++y) {
            const PV<T, V> o = sv[y * blockDim.x + threadIdx.x];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                if (KIND == 3) {
                    const int32_t oi = si[(y * blockDim.x + threadIdx.x) * V + e];
                    if (o.v[e] > acc.v[e] || (o.v[e] == acc.v[e] && oi < idx[e])) {
                        acc.v[e] = o.v[e];
                        idx[e] = oi;
                    }
                } else {
                    acc.v[e] += o.v[e];
                }
            }
        }
        pv_store<T, V>(partial + (size_t)blockIdx.y * cols + cp * V, acc);
        if (KIND == 3) {
#pragma unroll
            for (int e = 0; e < V; ++e) partial_idx[(size_t)blockIdx.y * cols + cp * V + e] = idx[e];
        }
    }
}

// fold the partial rows in order.  FIN 0: out = s ; 1: out += s ; 2: out = s / T(rows) (mean) ; 3: max (+idx)
template <typename T, int FIN>
__global__ void __launch_bounds__(256) colreduce_final_kernel(size_t nparts, size_t cols, const T* __restrict__ partial,
                                                              const int32_t* __restrict__ partial_idx, T* out, int32_t* idx_out,
                                                              T divisor) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    T acc = partial[c];
    int32_t ai = (FIN == 3) ? partial_idx[c] : 0;
    for (size_t p = 1; p < nparts; ++p) {
        const T v = partial[p * cols + c];
        if (FIN == 3) {
            const int32_t vi = partial_idx[p * cols + c];
            if (v > acc || (v == acc && vi < ai)) {
                acc = v;
                ai = vi;
            }
        } else {
            acc += v;
        }
    }
    if (FIN == 0) out[c] = acc;
    if (FIN == 1) out[c] += acc;
    if (FIN == 2) out[c] = acc / divisor;
    if (FIN == 3) {
        out[c] = acc;
        if (idx_out) idx_out[c] = ai;
    }
}

// ------------------------------------------------------------------ row-direction reductions (output length = rows)
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ void warp_maxidx(T& v, int32_t& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int32_t oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) {
            v = ov;
            i = oi;
        }
    }
}
__device__ __forceinline__ int32_t warp_min_i32(int32_t i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int32_t oi = __shfl_xor_sync(0xffffffffu, i, o);
        i = oi < i ? oi : i;
    }
    return i;
}

// TPR threads per row (1, 32 or 256), 256 threads per block.  KIND 0 sum, 1 mean (sum / T(cols)), 2 max (+ first idx).
template <typename T, int V, int TPR, int KIND>
__global__ void __launch_bounds__(256) rowreduce_kernel(size_t rows, size_t cols, const T* __restrict__ x, T* __restrict__ out,
                                                        int32_t* __restrict__ idx_out) {
    constexpr int ROWS_PER_BLOCK = 256 / TPR;
    __shared__ T s_val[8];
    __shared__ int32_t s_idx[8];
    const int lane = threadIdx.x % TPR;
    const int sub = threadIdx.x / TPR;
    const size_t colpacks = cols / V;
    for (size_t rbase = (size_t)blockIdx.x * ROWS_PER_BLOCK; rbase < rows; rbase += (size_t)gridDim.x * ROWS_PER_BLOCK) {
        const size_t r = rbase + sub;
        const bool active = r < rows;
        T acc = T(0);
        int32_t ai = INT_MAX;
        if (active) {
            const T* row = x + r * cols;
            if (KIND == 2) acc = __ldg(row);
            for (size_t cp0 = lane; cp0 < colpacks; cp0 += (size_t)TPR * RB) {
                PV<T, V> a[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const size_t cp = cp0 + (size_t)j * TPR;
                    if (cp < colpacks) a[j] = pv_load_stream<T, V>(row + cp * V);
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const size_t cp = cp0 + (size_t)j * TPR;
                    if (cp < colpacks) {
#pragma unroll
                        for (int e = 0; e < V; ++e) {
                            const T v = a[j].v[e];
                            if (KIND == 2) {
                                const int32_t ci = (int32_t)(cp * V + e);
                                if (v > acc || (v == acc && ci < ai)) {
                                    acc = v;
                                    ai = ci;
                                }
                            } else {
                                acc += v;
                            }
                        }
                    }
                }
            }
        }
        if (TPR >= 32) {
            if (KIND == 2) warp_maxidx(acc, ai);
            else acc = warp_sum(acc);
        }
        if (TPR == 256) {
            const int w = threadIdx.x >> 5;
            __syncthreads();  // protect s_val reuse across loop iterations
            if ((threadIdx.x & 31) == 0) {
                s_val[w] = acc;
                s_idx[w] = ai;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                acc = s_val[0];
                ai = s_idx[0];
                for (int k = 1; k < 8; ++k) {
                    if (KIND == 2) {
                        if (s_val[k] > acc || (s_val[k] == acc && s_idx[k] < ai)) {
                            acc = s_val[k];
                            ai = s_idx[k];
                        }
                    } else {
                        acc += s_val[k];
                    }
                }
            }
        }
        if (active && lane == 0) {
            if (KIND == 1) acc = acc / T(cols);
            out[r] = acc;
            if (KIND == 2 && idx_out) idx_out[r] = ai;
        }
    }
}

// max_cols_grad: first c with x[r,c] == out[r] gets += og[r]
template <typename T, int TPR>
__global__ void __launch_bounds__(256) max_cols_grad_kernel(size_t rows, size_t cols, const T* __restrict__ out, const T* __restrict__ x,
                                                            T* xg, const T* __restrict__ og, bool vec) {
    constexpr int ROWS_PER_BLOCK = 256 / TPR;
    __shared__ int32_t s_idx[8];
    const int lane = threadIdx.x % TPR;
    const int sub = threadIdx.x / TPR;
    for (size_t rbase = (size_t)blockIdx.x * ROWS_PER_BLOCK; rbase < rows; rbase += (size_t)gridDim.x * ROWS_PER_BLOCK) {
        const size_t r = rbase + sub;
        const bool active = r < rows;
        int32_t ai = INT_MAX;
        if (active) {
            const T mx = __ldg(out + r);
            const T* row = x + r * cols;
            if (vec) {
                // whole row in 128-bit packs, 4 packs in flight per thread (no early exit: the row is read exactly once, 4 B/elem)
                constexpr int V = Pack<T>::N;
                const size_t colpacks = cols / V;
                for (size_t cp0 = lane; cp0 < colpacks; cp0 += (size_t)TPR * RB) {
                    Pack<T> a[RB];
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        const size_t cp = cp0 + (size_t)j * TPR;
                        if (cp < colpacks) a[j] = ld_stream(row + cp * V);
                    }
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        const size_t cp = cp0 + (size_t)j * TPR;
                        if (cp < colpacks) {
#pragma unroll
                            for (int e = V - 1; e >= 0; --e)
                                if (a[j].v[e] == mx && (int32_t)(cp * V + e) < ai) ai = (int32_t)(cp * V + e);
                        }
                    }
                }
            } else {
                for (size_t c = lane; c < cols; c += TPR) {
                    if (__ldg(row + c) == mx) {
                        ai = (int32_t)c;
                        break;  // this thread's columns are visited in increasing order
                    }
                }
            }
        }
        if (TPR >= 32) ai = warp_min_i32(ai);
        if (TPR == 256) {
            const int w = threadIdx.x >> 5;
            __syncthreads();
            if ((threadIdx.x & 31) == 0) s_idx[w] = ai;
            __syncthreads();
            if (threadIdx.x == 0) {
                ai = s_idx[0];
                for (int k = 1; k < 8; ++k) ai = s_idx[k] < ai ? s_idx[k] : ai;
            }
        }
        if (active && lane == 0 && ai != INT_MAX) xg[r * cols + ai] += __ldg(og + r);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) max_cols_grad_idx_kernel(size_t rows, size_t cols, const int32_t* __restrict__ idx, T* xg,
                                                                const T* __restrict__ og) {
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x)
        xg[r * cols + idx[r]] += og[r];
}

// col_op_grad rhs: rg[r] += sum_c dr(l[r,c], rhs[r]) * og[r,c]
template <typename T, int TPR, int OP>
__global__ void __launch_bounds__(256) col_op_grad_rhs_kernel(size_t rows, size_t cols, const T* __restrict__ lhs, const T* __restrict__ rhs,
                                                              T* rg, const T* __restrict__ og) {
    constexpr int ROWS_PER_BLOCK = 256 / TPR;
    __shared__ T s_val[8];
    const int lane = threadIdx.x % TPR;
    const int sub = threadIdx.x / TPR;
    for (size_t rbase = (size_t)blockIdx.x * ROWS_PER_BLOCK; rbase < rows; rbase += (size_t)gridDim.x * ROWS_PER_BLOCK) {
        const size_t r = rbase + sub;
        const bool active = r < rows;
        T acc = T(0);
        if (active) {
            const T b = __ldg(rhs + r);
            for (size_t c = lane; c < cols; c += TPR) acc += binop_dr<OP>(__ldg(lhs + r * cols + c), b) * __ldg(og + r * cols + c);
        }
        if (TPR >= 32) acc = warp_sum(acc);
        if (TPR == 256) {
            const int w = threadIdx.x >> 5;
            __syncthreads();
            if ((threadIdx.x & 31) == 0) s_val[w] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                acc = s_val[0];
                for (int k = 1; k < 8; ++k) acc += s_val[k];
            }
        }
        if (active && lane == 0) rg[r] += acc;
    }
}

// ------------------------------------------------------------------ scalar reductions (whole buffer)
// KIND 0 sum, 2 max
template <typename T, int KIND>
__global__ void __launch_bounds__(256) scalar_partial_kernel(size_t n, const T* __restrict__ x, T* __restrict__ partial) {
    __shared__ T s_val[8];
    T acc = KIND == 2 ? __ldg(x) : T(0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const T v = __ldg(x + i);
        if (KIND == 2) acc = v > acc ? v : acc;
        else acc += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(0xffffffffu, acc, o);
        if (KIND == 2) acc = ov > acc ? ov : acc;
        else acc += ov;
    }
    if ((threadIdx.x & 31) == 0) s_val[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        acc = s_val[0];
        for (int k = 1; k < 8; ++k) {
            if (KIND == 2) acc = s_val[k] > acc ? s_val[k] : acc;
            else acc += s_val[k];
        }
        partial[blockIdx.x] = acc;
    }
}
// FIN 0: sum, 1: mean (sum / T(n)), 2: max
template <typename T, int FIN>
__global__ void scalar_final_kernel(size_t nparts, const T* __restrict__ partial, T* out, T divisor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        T acc = partial[0];
        for (size_t p = 1; p < nparts; ++p) {
            if (FIN == 2) acc = partial[p] > acc ? partial[p] : acc;
            else acc += partial[p];
        }
        if (FIN == 1) acc = acc / divisor;
        *out = acc;
    }
}

// ------------------------------------------------------------------ host-side launchers
#define SL_BINOP_SWITCH(op, OPC, ...)                                        \
    switch (op) {                                                            \
    case SL_ADD: { constexpr int OPC = SL_ADD; __VA_ARGS__; } break;         \
    case SL_SUB: { constexpr int OPC = SL_SUB; __VA_ARGS__; } break;         \
    case SL_MUL: { constexpr int OPC = SL_MUL; __VA_ARGS__; } break;         \
    case SL_DIV: { constexpr int OPC = SL_DIV; __VA_ARGS__; } break;         \
    default: return sl_set_error(ctx, SL_ERR_INVALID_ARG, "%s: bad binop %d", __func__, (int)(op)); \
    }

#define SL_VEC_SWITCH(g, T, VV, ...)                                         \
    if ((g).V == 1) { constexpr int VV = 1; __VA_ARGS__; }                   \
    else { constexpr int VV = Pack<T>::N; __VA_ARGS__; }

template <typename T>
int row_op_t(sl_ctx* ctx, int op, size_t rows, size_t cols, const void* lhs, const void* rhs, void* out) {
    Geo2D g = geo2d<T>(ctx, rows, cols, vec_ok_cols<T>(cols, {lhs, rhs, out}));
    if (lhs == out) {
        if (op != SL_ADD) return sl_set_error(ctx, SL_ERR_UNSUPPORTED, "in-place row_op supports ADD only");
        SL_VEC_SWITCH(g, T, VV, SL_LAUNCH(ctx, (add_row_mut_kernel<T, VV>), g.grid, g.block, 0, rows, cols, g.colpacks, (T*)out, (const T*)rhs));
        return SL_OK;
    }
    SL_BINOP_SWITCH(op, OPC, SL_VEC_SWITCH(g, T, VV, SL_LAUNCH(ctx, (row_op_kernel<T, VV, OPC>), g.grid, g.block, 0, rows, cols, g.colpacks,
                                                               (const T*)lhs, (const T*)rhs, (T*)out)));
    return SL_OK;
}

template <typename T>
int col_op_t(sl_ctx* ctx, int op, size_t rows, size_t cols, const void* lhs, const void* rhs, void* out) {
    Geo2D g = geo2d<T>(ctx, rows, cols, vec_ok_cols<T>(cols, {lhs, out}));
    SL_BINOP_SWITCH(op, OPC, SL_VEC_SWITCH(g, T, VV, SL_LAUNCH(ctx, (col_op_kernel<T, VV, OPC>), g.grid, g.block, 0, rows, cols, g.colpacks,
                                                               (const T*)lhs, (const T*)rhs, (T*)out)));
    return SL_OK;
}

template <typename T, int MODE>
int bcast_acc_t(sl_ctx* ctx, size_t rows, size_t cols, void* xg, const void* og, T factor) {
    Geo2D g = geo2d<T>(ctx, rows, cols, vec_ok_cols<T>(cols, {xg, (MODE <= 1) ? og : nullptr}));
    SL_VEC_SWITCH(g, T, VV, SL_LAUNCH(ctx, (bcast_acc_kernel<T, VV, MODE, 0>), g.grid, g.block, 0, rows, cols, g.colpacks, (T*)xg,
                                      (const T*)og, (const T*)nullptr, factor));
    return SL_OK;
}

// column reduction driver.  fin: 0 SET, 1 ACC, 2 mean, 3 max
template <typename T, int KIND, int OPG>
int colreduce_t(sl_ctx* ctx, size_t rows, size_t cols, const void* x, const void* aux, void* copy_dst, void* out, int32_t* idx_out, int fin) {
    Geo2D g = geo2d<T>(ctx, rows, cols, vec_ok_cols<T>(cols, {x, aux, copy_dst}));
    const size_t nparts = g.grid.y;
    const size_t part_bytes = nparts * cols * sizeof(T);
    const size_t part_bytes_al = (part_bytes + 255) & ~size_t(255);
    void* ws = nullptr;
    int rc = sl_ws_reserve(ctx, part_bytes_al + (KIND == 3 ? nparts * cols * sizeof(int32_t) : 0), &ws);
    if (rc != SL_OK) return rc;
    T* partial = (T*)ws;
    int32_t* partial_idx = KIND == 3 ? (int32_t*)((char*)ws + part_bytes_al) : nullptr;
    const size_t threads = (size_t)g.block.x * g.block.y;
    SL_VEC_SWITCH(g, T, VV, {
        size_t smem = threads * sizeof(PV<T, VV>) + (KIND == 3 ? threads * VV * sizeof(int32_t) : 0);
        SL_LAUNCH(ctx, (colreduce_partial_kernel<T, VV, KIND, OPG>), g.grid, g.block, smem, rows, cols, g.colpacks, (const T*)x,
                  (const T*)aux, (T*)copy_dst, partial, partial_idx);
    });
    const unsigned fgrid = (unsigned)((cols + 255) / 256);
    switch (fin) {
    case 0: SL_LAUNCH(ctx, (colreduce_final_kernel<T, 0>), fgrid, 256, 0, nparts, cols, partial, partial_idx, (T*)out, idx_out, T(1)); break;
    case 1: SL_LAUNCH(ctx, (colreduce_final_kernel<T, 1>), fgrid, 256, 0, nparts, cols, partial, partial_idx, (T*)out, idx_out, T(1)); break;
    case 2: SL_LAUNCH(ctx, (colreduce_final_kernel<T, 2>), fgrid, 256, 0, nparts, cols, partial, partial_idx, (T*)out, idx_out, (T)rows); break;
    default: SL_LAUNCH(ctx, (colreduce_final_kernel<T, 3>), fgrid, 256, 0, nparts, cols, partial, partial_idx, (T*)out, idx_out, T(1)); break;
    }
    return SL_OK;
}

template <typename T, int KIND>
int rowreduce_t(sl_ctx* ctx, size_t rows, size_t cols, const void* x, void* out, int32_t* idx_out) {
    const bool vec = (cols % Pack<T>::N == 0) && sl_aligned16(x) && cols >= 64;
    const size_t cap = (size_t)ctx->num_sms * 8;
    if (cols <= 32) {
        size_t blocks = (rows + 255) / 256;
        unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
        SL_LAUNCH(ctx, (rowreduce_kernel<T, 1, 1, KIND>), grid, 256, 0, rows, cols, (const T*)x, (T*)out, idx_out);
    } else if (cols <= 4096) {
        size_t blocks = (rows + 7) / 8;
        unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
        if (vec) SL_LAUNCH(ctx, (rowreduce_kernel<T, Pack<T>::N, 32, KIND>), grid, 256, 0, rows, cols, (const T*)x, (T*)out, idx_out);
        else SL_LAUNCH(ctx, (rowreduce_kernel<T, 1, 32, KIND>), grid, 256, 0, rows, cols, (const T*)x, (T*)out, idx_out);
    } else {
        unsigned grid = (unsigned)(rows < cap ? rows : cap);
        if (vec) SL_LAUNCH(ctx, (rowreduce_kernel<T, Pack<T>::N, 256, KIND>), grid, 256, 0, rows, cols, (const T*)x, (T*)out, idx_out);
        else SL_LAUNCH(ctx, (rowreduce_kernel<T, 1, 256, KIND>), grid, 256, 0, rows, cols, (const T*)x, (T*)out, idx_out);
    }
    return SL_OK;
}

template <typename T, int KIND>
int scalar_reduce_t(sl_ctx* ctx, const void* x, size_t n, void* out_dev) {
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)ctx->num_sms * 8;
    unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
    if (grid < 1) grid = 1;
    void* ws = nullptr;
    int rc = sl_ws_reserve(ctx, grid * sizeof(T), &ws);
    if (rc != SL_OK) return rc;
    constexpr int PK = KIND == 2 ? 2 : 0;
    SL_LAUNCH(ctx, (scalar_partial_kernel<T, PK>), grid, 256, 0, n, (const T*)x, (T*)ws);
    SL_LAUNCH(ctx, (scalar_final_kernel<T, KIND>), 1, 32, 0, (size_t)grid, (const T*)ws, (T*)out_dev, (T)n);
    return SL_OK;
}

#define SL_TPR_SWITCH(cols, TPRV, ...)                                       \
    if ((cols) <= 32) { constexpr int TPRV = 1; __VA_ARGS__; }               \
    else if ((cols) <= 4096) { constexpr int TPRV = 32; __VA_ARGS__; }       \
    else { constexpr int TPRV = 256; __VA_ARGS__; }

static unsigned tpr_grid(sl_ctx* ctx, size_t rows, size_t cols) {
    const size_t rpb = cols <= 32 ? 256 : (cols <= 4096 ? 8 : 1);
    size_t blocks = (rows + rpb - 1) / rpb;
    const size_t cap = (size_t)ctx->num_sms * 8;
    return (unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

#define SL_COMMON_2D_CHECKS()                                    \
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");              \
    if (rows == 0 || cols == 0) return SL_OK;

extern "C" {

int sl_row_op(sl_ctx* ctx, int dtype, int binop, size_t rows, size_t cols, const void* lhs, const void* rhs, void* out) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, out);
    SL_REQUIRE(ctx, lhs && rhs && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return row_op_t<T>(ctx, binop, rows, cols, lhs, rhs, out));
    return SL_OK;
}

int sl_add_row(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* lhs, const void* rhs, void* out) {
    return sl_row_op(ctx, dtype, SL_ADD, rows, cols, lhs, rhs, out);
}

int sl_add_row_mut(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* lhs, const void* rhs) {
    return sl_row_op(ctx, dtype, SL_ADD, rows, cols, lhs, rhs, lhs);
}

int sl_add_row_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* lhs_grad, void* rhs_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, lhs_grad, rhs_grad);
    SL_REQUIRE(ctx, lhs_grad && rhs_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (colreduce_t<T, 1, 0>(ctx, rows, cols, out_grad, nullptr, lhs_grad, rhs_grad, nullptr, 1)));
    return SL_OK;
}

int sl_add_row_mut_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* rhs_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, rhs_grad);
    SL_REQUIRE(ctx, rhs_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (colreduce_t<T, 0, 0>(ctx, rows, cols, out_grad, nullptr, nullptr, rhs_grad, nullptr, 1)));
    return SL_OK;
}

int sl_row_op_grad(sl_ctx* ctx, int dtype, int binop, size_t rows, size_t cols, const void* lhs, const void* rhs, void* lhs_grad,
                   void* rhs_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, lhs_grad, rhs_grad);
    SL_REQUIRE(ctx, out_grad != nullptr, "NULL out_grad");
    SL_REQUIRE(ctx, binop == SL_ADD || binop == SL_SUB || binop == SL_MUL, "row_op_grad supports ADD/SUB/MUL");
    SL_REQUIRE(ctx, binop != SL_MUL || (lhs && rhs), "MUL needs lhs and rhs");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        if (lhs_grad) {
            Geo2D g = geo2d<T>(ctx, rows, cols, vec_ok_cols<T>(cols, {lhs_grad, out_grad, rhs}));
            if (binop == SL_MUL) {
                SL_VEC_SWITCH(g, T, VV, SL_LAUNCH(ctx, (bcast_acc_kernel<T, VV, 4, SL_MUL>), g.grid, g.block, 0, rows, cols, g.colpacks,
                                                  (T*)lhs_grad, (const T*)out_grad, (const T*)rhs, T(0)));
            } else {
                SL_VEC_SWITCH(g, T, VV, SL_LAUNCH(ctx, (bcast_acc_kernel<T, VV, 4, SL_ADD>), g.grid, g.block, 0, rows, cols, g.colpacks,
                                                  (T*)lhs_grad, (const T*)out_grad, (const T*)out_grad, T(0)));
            }
        }
        if (rhs_grad) {
            int rc;
            if (binop == SL_MUL) rc = colreduce_t<T, 2, SL_MUL>(ctx, rows, cols, out_grad, lhs, nullptr, rhs_grad, nullptr, 1);
            else if (binop == SL_SUB) rc = colreduce_t<T, 2, SL_SUB>(ctx, rows, cols, out_grad, nullptr, nullptr, rhs_grad, nullptr, 1);
            else rc = colreduce_t<T, 2, SL_ADD>(ctx, rows, cols, out_grad, nullptr, nullptr, rhs_grad, nullptr, 1);
            if (rc != SL_OK) return rc;
        }
    });
    return SL_OK;
}

int sl_col_op(sl_ctx* ctx, int dtype, int binop, size_t rows, size_t cols, const void* lhs, const void* rhs, void* out) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, out);
    SL_REQUIRE(ctx, lhs && rhs && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return col_op_t<T>(ctx, binop, rows, cols, lhs, rhs, out));
    return SL_OK;
}

int sl_col_op_grad(sl_ctx* ctx, int dtype, int binop, size_t rows, size_t cols, const void* lhs, const void* rhs, void* lhs_grad,
                   void* rhs_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, lhs_grad, rhs_grad);
    SL_REQUIRE(ctx, lhs && rhs && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        if (lhs_grad) {
            Geo2D g = geo2d<T>(ctx, rows, cols, vec_ok_cols<T>(cols, {lhs, lhs_grad, out_grad}));
            SL_BINOP_SWITCH(binop, OPC, SL_VEC_SWITCH(g, T, VV, SL_LAUNCH(ctx, (col_op_grad_lhs_kernel<T, VV, OPC>), g.grid, g.block, 0, rows,
                                                                          cols, g.colpacks, (const T*)lhs, (const T*)rhs, (T*)lhs_grad,
                                                                          (const T*)out_grad)));
        }
        if (rhs_grad) {
            const unsigned grid = tpr_grid(ctx, rows, cols);
            SL_BINOP_SWITCH(binop, OPC, SL_TPR_SWITCH(cols, TPRV, SL_LAUNCH(ctx, (col_op_grad_rhs_kernel<T, TPRV, OPC>), grid, 256, 0, rows, cols,
                                                                            (const T*)lhs, (const T*)rhs, (T*)rhs_grad, (const T*)out_grad)));
        }
    });
    return SL_OK;
}

int sl_sum(sl_ctx* ctx, int dtype, const void* x, size_t n, void* out_dev) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out_dev);
    SL_REQUIRE(ctx, out_dev && (n == 0 || x), "NULL pointer");
    if (n == 0) return sl_clear(ctx, out_dev, sl_dtype_size(dtype));
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (scalar_reduce_t<T, 0>(ctx, x, n, out_dev)));
    return SL_OK;
}
int sl_mean(sl_ctx* ctx, int dtype, const void* x, size_t n, void* out_dev) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out_dev);
    SL_REQUIRE(ctx, out_dev && x && n > 0, "NULL pointer or empty buffer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (scalar_reduce_t<T, 1>(ctx, x, n, out_dev)));
    return SL_OK;
}
int sl_max(sl_ctx* ctx, int dtype, const void* x, size_t n, void* out_dev) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out_dev);
    SL_REQUIRE(ctx, out_dev && x && n > 0, "Buffer should contain at least an element.");  // src/ops2/max/cpu.rs:24
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (scalar_reduce_t<T, 2>(ctx, x, n, out_dev)));
    return SL_OK;
}

int sl_sum_rows(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    if (cols == 0) return SL_OK;
    if (rows == 0) return sl_clear(ctx, out, cols * sl_dtype_size(dtype));
    SL_REQUIRE(ctx, x && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (colreduce_t<T, 0, 0>(ctx, rows, cols, x, nullptr, nullptr, out, nullptr, 0)));
    return SL_OK;
}
int sl_mean_rows(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, out);
    SL_REQUIRE(ctx, x && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (colreduce_t<T, 0, 0>(ctx, rows, cols, x, nullptr, nullptr, out, nullptr, 2)));
    return SL_OK;
}
int sl_max_rows(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out, int32_t* idx_out) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, out, idx_out);
    SL_REQUIRE(ctx, x && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (colreduce_t<T, 3, 0>(ctx, rows, cols, x, nullptr, nullptr, out, idx_out, 3)));
    return SL_OK;
}
int sl_sum_cols(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, out);
    if (rows == 0) return SL_OK;
    if (cols == 0) return sl_clear(ctx, out, rows * sl_dtype_size(dtype));
    SL_REQUIRE(ctx, x && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (rowreduce_t<T, 0>(ctx, rows, cols, x, out, nullptr)));
    return SL_OK;
}
int sl_mean_cols(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, out);
    SL_REQUIRE(ctx, x && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (rowreduce_t<T, 1>(ctx, rows, cols, x, out, nullptr)));
    return SL_OK;
}
int sl_max_cols(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out, int32_t* idx_out) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, out, idx_out);
    SL_REQUIRE(ctx, x && out, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (rowreduce_t<T, 2>(ctx, rows, cols, x, out, idx_out)));
    return SL_OK;
}

int sl_sum_rows_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* x_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, x_grad);
    SL_REQUIRE(ctx, x_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (bcast_acc_t<T, 0>(ctx, rows, cols, x_grad, out_grad, T(0))));
    return SL_OK;
}
int sl_mean_rows_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* x_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, x_grad);
    SL_REQUIRE(ctx, x_grad && out_grad, "NULL pointer");
    // factor = T(cols) / T(len), computed in T exactly as src/ops2/mean/grad/cpu.rs:45-47 (integer T -> integer division)
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (bcast_acc_t<T, 1>(ctx, rows, cols, x_grad, out_grad, (T)((T)cols / (T)(rows * cols)))));
    return SL_OK;
}
int sl_sum_cols_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* x_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, x_grad);
    SL_REQUIRE(ctx, x_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (bcast_acc_t<T, 2>(ctx, rows, cols, x_grad, out_grad, T(0))));
    return SL_OK;
}
int sl_mean_cols_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* x_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, x_grad);
    SL_REQUIRE(ctx, x_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, return (bcast_acc_t<T, 3>(ctx, rows, cols, x_grad, out_grad, (T)cols)));
    return SL_OK;
}

int sl_max_rows_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* out, const void* x, void* x_grad,
                     const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, x_grad);
    SL_REQUIRE(ctx, out && x && x_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        Geo2D g = geo2d<T>(ctx, rows, cols, vec_ok_cols<T>(cols, {out, x, x_grad, out_grad}));
        SL_VEC_SWITCH(g, T, VV, SL_LAUNCH(ctx, (max_rows_grad_kernel<T, VV>), g.grid, g.block, 0, rows, cols, g.colpacks, (const T*)out,
                                          (const T*)x, (T*)x_grad, (const T*)out_grad));
    });
    return SL_OK;
}

int sl_max_cols_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* out, const void* x, void* x_grad,
                     const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, x_grad);
    SL_REQUIRE(ctx, out && x && x_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        const unsigned grid = tpr_grid(ctx, rows, cols);
        const bool vec = cols >= 64 && cols % Pack<T>::N == 0 && sl_aligned16(x);
        SL_TPR_SWITCH(cols, TPRV, SL_LAUNCH(ctx, (max_cols_grad_kernel<T, TPRV>), grid, 256, 0, rows, cols, (const T*)out, (const T*)x,
                                            (T*)x_grad, (const T*)out_grad, vec));
    });
    return SL_OK;
}

int sl_max_cols_grad_idx(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const int32_t* idx, void* x_grad, const void* out_grad) {
    SL_COMMON_2D_CHECKS();
    sl_note_writes(ctx, x_grad);
    SL_REQUIRE(ctx, idx && x_grad && out_grad, "NULL pointer");
    SL_DISPATCH_DTYPE(ctx, dtype, T, {
        size_t blocks = (rows + 255) / 256;
        const size_t cap = (size_t)ctx->num_sms * 8;
        unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
        SL_LAUNCH(ctx, (max_cols_grad_idx_kernel<T>), grid, 256, 0, rows, cols, idx, (T*)x_grad, (const T*)out_grad);
    });
    return SL_OK;
}

}  // extern "C"
