// mlp_small.cu — the whole training step of a SMALL squared-error MLP in one launch (examples/sine_net.rs:135-163 at its shipped
// sizes: 1-64-64-1 on 1000 samples).  Op by op that step is ~29 launches over matrices of at most 1000 x 64: every kernel is
// launch- and latency-bound (measured 300 us per step even replayed as a CUDA graph), while the arithmetic is 12.5 MFLOP.
//
// One thread-block cluster of 16 CTAs (8 below 256 samples) runs it: the samples are split over the CTAs, every CTA keeps the weights (and their
// transposes), its slice of every activation and its partial parameter gradients in shared memory, and
//   forward   z_l = a_{l-1} W_l + b_l ; a_l = (z_l >= 0) * z_l on hidden layers         (gemm + add_row_mut + relu, matrix.rs:181)
//   loss      sum (out - y)^2 ; d out = 2 (out - y)                                      (sub, pow(2.), mean * len, backward() seed 1)
//   backward  dW_l += a_{l-1}^T g_l ; db_l += colsum g_l ; g_{l-1} = (z_{l-1} >= 0) * (g_l W_l^T)
//   exchange  the partial gradients are summed over distributed shared memory in rank order (deterministic)
//   SGD       w -= grad * lr                                                             (sine_net.rs:108-116)
// never leave the SM.  All three contractions are FP32 FMA register tiles (4 x 4 per thread) fed by 128-bit shared-memory loads:
// activations / activation gradients are stored feature-major ([feature][sample], padded stride) so that forward, input-gradient and
// weight-gradient contractions all read float4s; weights live twice (W and W^T).
// Summation orders differ from the op-by-op tape (which is the bit-exact mirror of the oracle), so this path is tolerance-compared
// with it (tests/test_gpu_mlp.py::test_small_step_*).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int SM_MAXL = 4;      // Linear layers
constexpr int SM_MAXW = 64;     // widest layer
constexpr int SM_MAXC = 16;     // CTAs per cluster: 16 (non-portable size, one GPC) for >= 256 samples, else 8
constexpr int SM_NT = 256;      // threads per CTA == max (I/4) * (O/4) weight-gradient tiles
constexpr int SM_SCRATCH = 16 * SM_NT;   // floats per split-reduction scratch region

struct SmallMlpArgs {
    int L, batch, rows_per_cta, rt, rs;          // rt: samples per pass (multiple of 4), rs: padded sample stride of the feature-major arrays
    int dims[SM_MAXL + 1], dp[SM_MAXL + 1];      // widths and widths rounded up to 4
    int w_off[SM_MAXL], b_off[SM_MAXL];          // float offsets into the flat parameter / gradient buffers
    int sW[SM_MAXL], sWt[SM_MAXL], sB[SM_MAXL], sGW[SM_MAXL], sGB[SM_MAXL];   // shared-memory float offsets (W^T rows are padded by 4: ldt = dp[l] + 4)
    int sA[SM_MAXL + 1], sM[SM_MAXL], sG[2], sRed, sScr[2], sZeroEnd;         // [0, sZeroEnd): parameters + partial gradients, zero-filled first
    float lr;
    const float *x, *y;
    float *params, *grads, *loss_sum;
};

// C[r][n] = sum_k At[k][r] * B[k][n] on 4 (samples) x 4 (n) register tiles; At feature-major with stride rs, B row-major with stride ldb.
// When there are fewer tiles than threads the contraction is split over k between thread groups and the partial tiles are summed in
// group order through `scratch` (thin layers: 1000 x 64 times 64 x 1 would otherwise keep 32 threads busy).
// EPI 0: + bias, relu, store activation + mask (hidden layer)   1: + bias, store (output layer)   2: * mask of the consumer's input (input gradient)
template <int EPI>
__device__ __forceinline__ void tile_contract(const float* __restrict__ At, const float* __restrict__ B, int K, int np, int ldb, int rs, int rows4,
                                              const float* __restrict__ bias, float* __restrict__ Ct, unsigned char* __restrict__ mask,
                                              float* __restrict__ scratch) {
    const int tn = np >> 2;
    const int ntiles = (rows4 >> 2) * tn;
    int S = 1;
    while (S < 8 && ntiles * S * 2 <= SM_NT && K % (S * 8) == 0) S *= 2;
    const int kper = K / S;
    for (int base = 0; base < ntiles * S; base += SM_NT) {
        const int t = base + threadIdx.x;
        const int part = t / ntiles, tile = t - part * ntiles;
        const bool live = part < S;
        const int n0 = (tile % tn) << 2, r0 = (tile / tn) << 2;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        if (live) {
            const float* ap = At + r0 + part * kper * rs;
            const float* bp = B + n0 + part * kper * ldb;
#pragma unroll 4
            for (int k = 0; k < kper; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(ap + k * rs);
                const float4 b4 = *reinterpret_cast<const float4*>(bp + k * ldb);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
            }
        }
        if (S > 1) {   // (uniform over the block; then the loop has exactly one trip)
            if (live && part > 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) scratch[((part - 1) * 16 + j) * ntiles + tile] = acc[j >> 2][j & 3];
            }
            __syncthreads();
            if (part != 0) continue;
            for (int q = 1; q < S; ++q)
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j >> 2][j & 3] += scratch[((q - 1) * 16 + j) * ntiles + tile];
        }
        if (!live) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int n = n0 + b;
            float v[4];
            if (EPI == 2) {
                const uchar4 m = *reinterpret_cast<const uchar4*>(mask + n * rs + r0);
                v[0] = acc[0][b] * (float)m.x; v[1] = acc[1][b] * (float)m.y; v[2] = acc[2][b] * (float)m.z; v[3] = acc[3][b] * (float)m.w;
            } else {
                const float bb = bias[n];
#pragma unroll
                for (int a = 0; a < 4; ++a) v[a] = acc[a][b] + bb;
                if (EPI == 0) {
                    uchar4 m;
                    m.x = v[0] >= 0.f; m.y = v[1] >= 0.f; m.z = v[2] >= 0.f; m.w = v[3] >= 0.f;
                    *reinterpret_cast<uchar4*>(mask + n * rs + r0) = m;
                    v[0] = (float)m.x * v[0]; v[1] = (float)m.y * v[1]; v[2] = (float)m.z * v[2]; v[3] = (float)m.w * v[3];   // x.geq(0).mul(x)
                }
            }
            *reinterpret_cast<float4*>(Ct + n * rs + r0) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// gW[i][o] += sum_r At[i][r] * Gt[o][r]: 4 x 4 tiles with interleaved rows / columns (i = it + TI a, o = ot + TO b) so that the 128-bit
// loads of a warp hit distinct banks; with fewer tiles than threads the samples are split between thread groups (summed in group
// order through `scratch`).  gb[o] += sum_r Gt[o][r]: four threads per column + a fixed shuffle tree.
__device__ __forceinline__ void weight_grad(const float* __restrict__ At, const float* __restrict__ Gt, int ip, int op, int rs, int rows4,
                                            float* __restrict__ gW, float* __restrict__ gB, float* __restrict__ scratch) {
    const int ti = ip >> 2, to = op >> 2;
    const int tiles = ti * to;
    const int quads = rows4 >> 2;
    int S = 1;
    while (S < 16 && tiles * S * 2 <= SM_NT && quads >= S * 2) S *= 2;
    const int t = threadIdx.x;
    const int part = t / tiles, tile = t - part * tiles;
    const int it = tile / to, ot = tile % to;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    if (part < S) {
        const int qper = (quads + S - 1) / S;
        const int r_end = min(rows4, (part + 1) * qper * 4);
        for (int r = part * qper * 4; r < r_end; r += 4) {
            float4 av[4], gv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(At + (it + ti * a) * rs + r);
#pragma unroll
            for (int b = 0; b < 4; ++b) gv[b] = *reinterpret_cast<const float4*>(Gt + (ot + to * b) * rs + r);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float s = acc[a][b];
                    s = fmaf(av[a].x, gv[b].x, s);
                    s = fmaf(av[a].y, gv[b].y, s);
                    s = fmaf(av[a].z, gv[b].z, s);
                    s = fmaf(av[a].w, gv[b].w, s);
                    acc[a][b] = s;
                }
        }
    }
    if (S > 1) {
        if (part > 0 && part < S) {
#pragma unroll
            for (int j = 0; j < 16; ++j) scratch[((part - 1) * 16 + j) * tiles + tile] = acc[j >> 2][j & 3];
        }
        __syncthreads();
        if (part == 0)
            for (int q = 1; q < S; ++q)
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j >> 2][j & 3] += scratch[((q - 1) * 16 + j) * tiles + tile];
    }
    if (part == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) gW[(it + ti * a) * op + ot + to * b] += acc[a][b];
    }
    // bias gradient
    {
        const int col = t >> 2, sub = t & 3;
        float s = 0.f;
        if (col < op)
            for (int r = sub * 4; r < rows4; r += 16) {
                const float4 g = *reinterpret_cast<const float4*>(Gt + col * rs + r);
                s += g.x; s += g.y; s += g.z; s += g.w;
            }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (col < op && sub == 0) gB[col] += s;
    }
}

__global__ void __launch_bounds__(SM_NT, 1) mlp_small_step_kernel(const SmallMlpArgs p) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int nrank = (int)cluster.num_blocks();
    const int tid = threadIdx.x;
    const int L = p.L;
    const int rs = p.rs;
    const int row_begin = min(p.batch, rank * p.rows_per_cta);
    const int row_end = min(p.batch, row_begin + p.rows_per_cta);

    // ---- parameters -> shared memory (W, W^T, b; zero padded), partial gradients = 0.  All global loads of a batch are issued before
    //      the first dependent store: the step is latency-bound, not bandwidth-bound.
    for (int e = tid * 4; e < p.sZeroEnd; e += SM_NT * 4) *reinterpret_cast<float4*>(sm + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int l = 0; l < L; ++l) {
        const int I = p.dims[l], O = p.dims[l + 1], ip = p.dp[l], op = p.dp[l + 1], ldt = ip + 4;
        const float* Wg = p.params + p.w_off[l];
        float *W = sm + p.sW[l], *Wt = sm + p.sWt[l];
        const int n = I * O;
        if ((O & 3) == 0 && (p.w_off[l] & 3) == 0) {
            for (int base = 0; base < n; base += 16 * SM_NT) {
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = base + (j * SM_NT + tid) * 4;
                    v[j] = e < n ? *reinterpret_cast<const float4*>(Wg + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = base + (j * SM_NT + tid) * 4;
                    if (e < n) {
                        const int i = e / O, o = e - i * O;
                        *reinterpret_cast<float4*>(W + i * op + o) = v[j];
                        Wt[o * ldt + i] = v[j].x; Wt[(o + 1) * ldt + i] = v[j].y; Wt[(o + 2) * ldt + i] = v[j].z; Wt[(o + 3) * ldt + i] = v[j].w;
                    }
                }
            }
        } else {
            for (int base = 0; base < n; base += 4 * SM_NT) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = base + j * SM_NT + tid;
                    v[j] = e < n ? Wg[e] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = base + j * SM_NT + tid;
                    if (e < n) {
                        const int i = e / O, o = e - i * O;
                        W[i * op + o] = v[j];
                        Wt[o * ldt + i] = v[j];
                    }
                }
            }
        }
        if (tid < O) sm[p.sB[l] + tid] = p.params[p.b_off[l] + tid];
    }
    float loss_acc = 0.f;

    for (int rb = row_begin; rb < row_end; rb += p.rt) {
        const int R = min(p.rt, row_end - rb);
        const int rows4 = (R + 3) & ~3;
        const int O_out = p.dims[L], op_out = p.dp[L];
        // the first targets this thread needs after the forward pass: requested now, consumed in the loss stage
        float y_pre[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int e = j * SM_NT + tid, r = e / op_out, o = e - r * op_out;
            y_pre[j] = (e < op_out * rows4 && r < R && o < O_out) ? p.y[(size_t)(rb + r) * O_out + o] : 0.f;
        }
        // ---- input, feature-major; samples past R and padded features are zero
        {
            const int I = p.dims[0], ip = p.dp[0];
            float* A0 = sm + p.sA[0];
            const int n = ip * rows4;
            for (int base = 0; base < n; base += 4 * SM_NT) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = base + j * SM_NT + tid, r = e / ip, k = e - r * ip;   // consecutive threads read consecutive floats of x
                    v[j] = (e < n && r < R && k < I) ? p.x[(size_t)(rb + r) * I + k] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = base + j * SM_NT + tid, r = e / ip, k = e - r * ip;
                    if (e < n) A0[k * rs + r] = v[j];
                }
            }
        }
        __syncthreads();
        // ---- forward
        for (int l = 0; l < L; ++l) {
            const float* At = sm + p.sA[l];
            float* Ct = sm + p.sA[l + 1];
            if (l + 1 < L)
                tile_contract<0>(At, sm + p.sW[l], p.dp[l], p.dp[l + 1], p.dp[l + 1], rs, rows4, sm + p.sB[l], Ct,
                                 reinterpret_cast<unsigned char*>(sm + p.sM[l]), sm + p.sScr[l & 1]);
            else
                tile_contract<1>(At, sm + p.sW[l], p.dp[l], p.dp[l + 1], p.dp[l + 1], rs, rows4, sm + p.sB[l], Ct, nullptr, sm + p.sScr[l & 1]);
            __syncthreads();
        }
        // ---- loss and its gradient: loss = (out - y)^2 (sine_net.rs:150), d out = 2 * (out - y) * 1
        {
            const float* out = sm + p.sA[L];
            float* G = sm + p.sG[L & 1];
            const int n = op_out * rows4;
            for (int base = 0, j = 0; base < n; base += SM_NT, ++j) {
                const int e = base + tid, r = e / op_out, o = e - r * op_out;
                if (e >= n) break;
                float g = 0.f;
                if (r < R && o < O_out) {
                    const float yv = j == 0 ? y_pre[0] : j == 1 ? y_pre[1] : p.y[(size_t)(rb + r) * O_out + o];
                    const float d = out[o * rs + r] - yv;
                    loss_acc += d * d;
                    g = 2.f * d;
                }
                G[o * rs + r] = g;
            }
        }
        __syncthreads();
        // ---- backward (g_l lives in sG[l & 1]); the weight gradient of layer l and the input gradient through it run in one phase
        for (int l = L; l-- > 0;) {
            const float* Gt = sm + p.sG[(l + 1) & 1];
            weight_grad(sm + p.sA[l], Gt, p.dp[l], p.dp[l + 1], rs, rows4, sm + p.sGW[l], sm + p.sGB[l], sm + p.sScr[0]);
            if (l > 0)
                tile_contract<2>(Gt, sm + p.sWt[l], p.dp[l + 1], p.dp[l], p.dp[l] + 4, rs, rows4, nullptr, sm + p.sG[l & 1],
                                 reinterpret_cast<unsigned char*>(sm + p.sM[l - 1]), sm + p.sScr[1]);
            __syncthreads();
        }
    }

    // ---- loss: fixed-order block reduction, then rank order over the cluster
    {
        float v = loss_acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        float* red = sm + p.sRed;
        if ((tid & 31) == 0) red[1 + (tid >> 5)] = v;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < SM_NT / 32; ++w) s += red[1 + w];
            red[0] = s;
        }
    }
    cluster.sync();
    if (rank == 0 && tid == 0) {
        float s = 0.f;
        for (int q = 0; q < nrank; ++q) s += cluster.map_shared_rank(sm + p.sRed, q)[0];
        p.loss_sum[0] = s;
    }
    // ---- gradient exchange over distributed shared memory + SGD: every CTA owns a contiguous slice of each parameter segment
    for (int l = 0; l < L; ++l) {
        const int I = p.dims[l], O = p.dims[l + 1], op = p.dp[l + 1];
        const int nW = I * O, per = (nW + nrank - 1) / nrank;
        const int e_end = min(nW, (rank + 1) * per);
        for (int e = rank * per + tid; e < e_end; e += SM_NT) {
            const int so = (e / O) * op + (e % O);
            float part[SM_MAXC];
#pragma unroll
            for (int q = 0; q < SM_MAXC; ++q) part[q] = q < nrank ? cluster.map_shared_rank(sm + p.sGW[l], q)[so] : 0.f;
            float g = 0.f;
#pragma unroll
            for (int q = 0; q < SM_MAXC; ++q) g += part[q];
            p.grads[p.w_off[l] + e] = g;
            p.params[p.w_off[l] + e] = sm[p.sW[l] + so] - g * p.lr;   // *value -= *grad * self.lr
        }
        if (rank == (l % nrank))
            for (int o = tid; o < O; o += SM_NT) {
                float g = 0.f;
                for (int q = 0; q < nrank; ++q) g += cluster.map_shared_rank(sm + p.sGB[l], q)[o];
                p.grads[p.b_off[l] + o] = g;
                p.params[p.b_off[l] + o] = sm[p.sB[l] + o] - g * p.lr;
            }
    }
    cluster.sync();   // no CTA may leave while its shared memory is still being read
}

// shared-memory plan for a cluster of `nrank` CTAs; returns the bytes needed (0: the shape is outside what this kernel takes)
size_t plan(SmallMlpArgs& a, int n_layers, const size_t* dims, size_t batch, int nrank) {
    if (n_layers < 1 || n_layers > SM_MAXL || batch < 1 || batch > (1u << 20)) return 0;
    a.L = n_layers;
    a.batch = (int)batch;
    for (int l = 0; l <= n_layers; ++l) {
        if (dims[l] < 1 || dims[l] > (size_t)SM_MAXW) return 0;
        a.dims[l] = (int)dims[l];
        a.dp[l] = ((int)dims[l] + 3) & ~3;
    }
    const int per = (int)((batch + nrank - 1) / nrank);
    a.rows_per_cta = (per + 3) & ~3;
    for (int rt : {128, 64, 32}) {
        a.rt = a.rows_per_cta < rt ? a.rows_per_cta : rt;
        a.rs = a.rt + 4;
        int off = 0;
        auto take = [&](int n) { const int o = off; off += (n + 3) & ~3; return o; };
        int gmax = 0;
        for (int l = 0; l < n_layers; ++l) {
            const int n = a.dp[l] * a.dp[l + 1];
            a.sW[l] = take(n); a.sWt[l] = take(a.dp[l + 1] * (a.dp[l] + 4)); a.sGW[l] = take(n);
            a.sB[l] = take(a.dp[l + 1]); a.sGB[l] = take(a.dp[l + 1]);
            gmax = a.dp[l + 1] > gmax ? a.dp[l + 1] : gmax;
        }
        a.sZeroEnd = off;
        for (int l = 0; l < n_layers; ++l) a.sM[l] = take((a.dp[l + 1] * a.rs + 3) / 4);
        for (int l = 0; l <= n_layers; ++l) a.sA[l] = take(a.dp[l] * a.rs);
        a.sG[0] = take(gmax * a.rs);
        a.sG[1] = take(gmax * a.rs);
        a.sScr[0] = take(SM_SCRATCH);
        a.sScr[1] = take(SM_SCRATCH);
        a.sRed = take(4 + SM_NT / 32);
        if ((size_t)off * 4 <= 227u * 1024) return (size_t)off * 4;
    }
    return 0;
}

// 16 CTAs (a non-portable cluster size: one whole GPC) when there are enough samples to feed them and the device can place such a
// cluster with this much shared memory; else the portable 8
int pick_cluster(sl_ctx* ctx, int n_layers, const size_t* dims, size_t batch) {
    int& can16 = ctx->mlp_small_can16;
    if (batch < 256) return 8;
    SmallMlpArgs a{};
    const size_t bytes = plan(a, n_layers, dims, batch, 16);
    if (!bytes) return 8;
    if (can16 < 0) {
        can16 = 0;
        if (cudaFuncSetAttribute(mlp_small_step_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(16);
            cfg.blockDim = dim3(SM_NT);
            cfg.dynamicSmemBytes = 227 * 1024;
            cudaLaunchAttribute at{};
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = 16; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, mlp_small_step_kernel, &cfg) == cudaSuccess && n >= 1) can16 = 1;
        }
        cudaGetLastError();
    }
    return can16 == 1 ? 16 : 8;
}

}  // namespace

extern "C" int sl_mlp_small_fits(sl_ctx* ctx, int n_layers, const size_t* dims, size_t batch) {
    if (!ctx || !dims) return 0;
    SmallMlpArgs a{};
    return plan(a, n_layers, dims, batch, 8) != 0;
}

extern "C" int sl_mlp_small_step(sl_ctx* ctx, int dtype, int n_layers, const size_t* dims, const size_t* seg_off, size_t batch, const void* x,
                                 const void* y, void* params, void* grads, double lr, void* loss_sum_dev) {
    if (!ctx) return SL_ERR_INVALID_ARG;
    SL_REQUIRE(ctx, dtype == SL_F32, "f32 only");
    SL_REQUIRE(ctx, dims && seg_off && x && y && params && grads && loss_sum_dev, "null argument");
    SL_REQUIRE(ctx, sl_aligned16(params) && sl_aligned16(grads), "params / grads must be 16-byte aligned");
    if (!ctx->mlp_small_attr) {
        SL_CUDA(ctx, cudaFuncSetAttribute(mlp_small_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ctx->mlp_small_attr = true;
    }
    SmallMlpArgs a{};
    if (!plan(a, n_layers, dims, batch, 8))
        return sl_set_error(ctx, SL_ERR_UNSUPPORTED, "sl_mlp_small_step: needs 1..%d layers of width <= %d that fit in shared memory", SM_MAXL, SM_MAXW);
    const int nrank = pick_cluster(ctx, n_layers, dims, batch);
    const size_t bytes = plan(a, n_layers, dims, batch, nrank);
    for (int l = 0; l < n_layers; ++l) {
        a.w_off[l] = (int)seg_off[2 * l];
        a.b_off[l] = (int)seg_off[2 * l + 1];
    }
    a.lr = (float)lr;
    a.x = (const float*)x;
    a.y = (const float*)y;
    a.params = (float*)params;
    a.grads = (float*)grads;
    a.loss_sum = (float*)loss_sum_dev;
    sl_note_writes(ctx, params, grads, loss_sum_dev);

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nrank);
    cfg.blockDim = dim3(SM_NT);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = nrank; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    sl_ctx::ProfRec pr{};
    const bool prof = ctx->profiling && ctx->profiling_all;
    if (prof) {
        cudaEventCreate(&pr.a);
        cudaEventCreate(&pr.b);
        pr.name = "mlp_small_step_kernel";
        cudaEventRecord(pr.a, ctx->stream);
    }
    SL_CUDA(ctx, cudaLaunchKernelEx(&cfg, mlp_small_step_kernel, a));
    if (prof) {
        cudaEventRecord(pr.b, ctx->stream);
        ctx->prof.push_back(pr);
    }
    ctx->launches++;
    return SL_OK;
}
