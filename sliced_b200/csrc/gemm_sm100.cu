// gemm_sm100.cu — family G: the dense contractions (gemm, gemmT, Tgemm and therefore gemm_grad) on the 5th-generation
// tensor cores of B200: TMA -> 128B-swizzled shared memory -> tcgen05.mma kind::tf32 -> fp32 accumulators in TMEM ->
// tcgen05.ld epilogue.  Hand-written PTX; no CUTLASS / cuBLAS on this path.
//
// Precision modes (sl_gemm_mode):
//   SL_GEMM_3XTF32 (default)  every fp32 operand x is split into  hi = rna_tf32(x),  lo = rna_tf32(x - hi)  and each
//                             k-slice issues three MMAs into the same TMEM accumulator:  lo*hi + hi*lo + hi*hi.
//                             The dropped lo*lo term and the rounding of lo are both ~2^-24 relative, i.e. fp32-class.
//   SL_GEMM_TF32              one MMA per k-slice on the raw fp32 bits (the tensor core reads the top 19 bits).
//
// tcgen05.mma takes both operands from shared memory through descriptors, so there is no register stage in which to
// split an operand.  The split is therefore done by a bandwidth-bound "prep" pass that writes K-major hi / lo planes
// into the context's scratch; the same pass transposes an operand whose contraction index is not contiguous
// (gemm's rhs, Tgemm's lhs and rhs), so that the MMA kernel only ever sees K-major A[M x K] and B[N x K] tiles
// (the canonical, swizzle-128B K-major UMMA layout).  Extra HBM traffic: 12 B per operand element once per gemm,
// O((MK + KN) / (MNK)) of the MMA work — <2 % at 4096^3 and above.
//
// Kernel anatomy (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: per k-block, one cp.async.bulk.tensor per operand plane into the stage ring
//   warp 1      MMA issuer (one lane): tcgen05.mma x (BLOCK_K/8) x TERMS per k-block, tcgen05.commit frees the stage;
//               owns the TMEM allocation (2 accumulator buffers so the epilogue of tile i overlaps the mainloop of i+1)
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns at a time, optional += C / + bias / relu, 128-bit stores
//   barriers    full[STAGES], empty[STAGES] (TMA <-> MMA), tmem_full[2], tmem_empty[2] (MMA <-> epilogue)
#include <cuda.h>

#include <cuda_fp16.h>

#include "common.cuh"

int sl_gemm_simt(sl_ctx* ctx, int dtype, int trans_a, int trans_b, size_t m, size_t n, size_t k, const void* a, const void* b, void* c,
                 int accumulate);
int sl_gemm_skinny_f32(sl_ctx* ctx, int trans_a, int trans_b, size_t m, size_t n, size_t k, const float* a, const float* b, float* c,
                       int accumulate, const float* mask_src, int* mask_done, const uint32_t* mask_bits = nullptr);

namespace {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a protocol bug must abort the kernel (sticky error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > 4000000000ull) {
            printf("sliced_b200 gemm: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread on behalf of the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t <-> TMEM lane base+t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, K-major operand tile [rows x BLOCK_K] fp32 written by TMA with SWIZZLE_(BLOCK_K*4)B:
// row r sits at r * (BLOCK_K*4) bytes, 16-byte chunks XOR-swizzled within each group of 8 rows.
//   bits [0,14)  start address >> 4         bits [16,30) leading byte offset >> 4 (unused for swizzled K-major; 1)
//   bits [32,46) stride byte offset >> 4 = distance between 8-row groups        bits [46,48) descriptor version = 1
//   bits [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
template <int BLOCK_K>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    constexpr uint64_t row_bytes = BLOCK_K * 4;
    constexpr uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}
// MN-major operand tile [BLOCK_K x 128 mn] fp32 (the contraction index is the SLOW one in memory: gemm's rhs, Tgemm's operands):
// four 32-mn chunks, each written by its own TMA box {32 mn, BLOCK_K k} as BLOCK_K rows of 128 bytes.  For 32-bit MN-major
// operands the tensor core only accepts the "128B swizzle with 32-byte atoms" layout (UMMA layout type 1, TMA
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): within every group of 4 k-rows the 32-byte chunks of a row are XOR-ed with the row
// index (address bits [5,7) ^= bits [7,9)).  Canonical form ((T,8,m),(4,k)) : ((1,T,LBO),(8T,SBO)), T = 4 tf32 per 16 B:
// LBO = distance between 32-mn chunks, SBO = distance between groups of 4 k-rows (512 B).  One MMA (K = 8) consumes two
// such groups (1024 B) of every chunk.
template <int BLOCK_K>
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((BLOCK_K * 128) >> 4) << 16;   // LBO: one chunk = BLOCK_K rows x 128 B
    d |= (uint64_t)(512 >> 4) << 32;               // SBO: 4 k-rows x 128 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                        // SWIZZLE_128B_BASE32B
    return d;
}
// 16-bit MN-major operand tile [64 k x 128 mn] fp16: two 64-mn chunks, each one TMA box {64 mn, 64 k} = 64 rows of 128 bytes with
// the standard 128-byte swizzle (UMMA layout type 2).  Canonical form ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO)), T = 8 halves per
// 16 B: LBO = distance between 64-mn chunks (8192 B), SBO = distance between groups of 8 k-rows (1024 B).  One MMA (K = 16)
// consumes two such groups (2048 B) of every chunk.
__device__ __forceinline__ uint64_t make_smem_desc_mn16(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(8192 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor: D = F32 (bits[4,6)=1), A = B = TF32 (bits[7,10)=2, [10,13)=2), A / B major (bit 15 / 16: 0 = K-major,
// 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
// same with A = B = F16 (format code 0), for kind::f16
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ the MMA kernel
//
// Accumulation scheme ("promotion").  The tensor core adds each MMA's 8 products into the fp32 TMEM accumulator with
// round-toward-zero, so a long K chain drifts linearly in K (measured: 13x the error of a CPU sgemm at K = 4096).
// The kernel therefore never lets a TMEM chain run longer than `kc_blocks` k-blocks: the MMA warp ping-pongs between
// two TMEM buffers, one K-chunk each, and the epilogue warps fold every finished chunk into per-thread fp32 REGISTER
// accumulators with round-to-nearest adds (sqrt(K) growth, like a blocked CPU summation).  Folding a chunk (a
// tcgen05.ld of the buffer + 128 FADDs per thread) overlaps the next chunk's MMAs, so it is free.
// In TF32 fast mode the operand rounding dominates and kc_blocks is the whole K (single chunk).
constexpr int BLOCK_M = 128;
constexpr int UMMA_K = 8;  // 32 bytes of tf32 per MMA
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int GROUP_M = 8;

struct GemmParams {
    int M, N, K;
    float* C;           // [M x N] row-major
    const float* bias;  // [N] or NULL
    int accumulate;     // C += acc
    int relu;           // C = (v >= 0) * v after bias
    int c_vec_ok;       // N % 4 == 0 and C 16-byte aligned
    int kc_blocks;      // k-blocks per TMEM accumulation chunk
    int group;              // tile raster: blocks per group (GROUP_M when 0)
    int group_along_n;      // group over N (sweep M inside a band of `group` n-blocks) instead of over M
    int splits;             // split-K factor (2-CTA kernel): partial s is written to C + s*M*N, folded afterwards by the host
    float* C2;              // optional second output: C2 = (v >= 0) * v of the value stored to C (fused Matrix::relu)
    const float* mask_src;  // optional [M x N]: v *= (mask_src >= 0) before the store (fused relu gradient)
    const float* row_scale; // 3xFP16 mode: acc is multiplied by row_scale[row] * col_scale[col] (exact powers of two undoing the
    const float* col_scale; //   operand scaling) before anything else; both NULL otherwise
    int kc_first;           // 2-CTA kernel: k-blocks of the FIRST TWO chunks of every tile (>= kc_blocks; see the kernel)
    int debug;              // bit 0: skip the final store (timing experiments only, SLICED_GEMM_DEBUG)
    unsigned long long hint_a, hint_b;   // L2 eviction-priority policy of the A / B operand loads (0 = none)
    int dynamic;            // 2-CTA kernel: 1 = one cluster per work unit, units handed out by cluster launch control (work stealing)
    // relu mask carried as one bit per element ([M x bits_words] uint32, bit c%32 of word c/32; N % 32 == 0, 2-CTA kernel only):
    uint32_t* bits_out;          //   written: (v >= 0) of the value after bias, before the in-place relu
    const uint32_t* bits_in;     //   read: v *= bit (the relu gradient) — replaces mask_src (4 bytes per element) by 1/32 of the traffic
    int bits_words;
    unsigned long long* stats;   // SLICED_GEMM_STATS (diagnosis): per leader CTA [total, wait full, wait tmem_empty, store phase, units] clocks
};

// shared epilogue arithmetic of both MMA kernels: 4 consecutive columns of one output row
__device__ __forceinline__ void epilogue_store4(const GemmParams& p, float* crow, size_t row_off, int n0, float4 v, float rs = 1.f) {
    if (p.col_scale) {
        const float4 cs = __ldg(reinterpret_cast<const float4*>(p.col_scale + n0));
        v.x = (v.x * rs) * cs.x; v.y = (v.y * rs) * cs.y; v.z = (v.z * rs) * cs.z; v.w = (v.w * rs) * cs.w;
    } else if (p.row_scale) {
        v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
    }
    if (p.accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(crow + n0);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    if (p.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (p.mask_src) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(p.mask_src + row_off + n0));
        v.x = (m.x >= 0.f ? 1.f : 0.f) * v.x; v.y = (m.y >= 0.f ? 1.f : 0.f) * v.y;
        v.z = (m.z >= 0.f ? 1.f : 0.f) * v.z; v.w = (m.w >= 0.f ? 1.f : 0.f) * v.w;
    }
    if (p.relu) {
        v.x = (v.x >= 0.f ? 1.f : 0.f) * v.x; v.y = (v.y >= 0.f ? 1.f : 0.f) * v.y;
        v.z = (v.z >= 0.f ? 1.f : 0.f) * v.z; v.w = (v.w >= 0.f ? 1.f : 0.f) * v.w;
    }
    *reinterpret_cast<float4*>(crow + n0) = v;
    if (p.C2) {
        float4 r;
        r.x = (v.x >= 0.f ? 1.f : 0.f) * v.x; r.y = (v.y >= 0.f ? 1.f : 0.f) * v.y;
        r.z = (v.z >= 0.f ? 1.f : 0.f) * v.z; r.w = (v.w >= 0.f ? 1.f : 0.f) * v.w;
        *reinterpret_cast<float4*>(p.C2 + row_off + n0) = r;
    }
}
__device__ __forceinline__ void epilogue_store1(const GemmParams& p, float* crow, size_t row_off, int n, float v, float rs = 1.f) {
    if (p.col_scale) v = (v * rs) * __ldg(p.col_scale + n);
    else if (p.row_scale) v *= rs;
    if (p.accumulate) v += crow[n];
    if (p.bias) v += __ldg(p.bias + n);
    if (p.mask_src) v = (__ldg(p.mask_src + row_off + n) >= 0.f ? 1.f : 0.f) * v;
    if (p.relu) v = (v >= 0.f ? 1.f : 0.f) * v;
    crow[n] = v;
    if (p.C2) p.C2[row_off + n] = (v >= 0.f ? 1.f : 0.f) * v;
}

template <int BLOCK_N, int BLOCK_K, int TERMS, int STAGES>
struct GemmCfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 4;
    static constexpr int PLANES = TERMS == 3 ? 2 : 1;
    static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
    static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two chunk buffers; 256 or 512 (power of two)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int COLS_PER_WARP = BLOCK_N / (EPI_WARPS / 4);  // each lane quadrant is shared by EPI_WARPS/4 warps
    static_assert(SMEM_BYTES <= 227 * 1024, "stage ring does not fit in shared memory");
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
};

template <int BLOCK_N, int BLOCK_K, int TERMS, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, BLOCK_K, TERMS, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: required by the 128B swizzle atom (8 rows x 128 B)
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_b_hi);
        if (TERMS == 3) {
            tma_prefetch_desc(&map_a_lo);
            tma_prefetch_desc(&map_b_lo);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full_bar(s), 1);
            mbar_init(tmem_empty_bar(s), EPI_WARPS);  // one arrival per epilogue warp
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = num_m * num_n;
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int kc = p.kc_blocks > 0 ? p.kc_blocks : num_kb;
    const int num_chunks = (num_kb + kc - 1) / kc;
    auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
        // grouped raster: `grp` blocks of the grouped dimension stay L2-resident while the other dimension is swept
        if (p.group_along_n) {
            const int group_size = p.group * num_m;
            const int group = tile / group_size;
            const int first_n = group * p.group;
            const int gsz = min(num_n - first_n, p.group);
            const int in_group = tile - group * group_size;
            n_blk = first_n + in_group % gsz;
            m_blk = in_group / gsz;
        } else {
            const int group_size = p.group * num_n;
            const int group = tile / group_size;
            const int first_m = group * p.group;
            const int gsz = min(num_m - first_m, p.group);
            const int in_group = tile - group * group_size;
            m_blk = first_m + in_group % gsz;
            n_blk = in_group / gsz;
        }
    };

    if (warp == 0) {
        // ================================================================= TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_blk, n_blk;
                tile_coords(tile, m_blk, n_blk);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::PLANES * Cfg::A_BYTES;
                    mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                    tma_load_2d(sa, &map_a_hi, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
                    tma_load_2d(sb, &map_b_hi, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
                    if (TERMS == 3) {
                        tma_load_2d(sa + Cfg::A_BYTES, &map_a_lo, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
                        tma_load_2d(sb + Cfg::B_BYTES, &map_b_lo, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t g = 0;  // global chunk counter: TMEM buffer = g & 1
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int ch = 0; ch < num_chunks; ++ch, ++g) {
                    const uint32_t buf = g & 1;
                    mbar_wait(tmem_empty_bar(buf), ((g >> 1) & 1) ^ 1);  // epilogue has folded this buffer's previous chunk
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * BLOCK_N;
                    const int kb_end = min(num_kb, (ch + 1) * kc);
                    for (int kb = ch * kc; kb < kb_end; ++kb) {
                        mbar_wait(full_bar(stage), phase);  // TMA bytes have landed
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        const uint32_t sb = sa + Cfg::PLANES * Cfg::A_BYTES;
                        const uint64_t da_hi = make_smem_desc<BLOCK_K>(sa);
                        const uint64_t db_hi = make_smem_desc<BLOCK_K>(sb);
                        const uint64_t da_lo = make_smem_desc<BLOCK_K>(sa + Cfg::A_BYTES);
                        const uint64_t db_lo = make_smem_desc<BLOCK_K>(sb + Cfg::B_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);  // advance the start address inside the swizzle atom
                            const uint32_t first = (kb == ch * kc && k == 0) ? 0u : 1u;
                            if (TERMS == 3) {
                                umma_tf32(tmem_d, da_lo + koff, db_hi + koff, idesc, first);
                                umma_tf32(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
                                umma_tf32(tmem_d, da_hi + koff, db_hi + koff, idesc, 1u);
                            } else {
                                umma_tf32(tmem_d, da_hi + koff, db_hi + koff, idesc, first);
                            }
                        }
                        umma_commit(empty_bar(stage));  // stage is free once these MMAs have read it
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit(tmem_full_bar(buf));  // chunk complete -> epilogue
                }
            }
        }
        __syncwarp();
    } else {
        // ================================================================= epilogue (warps 2..9)
        constexpr int CPW = Cfg::COLS_PER_WARP;
        const int ew = warp - 2;
        const int quad = warp & 3;            // TMEM lanes [32*quad, 32*quad+32) are the only ones this warp may read
        const int col0 = (ew >> 2) * CPW;     // this warp's column slice of the tile
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(tile, m_blk, n_blk);
            float acc[CPW];
            for (int ch = 0; ch < num_chunks; ++ch, ++g) {
                const uint32_t buf = g & 1;
                mbar_wait(tmem_full_bar(buf), (g >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < CPW / 32; ++c) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BLOCK_N + col0 + c * 32);
                    tmem_ld_32x32(taddr, r);
                    tmem_ld_wait();
                    if (ch == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c * 32 + j] = __uint_as_float(r[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);  // round-to-nearest promotion
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar(buf));
            }
            // ---- write the tile slice: optional += C, + bias, relu
            const int row = m_blk * BLOCK_M + quad * 32 + lane;
            if (row < p.M) {
                float* crow = p.C + (size_t)row * p.N;
                const int nbase = n_blk * BLOCK_N + col0;
#pragma unroll
                for (int j = 0; j < CPW; j += 4) {
                    const int n0 = nbase + j;
                    if (n0 >= p.N) continue;
                    if (p.c_vec_ok && n0 + 4 <= p.N) {
                        epilogue_store4(p, crow, (size_t)row * p.N, n0, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n0 + e < p.N) epilogue_store1(p, crow, (size_t)row * p.N, n0 + e, acc[j + e]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ the 2-CTA MMA kernel
//
// cta_group::2: two CTAs on the SMs of one TPC (a cluster of 2) compute one 256 x 256 tile.  Each CTA stages ITS 128 rows
// of A and ITS 128 rows (N) of B; the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256),
// which reads both CTAs' shared memory and writes each CTA's 128 accumulator rows into that CTA's own TMEM.
// Per CTA and k-block this moves 64 KB (3xTF32) from L2 instead of the 96 KB of the 128 x 256 single-CTA tile: the
// single-CTA kernel is bound by L2 -> SM traffic (ncu: 12.3 TB/s of xbar reads at 77 % tensor-pipe activity).
//   full[s]        lives in the leader; both CTAs' TMA loads complete_tx on it (cp.async.bulk.tensor .cta_group::2)
//   empty[s]       one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tmem_full[b]   one per CTA (multicast commit);   tmem_empty[b] in the leader, 2 x 8 epilogue-warp arrivals (remote via mapa)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
// Arrival on a barrier of the pair's leader.  Default semantics (release at CTA scope): what the arrival has to order is this warp's
// tcgen05.ld of the accumulator chunk (tcgen05.fence::before_thread_sync does that) — a cluster-scope release additionally made the
// warp wait for every global store of the previous tile's store phase to be acknowledged (ncu: 13 % of all samples in ERRBAR).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {   // SLICED_GEMM_DEBUG bit 1: the old form, for A/B timing
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose mbarrier may live in the peer CTA of the pair (the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_leader, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_leader), "r"(c0), "r"(c1)
        : "memory");
}
// same with an L2 eviction-priority hint (createpolicy encodings: evict_first for a streamed operand, evict_last for a re-read one)
__device__ __forceinline__ void tma_load_2d_2sm_hint(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_leader, int32_t c0, int32_t c1,
                                                     uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_leader), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 1) umma_f16_2sm(tmem_d, desc_a, desc_b, idesc, accumulate);
    else umma_tf32_2sm(tmem_d, desc_a, desc_b, idesc, accumulate);
}
// arrive (once the issued MMAs retire) on the barrier at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}

// ---- dynamic work distribution: cluster launch control (Blackwell's hardware work-stealing).  The grid has one cluster per work
// unit; a running cluster asks the launcher to CANCEL a not-yet-launched cluster and takes over its unit.  The 16-byte response is
// multicast to the same shared-memory offset of both CTAs of the pair and completes a transaction barrier in each.
__device__ __forceinline__ void clc_try_cancel_2sm(uint32_t resp_addr, uint32_t bar) {
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
                 ::"r"(resp_addr), "r"(bar) : "memory");
}
// blockIdx.x of the first CTA of the cancelled cluster, or -1 when nothing was left to cancel
__device__ __forceinline__ int clc_response_ctaid_x(uint32_t resp_addr) {
    uint32_t x, valid;
    asm volatile(
        "{\n\t.reg .pred p1;\n\t.reg .b128 resp;\n\t"
        "ld.shared.b128 resp, [%2];\n\t"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, resp;\n\t"
        "selp.u32 %1, 1, 0, p1;\n\t"
        "mov.u32 %0, 0;\n\t"
        "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, resp;\n\t}"
        : "=r"(x), "=r"(valid)
        : "r"(resp_addr)
        : "memory");
    return valid ? (int)x : -1;
}

// ---- coalesced epilogue.  tcgen05.ld hands every thread one accumulator ROW, so storing straight from registers makes each
// warp-wide STG touch 32 different rows = 32 cache lines (measured: the store phase, not the MMAs, bounded the K = 4096 gemms at
// 75 % tensor-pipe activity).  Instead each warp bounces 32 x 32 blocks through a private 4 KB shared-memory tile (16-byte chunks
// XOR-swizzled by row so both the row-owner and the coalesced access pattern are conflict-free): global accesses then cover
// 4 rows x 128 contiguous bytes per instruction.
__device__ __forceinline__ int epi_slot(int r, int ch) { return r * 32 + ((ch ^ (r & 7)) << 2); }
// registers (lane = row, v[32] = 32 consecutive columns)  ->  dst[(row_base + r) * ld + col0 + c], rows < rows_valid, cols < cols_valid (multiple of 4)
__device__ __forceinline__ void epi_scatter(float* stage, int lane, const float (&v)[32], float* dst, size_t ld, int rows_valid, int cols_valid) {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<float4*>(stage + epi_slot(lane, ch)) = make_float4(v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
    __syncwarp();
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        const float4 x = *reinterpret_cast<const float4*>(stage + epi_slot(r, ch));
        if (r < rows_valid && ch * 4 < cols_valid) *reinterpret_cast<float4*>(dst + (size_t)r * ld + ch * 4) = x;
    }
    __syncwarp();
}
// the reverse: src tile -> the row-owning lane, combined into v[32] on the fly (MASK: v *= (src >= 0); else v += src); out-of-range
// elements read as 0.  The read-back applies chunk by chunk so no second 32-register array is live (the kernel is capped at 168).
template <bool MASK>
__device__ __forceinline__ void epi_gather(float* stage, int lane, float (&v)[32], const float* src, size_t ld, int rows_valid, int cols_valid) {
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_valid && ch * 4 < cols_valid) x = *reinterpret_cast<const float4*>(src + (size_t)r * ld + ch * 4);
        *reinterpret_cast<float4*>(stage + epi_slot(r, ch)) = x;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 x = *reinterpret_cast<const float4*>(stage + epi_slot(lane, c));
        if (MASK) {
            v[c * 4] = (x.x >= 0.f ? 1.f : 0.f) * v[c * 4]; v[c * 4 + 1] = (x.y >= 0.f ? 1.f : 0.f) * v[c * 4 + 1];
            v[c * 4 + 2] = (x.z >= 0.f ? 1.f : 0.f) * v[c * 4 + 2]; v[c * 4 + 3] = (x.w >= 0.f ? 1.f : 0.f) * v[c * 4 + 3];
        } else {
            v[c * 4] += x.x; v[c * 4 + 1] += x.y; v[c * 4 + 2] += x.z; v[c * 4 + 3] += x.w;
        }
    }
    __syncwarp();
}

// KIND 0: tf32 planes (4-byte elements, 32 k per 128-byte row, MMA K = 8); KIND 1: fp16 planes (64 k per row, MMA K = 16)
template <int TERMS, int STAGES, int KIND = 0>
struct Gemm2Cfg {
    static constexpr int ELEM = KIND == 1 ? 2 : 4;
    static constexpr int BLOCK_K = 128 / ELEM;
    static constexpr int MMA_K = 32 / ELEM;
    static constexpr int TILE_M = 256, TILE_N = 256;       // per CTA pair
    static constexpr int A_BYTES = 128 * 128;              // this CTA's 128 rows of A, 128 bytes of k each
    static constexpr int B_BYTES = 128 * 128;              // this CTA's 128 rows (N) of B
    static constexpr int PLANES = TERMS == 3 ? 2 : 1;
    static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
    static constexpr int TMEM_COLS = 2 * TILE_N;           // two chunk buffers of 256 fp32 columns
    static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 4096;   // one 32 x 32 fp32 transposition tile per epilogue warp
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_STAGE_BYTES;
    static constexpr int COLS_PER_WARP = TILE_N / (EPI_WARPS / 4);
    static_assert(SMEM_BYTES <= 227 * 1024, "stage ring does not fit in shared memory");
};

template <int TERMS, int STAGES, bool A_MN, bool B_MN, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)   // 10 warps -> 3 on one SM sub-partition -> 168 registers
gemm_tf32_2cta_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const GemmParams p) {
    using Cfg = Gemm2Cfg<TERMS, STAGES, KIND>;
    constexpr int BLOCK_K = Cfg::BLOCK_K;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    // dynamic scheduling (p.dynamic): a ring of SCHED_SLOTS launch-control responses, one full barrier per CTA (16 response bytes),
    // one empty barrier in the leader (every consumer of both CTAs: 2 producers, 1 MMA thread, 2 x 8 epilogue warps)
    constexpr int SCHED_SLOTS = 2;
    constexpr uint32_t SCHED_CONSUMERS = 2 * (1 + EPI_WARPS) + 1;
    auto sched_full_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + 6 + s); };
    auto sched_empty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + 6 + SCHED_SLOTS + s); };
    auto sched_resp = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + 6 + 2 * SCHED_SLOTS) + 16u * s; };
    static_assert(8 * (2 * STAGES + 6 + 2 * SCHED_SLOTS) + 16 * SCHED_SLOTS <= 256, "barrier area is 256 bytes");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // MN-major chunking: tf32 -> four 32-mn chunks of BLOCK_K rows x 128 B; fp16 -> two 64-mn chunks
    constexpr int MN_CHUNKS = KIND == 1 ? 2 : 4;
    constexpr int MN_CHUNK_ROWS = 128 / MN_CHUNKS;
    constexpr int MN_CHUNK_BYTES = BLOCK_K * 128;
    constexpr int MN_KSTEP_BYTES = Cfg::MMA_K * 128;
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_b_hi);
        if (TERMS == 3) {
            tma_prefetch_desc(&map_a_lo);
            tma_prefetch_desc(&map_b_lo);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);     // leader's producer arrives (+ both CTAs' transaction bytes); unused in the peer
            mbar_init(empty_bar(s), 1);    // one multicast commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full_bar(s), 1);
            mbar_init(tmem_empty_bar(s), 2 * EPI_WARPS);  // both CTAs' epilogue warps; only the leader's copy is used
        }
        for (int s = 0; s < SCHED_SLOTS; ++s) {
            mbar_init(sched_full_bar(s), 1);                 // this CTA's producer arms it (+ 16 response bytes)
            mbar_init(sched_empty_bar(s), SCHED_CONSUMERS);  // only the leader's copy is used
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // barriers of both CTAs are initialised before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int num_m = (p.M + Cfg::TILE_M - 1) / Cfg::TILE_M;
    const int num_n = (p.N + Cfg::TILE_N - 1) / Cfg::TILE_N;
    const int num_tiles = num_m * num_n;
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int kc = p.kc_blocks > 0 ? p.kc_blocks : num_kb;
    // Chunk schedule of one work unit.  While the epilogue warps store a finished tile they cannot drain TMEM, so the MMA warp can
    // only run two chunks into the next tile before it stalls: the first two chunks of every unit are therefore longer (kc_first
    // k-blocks) than the rest (kc), buying the store phase 2 * kc_first k-blocks of MMA time at the cost of two slightly longer
    // round-toward-zero chains per tile.
    const int kcf = p.kc_first > kc ? p.kc_first : kc;
    auto chunk_begin = [&](int ch) { return ch < 2 ? ch * kcf : 2 * kcf + (ch - 2) * kc; };
    auto chunk_count = [&](int nkb) { return nkb <= 2 * kcf ? (nkb + kcf - 1) / kcf : 2 + (nkb - 2 * kcf + kc - 1) / kc; };
    // split-K: a work unit is (tile, split); split s covers k-blocks [s*kbs, (s+1)*kbs) and writes its partial tile to
    // C + s*M*N (the host points C at scratch and folds the partials in order afterwards)
    // Units are numbered split-major (unit = split * num_tiles + tile): the ~74 units running at any time then belong to ONE k range
    // and share operand slices through L2 (tile-major numbering ran both halves of 37 tiles side by side, which share nothing:
    // 11 GB of DRAM reads per weight-gradient launch against 2.2 GB algorithmic)
    const int splits = p.splits > 0 ? p.splits : 1;
    const int kbs = (num_kb + splits - 1) / splits;
    const int num_units = num_tiles * splits;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    // Unit sequence of this cluster.  Static: cluster_id, + num_clusters, ... (persistent grid of one cluster per SM pair).
    // Dynamic: the grid holds one cluster per unit; a cluster starts with its own unit and then takes over the units of clusters
    // that have not been launched yet (clusterlaunchcontrol.try_cancel) until none is left.  SMs held by another kernel (the NCCL
    // exchange of a data-parallel step) then just mean fewer clusters sharing the same queue, instead of whole static shares of
    // the tiles waiting for those SMs.  Fetch number i (issued by the leader's producer when it STARTS its i-th unit, so the
    // round trip hides behind that unit) names the (i+1)-th unit; every role of both CTAs reads it when it finishes its i-th.
    const bool dyn = p.dynamic != 0;
    auto sched_prefetch = [&](uint32_t it) {   // producer threads of both CTAs
        const uint32_t slot = it % SCHED_SLOTS;
        if (leader) mbar_wait(sched_empty_bar(slot), ((it / SCHED_SLOTS) & 1) ^ 1);   // everybody has read fetch it - SCHED_SLOTS
        mbar_expect_tx(sched_full_bar(slot), 16);
        if (leader) clc_try_cancel_2sm(sched_resp(slot), sched_full_bar(slot));
    };
    auto next_unit = [&](int unit, uint32_t& it, bool whole_warp) -> int {
        if (!dyn) {
            const int n = unit + num_clusters;
            return n < num_units ? n : -1;
        }
        const uint32_t slot = it % SCHED_SLOTS, par = (it / SCHED_SLOTS) & 1;
        ++it;
        mbar_wait(sched_full_bar(slot), par);
        const int x = clc_response_ctaid_x(sched_resp(slot));
        fence_proxy_async();   // this (generic-proxy) read is ordered before the next (async-proxy) response written to the slot
        if (whole_warp) __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(sched_empty_bar(slot), 0));
        return x < 0 ? -1 : (x >> 1);
    };
    auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
        // grouped raster: `grp` blocks of the grouped dimension stay L2-resident while the other dimension is swept
        if (p.group_along_n) {
            const int group_size = p.group * num_m;
            const int group = tile / group_size;
            const int first_n = group * p.group;
            const int gsz = min(num_n - first_n, p.group);
            const int in_group = tile - group * group_size;
            n_blk = first_n + in_group % gsz;
            m_blk = in_group / gsz;
        } else {
            const int group_size = p.group * num_n;
            const int group = tile / group_size;
            const int first_m = group * p.group;
            const int gsz = min(num_m - first_m, p.group);
            const int in_group = tile - group * group_size;
            m_blk = first_m + in_group % gsz;
            n_blk = in_group / gsz;
        }
    };

    if (warp == 0) {
        // ================================================================= TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int unit = cluster_id; unit >= 0; unit = next_unit(unit, it, false)) {
                if (dyn) sched_prefetch(it);
                int m_blk, n_blk;
                tile_coords(unit % num_tiles, m_blk, n_blk);
                const int kb0 = (unit / num_tiles) * kbs, kb1 = min(num_kb, kb0 + kbs);
                const int row_a = m_blk * Cfg::TILE_M + (int)rank * 128;
                const int row_b = n_blk * Cfg::TILE_N + (int)rank * 128;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::PLANES * Cfg::A_BYTES;
                    const uint32_t fb = mapa_shared(full_bar(stage), 0);  // the leader's barrier, as a shared::cluster address
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
                    // K-major plane: one box {BLOCK_K k, 128 rows}; MN-major plane: one box {32 | 64 mn, BLOCK_K k} per chunk
                    auto load_box = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t hint) {
                        if (hint) tma_load_2d_2sm_hint(dst, map, fb, c0, c1, hint);
                        else tma_load_2d_2sm(dst, map, fb, c0, c1);
                    };
                    auto load_plane = [&](uint32_t dst, const CUtensorMap* map, int row0, bool mn, uint64_t hint) {
                        if (!mn) {
                            load_box(dst, map, kb * BLOCK_K, row0, hint);
                        } else {
#pragma unroll
                            for (int g = 0; g < MN_CHUNKS; ++g) load_box(dst + g * MN_CHUNK_BYTES, map, row0 + g * MN_CHUNK_ROWS, kb * BLOCK_K, hint);
                        }
                    };
                    load_plane(sa, &map_a_hi, row_a, A_MN, p.hint_a);
                    load_plane(sb, &map_b_hi, row_b, B_MN, p.hint_b);
                    if (TERMS == 3) {
                        load_plane(sa + Cfg::A_BYTES, &map_a_lo, row_a, A_MN, p.hint_a);
                        load_plane(sb + Cfg::B_BYTES, &map_b_lo, row_b, B_MN, p.hint_b);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================= MMA issuer (leader CTA only)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = KIND == 1 ? make_idesc_f16(Cfg::TILE_M, Cfg::TILE_N, A_MN ? 1 : 0, B_MN ? 1 : 0)
                                                 : make_idesc_tf32(Cfg::TILE_M, Cfg::TILE_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
            // K-major descriptors are byte-based (128-byte rows, 8-row groups) and identical for both element sizes
            auto desc_k = [](uint32_t addr) { return make_smem_desc<32>(addr); };
            auto desc_mn = [](uint32_t addr) { return KIND == 1 ? make_smem_desc_mn16(addr) : make_smem_desc_mn<32>(addr); };
            int stage = 0;
            uint32_t phase = 0;
            uint32_t g = 0;
            uint32_t it = 0;
            long long st_t0 = p.stats ? clock64() : 0, st_full = 0, st_empty = 0, st_units = 0;
            for (int unit = cluster_id; unit >= 0; unit = next_unit(unit, it, false)) {
                const int kb0 = (unit / num_tiles) * kbs, kb1 = min(num_kb, kb0 + kbs);
                const int num_chunks = chunk_count(kb1 - kb0);
                ++st_units;
                for (int ch = 0; ch < num_chunks; ++ch, ++g) {
                    const uint32_t buf = g & 1;
                    if (p.stats) {
                        const long long t = clock64();
                        mbar_wait(tmem_empty_bar(buf), ((g >> 1) & 1) ^ 1);
                        st_empty += clock64() - t;
                    } else
                        mbar_wait(tmem_empty_bar(buf), ((g >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * Cfg::TILE_N;
                    const int kb_begin = kb0 + chunk_begin(ch);
                    const int kb_end = min(kb1, kb0 + chunk_begin(ch + 1));
                    for (int kb = kb_begin; kb < kb_end; ++kb) {
                        if (p.stats) {
                            const long long t = clock64();
                            mbar_wait(full_bar(stage), phase);
                            st_full += clock64() - t;
                        } else
                            mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        const uint32_t sb = sa + Cfg::PLANES * Cfg::A_BYTES;
                        const uint64_t da_hi = A_MN ? desc_mn(sa) : desc_k(sa);
                        const uint64_t db_hi = B_MN ? desc_mn(sb) : desc_k(sb);
                        const uint64_t da_lo = A_MN ? desc_mn(sa + Cfg::A_BYTES) : desc_k(sa + Cfg::A_BYTES);
                        const uint64_t db_lo = B_MN ? desc_mn(sb + Cfg::B_BYTES) : desc_k(sb + Cfg::B_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // one k-slice (32 bytes of k): K-major -> 32 bytes further along the 128-byte row; MN-major -> MMA_K rows further
                            const uint64_t koff_a = (uint64_t)((A_MN ? k * MN_KSTEP_BYTES : k * 32) >> 4);
                            const uint64_t koff_b = (uint64_t)((B_MN ? k * MN_KSTEP_BYTES : k * 32) >> 4);
                            const uint32_t first = (kb == kb_begin && k == 0) ? 0u : 1u;
                            if (TERMS == 3) {
                                umma_2sm<KIND>(tmem_d, da_lo + koff_a, db_hi + koff_b, idesc, first);
                                umma_2sm<KIND>(tmem_d, da_hi + koff_a, db_lo + koff_b, idesc, 1u);
                                umma_2sm<KIND>(tmem_d, da_hi + koff_a, db_hi + koff_b, idesc, 1u);
                            } else {
                                umma_2sm<KIND>(tmem_d, da_hi + koff_a, db_hi + koff_b, idesc, first);
                            }
                        }
                        umma_commit_2sm(empty_bar(stage), 0x3);   // frees the stage in BOTH CTAs
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit_2sm(tmem_full_bar(buf), 0x3);     // chunk complete -> both CTAs' epilogues
                }
            }
            if (p.stats) {
                unsigned long long* o = p.stats + (size_t)(blockIdx.x >> 1) * 8;
                o[0] = clock64() - st_t0; o[1] = st_full; o[2] = st_empty; o[4] = st_units;
            }
        }
        __syncwarp();
    } else {
        // ================================================================= epilogue (warps 2..9, both CTAs, own 128 rows)
        constexpr int CPW = Cfg::COLS_PER_WARP;
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int col0 = (ew >> 2) * CPW;
        uint32_t g = 0;
        uint32_t it = 0;
        for (int unit = cluster_id; unit >= 0; unit = next_unit(unit, it, true)) {
            int m_blk, n_blk;
            tile_coords(unit % num_tiles, m_blk, n_blk);
            const int split = unit / num_tiles;
            const int kb0 = split * kbs, kb1 = min(num_kb, kb0 + kbs);
            const int num_chunks = chunk_count(kb1 - kb0);
            float acc[CPW];
            // The store phase reads the relu-mask source (dX of the MLP) or the old C (+=) of this warp's 32 x 128 slice straight from
            // global memory; ncu showed those loads stalling the epilogue warps (which then cannot drain TMEM: 67.6 % tensor-pipe
            // activity on dX vs 93 % on dW).  Pull the slice into L2 now, while the mainloop of this tile runs: one bulk prefetch per row.
            if (p.c_vec_ok && (p.mask_src || p.accumulate)) {
                const int prow = m_blk * Cfg::TILE_M + (int)rank * 128 + quad * 32 + lane;
                const int pn0 = n_blk * Cfg::TILE_N + col0;
                const int pcols = min(CPW, p.N - pn0);
                if (prow < p.M && pcols > 0) {
                    const float* src = p.mask_src ? p.mask_src + (size_t)prow * p.N + pn0
                                                  : p.C + (size_t)split * p.M * p.N + (size_t)prow * p.N + pn0;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(pcols * 4)) : "memory");
                }
            }
            for (int ch = 0; ch < num_chunks; ++ch, ++g) {
                const uint32_t buf = g & 1;
                mbar_wait(tmem_full_bar(buf), (g >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < CPW / 32; ++c) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * Cfg::TILE_N + col0 + c * 32);
                    tmem_ld_32x32(taddr, r);
                    tmem_ld_wait();
                    if (ch == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c * 32 + j] = __uint_as_float(r[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {   // the leader's barrier
                    if (p.debug & 2) mbar_arrive_cluster_release(mapa_shared(tmem_empty_bar(buf), 0));
                    else mbar_arrive_cluster(mapa_shared(tmem_empty_bar(buf), 0));
                }
            }
            const int row = m_blk * Cfg::TILE_M + (int)rank * 128 + quad * 32 + lane;
            if (p.debug & 1) continue;
            const long long st_s0 = p.stats ? clock64() : 0;
            if (p.c_vec_ok) {
                // coalesced path: 32 x 32 blocks through the warp's shared-memory tile
                float* stage = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw))) + ew * 1024;
                const int row_base = row - lane;
                const int rows_valid = p.M - row_base;     // may be <= 0 or > 32: the helpers compare r < rows_valid
                const float rs = (p.row_scale && row < p.M) ? __ldg(p.row_scale + row) : 1.f;
                const size_t tile_off = (size_t)row_base * p.N;
                float* cbase = p.C + (size_t)split * p.M * p.N + tile_off;
                // relu-gradient mask bits of this row's CPW columns: requested together, one L2 round trip per tile
                uint32_t wbits[CPW / 32];
                if (p.bits_in) {
#pragma unroll
                    for (int c = 0; c < CPW / 32; ++c) {
                        const int n0 = n_blk * Cfg::TILE_N + col0 + c * 32;
                        wbits[c] = (row < p.M && n0 < p.N) ? __ldg(p.bits_in + (size_t)row * p.bits_words + (n0 >> 5)) : 0u;
                    }
                }
#pragma unroll
                for (int c = 0; c < CPW / 32; ++c) {
                    const int n0 = n_blk * Cfg::TILE_N + col0 + c * 32;
                    const int cols_valid = p.N - n0;
                    if (cols_valid <= 0 || rows_valid <= 0) continue;   // warp-uniform (no break: keeps acc[] statically indexed)
                    float(&v)[32] = *reinterpret_cast<float(*)[32]>(&acc[c * 32]);   // in place: acc is dead after the store
                    if (p.col_scale) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (j < cols_valid) {
                                const float4 cs = __ldg(reinterpret_cast<const float4*>(p.col_scale + n0 + j));
                                v[j] = (v[j] * rs) * cs.x; v[j + 1] = (v[j + 1] * rs) * cs.y;
                                v[j + 2] = (v[j + 2] * rs) * cs.z; v[j + 3] = (v[j + 3] * rs) * cs.w;
                            }
                    } else if (p.row_scale) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= rs;
                    }
                    if (p.accumulate) epi_gather<false>(stage, lane, v, cbase + n0, (size_t)p.N, rows_valid, cols_valid);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (j < cols_valid) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                            }
                    }
                    if (p.mask_src) epi_gather<true>(stage, lane, v, p.mask_src + tile_off + n0, (size_t)p.N, rows_valid, cols_valid);
                    if (p.bits_in) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = ((wbits[c] >> j) & 1u ? 1.f : 0.f) * v[j];
                    }
                    if (p.bits_out) {
                        uint32_t w = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) w |= (v[j] >= 0.f ? 1u : 0u) << j;
                        if (row < p.M) p.bits_out[(size_t)row * p.bits_words + (n0 >> 5)] = w;
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (v[j] >= 0.f ? 1.f : 0.f) * v[j];
                    }
                    epi_scatter(stage, lane, v, cbase + n0, (size_t)p.N, rows_valid, cols_valid);
                    if (p.C2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (v[j] >= 0.f ? 1.f : 0.f) * v[j];
                        epi_scatter(stage, lane, v, p.C2 + tile_off + n0, (size_t)p.N, rows_valid, cols_valid);
                    }
                }
            } else if (row < p.M) {
                float* crow = p.C + (size_t)split * p.M * p.N + (size_t)row * p.N;
                const int nbase = n_blk * Cfg::TILE_N + col0;
                const float rs = p.row_scale ? __ldg(p.row_scale + row) : 1.f;
#pragma unroll
                for (int j = 0; j < CPW; j += 4) {
                    const int n0 = nbase + j;
                    if (n0 >= p.N) continue;
                    if (p.c_vec_ok && n0 + 4 <= p.N) {
                        epilogue_store4(p, crow, (size_t)row * p.N, n0, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]), rs);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n0 + e < p.N) epilogue_store1(p, crow, (size_t)row * p.N, n0 + e, acc[j + e], rs);
                    }
                }
            }
            if (p.stats && leader && warp == 2 && lane == 0) p.stats[(size_t)(blockIdx.x >> 1) * 8 + 3] += clock64() - st_s0;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still read this CTA's shared memory / signal its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ operand prep kernels
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rna_tf32(x);
    if (!isfinite(hi)) {  // inf/nan input, or rounding overflowed: keep the top bits, no correction term
        hi = isfinite(x) ? __uint_as_float(__float_as_uint(x) & 0xffffe000u) : x;
        lo = isfinite(x) ? rna_tf32(x - hi) : 0.f;
        return;
    }
    lo = rna_tf32(x - hi);
}

// src [R x K] row-major (K contiguous) -> hi / lo planes [R x ldp]; lo may be NULL (TF32 fast mode: hi = rna_tf32(x))
__global__ void __launch_bounds__(256) prep_kmajor_kernel(size_t R, size_t K, size_t ldp, const float* __restrict__ src, float* __restrict__ hi,
                                                          float* __restrict__ lo, int vec) {
    if (vec) {
        const size_t kp = K / 4;
        const size_t total = R * kp;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const size_t r = i / kp, c = (i % kp) * 4;
            const Pack<float> x = ld_stream(src + r * K + c);
            Pack<float> h, l;
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32(x.v[e], h.v[e], l.v[e]);
            st_stream(hi + r * ldp + c, h);
            if (lo) st_stream(lo + r * ldp + c, l);
        }
    } else {
        const size_t total = R * K;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const size_t r = i / K, c = i % K;
            float h, l;
            split_tf32(__ldg(src + i), h, l);
            hi[r * ldp + c] = h;
            if (lo) lo[r * ldp + c] = l;
        }
    }
}

// src [K x R] row-major (R contiguous) -> transposed hi / lo planes [R x ldp] (K contiguous)
__global__ void __launch_bounds__(256) prep_transpose_kernel(size_t K, size_t R, size_t ldp, const float* __restrict__ src,
                                                             float* __restrict__ hi, float* __restrict__ lo) {
    __shared__ float tile[64][65];
    const size_t tiles_r = (R + 63) / 64, tiles_k = (K + 63) / 64;
    const size_t ntiles = tiles_r * tiles_k;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t k0 = (t / tiles_r) * 64, r0 = (t % tiles_r) * 64;
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
            const size_t k = k0 + threadIdx.y + i;
#pragma unroll
            for (int j = 0; j < 64; j += 32) {
                const size_t r = r0 + threadIdx.x + j;
                if (k < K && r < R) tile[threadIdx.y + i][threadIdx.x + j] = __ldg(src + k * R + r);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
            const size_t r = r0 + threadIdx.y + i;
#pragma unroll
            for (int j = 0; j < 64; j += 32) {
                const size_t k = k0 + threadIdx.x + j;
                if (k < K && r < R) {
                    float h, l;
                    split_tf32(tile[threadIdx.x + j][threadIdx.y + i], h, l);
                    hi[r * ldp + k] = h;
                    if (lo) lo[r * ldp + k] = l;
                }
            }
        }
        __syncthreads();
    }
}

// ---- 3xFP16 operand prep.  x is first multiplied by an exact power of two chosen per OUTPUT index (per row of A, per column
// of B) so that the largest magnitude along the contraction lands in [2^14, 2^15); then hi = fp16(x'), lo = fp16(x' - hi).  fp16
// and tf32 carry the same 11 significant bits, so hi + lo represents x' to 22 bits exactly like the tf32 split; the scaling only
// exists to fit fp16's 5-bit exponent (elements below 2^-16 of their row's maximum start to lose low bits: absolute floor
// 2^-40 of the row maximum).  The epilogue multiplies by the inverse scales, again exact.
__device__ __forceinline__ float scale_for_max(float m, float* inv) {
    if (!(m > 0.f) || !isfinite(m)) { *inv = 1.f; return 1.f; }
    int e;
    frexpf(m, &e);                 // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(15 - e) in [2^14, 2^15)
    e = 15 - e;
    e = e < -126 ? -126 : (e > 126 ? 126 : e);
    *inv = ldexpf(1.f, -e);
    return ldexpf(1.f, e);
}
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    const float h = __half2float(hi);
    lo = isfinite(h) ? __float2half_rn(x - h) : __float2half_rn(0.f);
}
__device__ __forceinline__ float absmax4(float m, const float4& v) {
    return fmaxf(fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fabsf(v.z))), fabsf(v.w));
}
__device__ __forceinline__ void split4_store(const float4& v, float sx, float sy, float sz, float sw, __half* hi, __half* lo) {
    __half h[4], l[4];
    split_f16(v.x * sx, h[0], l[0]); split_f16(v.y * sy, h[1], l[1]);
    split_f16(v.z * sz, h[2], l[2]); split_f16(v.w * sw, h[3], l[3]);
    *reinterpret_cast<uint2*>(hi) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo) = *reinterpret_cast<const uint2*>(l);
}

// K-contiguous operand [R x K] (K % 4 == 0): one 128-thread block per row, several blocks resident per SM so that one row's
// load latency hides behind the others' reductions and stores.  VPT float4 per thread hold a row of up to 512 * VPT elements in
// registers between the max pass and the split pass (one HBM read); VPT = 0: longer rows, re-read (L2).
template <int VPT>
__global__ void __launch_bounds__(128) prep16_rows_kernel(size_t R, size_t K, const float* __restrict__ src, __half* __restrict__ hi,
                                                          __half* __restrict__ lo, float* __restrict__ scale_inv) {
    constexpr int NV = VPT > 0 ? VPT : 1;
    __shared__ float red[4];
    __shared__ float bcast;
    const size_t nv = K / 4;
    for (size_t row = blockIdx.x; row < R; row += gridDim.x) {
        const float4* s4 = reinterpret_cast<const float4*>(src + row * K);
        float4 v[NV];
        float m = 0.f;
        if (VPT > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const size_t idx = threadIdx.x + (size_t)i * 128;
                if (idx < nv) {
                    const Pack<float> p = ld_stream(src + row * K + idx * 4);
                    v[i] = make_float4(p.v[0], p.v[1], p.v[2], p.v[3]);
                    m = absmax4(m, v[i]);
                }
            }
        } else {
            for (size_t idx = threadIdx.x; idx < nv; idx += 128) m = absmax4(m, __ldg(s4 + idx));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            const float t = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
            float inv;
            bcast = scale_for_max(t, &inv);
            scale_inv[row] = inv;
        }
        __syncthreads();
        const float sc = bcast;
        __half* h = hi + row * K;
        __half* l = lo + row * K;
        if (VPT > 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const size_t idx = threadIdx.x + (size_t)i * 128;
                if (idx < nv) split4_store(v[i], sc, sc, sc, sc, h + idx * 4, l + idx * 4);
            }
        } else {
            for (size_t idx = threadIdx.x; idx < nv; idx += 128) split4_store(__ldg(s4 + idx), sc, sc, sc, sc, h + idx * 4, l + idx * 4);
        }
        // (the next iteration's first __syncthreads orders this row's read of `bcast` before the next write)
    }
}

// MN-contiguous operand [K x R] (R % 4 == 0), scale per column r over all K rows.
// pass 1: grid (column blocks of 1024, slabs of rows) -> partial[slab][R] = max |x| over the slab
// partial_sum (optional): the same pass also produces the slab's column SUMS (sl_linear_bwd_params: bias gradient).
// row_factor (optional, [K], exact powers of two): the maxima are those of src[k][r] * row_factor[k] (the sums stay those of src)
__global__ void __launch_bounds__(256) prep16_colmax_kernel(size_t K, size_t R, size_t rows_per_slab, const float* __restrict__ src,
                                                            float* __restrict__ partial, float* __restrict__ partial_sum,
                                                            const float* __restrict__ row_factor) {
    const size_t c = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (c >= R) return;
    const size_t k0 = (size_t)blockIdx.y * rows_per_slab;
    const size_t k1 = k0 + rows_per_slab < K ? k0 + rows_per_slab : K;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (size_t k = k0; k < k1; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + k * R + c));
        const float f = row_factor ? __ldg(row_factor + k) : 1.f;
        m.x = fmaxf(m.x, fabsf(v.x * f)); m.y = fmaxf(m.y, fabsf(v.y * f)); m.z = fmaxf(m.z, fabsf(v.z * f)); m.w = fmaxf(m.w, fabsf(v.w * f));
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * R + c) = m;
    if (partial_sum) *reinterpret_cast<float4*>(partial_sum + (size_t)blockIdx.y * R + c) = sum;
}
// pass 2: fold the slabs -> scale[r], scale_inv[r]
//         (and the slab sums, in a fixed order -> sum_acc[r] += total: deterministic)
// block (32 columns, 8 slab lanes): each thread walks every 8th slab with 4 loads in flight, then the 8 lanes fold in order.
__global__ void __launch_bounds__(256) prep16_colscale_kernel(size_t R, int slabs, const float* __restrict__ partial, float* __restrict__ scale,
                                                              float* __restrict__ scale_inv, const float* __restrict__ partial_sum,
                                                              float* __restrict__ sum_acc) {
    __shared__ float sm[8][33], ss[8][33];
    const size_t r = (size_t)blockIdx.x * 32 + threadIdx.x;
    float m = 0.f, t = 0.f;
    if (r < R) {
#pragma unroll 4
        for (int s = threadIdx.y; s < slabs; s += 8) {
            m = fmaxf(m, partial[(size_t)s * R + r]);
            if (partial_sum) t += partial_sum[(size_t)s * R + r];
        }
    }
    sm[threadIdx.y][threadIdx.x] = m;
    ss[threadIdx.y][threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.y == 0 && r < R) {
#pragma unroll
        for (int y = 1; y < 8; ++y) {
            m = fmaxf(m, sm[y][threadIdx.x]);
            t += ss[y][threadIdx.x];
        }
        float inv;
        scale[r] = scale_for_max(m, &inv);
        scale_inv[r] = inv;
        if (partial_sum) sum_acc[r] += t;
    }
}
// pass 3: element-wise split with the column's scale (planes keep the [K x R] layout)
__global__ void __launch_bounds__(256) prep16_cols_kernel(size_t K, size_t R, const float* __restrict__ src, const float* __restrict__ scale,
                                                          __half* __restrict__ hi, __half* __restrict__ lo, const float* __restrict__ row_factor) {
    const size_t rv = R / 4;
    const size_t total = K * rv;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = (i % rv) * 4;
        const Pack<float> x = ld_stream(src + i * 4);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
        const float f = row_factor ? __ldg(row_factor + i / rv) : 1.f;   // (both factors are powers of two: x * f is exact)
        split4_store(make_float4(x.v[0] * f, x.v[1] * f, x.v[2] * f, x.v[3] * f), sc.x, sc.y, sc.z, sc.w, hi + i * 4, lo + i * 4);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// K-major operand plane [rows x K] with leading dimension ld (floats); box = BLOCK_K x box_rows
static int make_map(sl_ctx* ctx, CUtensorMap* map, const void* base, size_t rows, size_t K, size_t ld, int block_k, int box_rows, int elem = 4) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return sl_set_error(ctx, SL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * elem};
    cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = block_k * elem == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = enc(map, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sl_set_error(ctx, SL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%zu K=%zu ld=%zu", (int)r, rows, K, ld);
    return SL_OK;
}

template <int BLOCK_N, int BLOCK_K, int TERMS, int STAGES>
static int launch_cfg(sl_ctx* ctx, const GemmParams& p, const float* a_hi, const float* a_lo, size_t lda, const float* b_hi, const float* b_lo,
                      size_t ldb) {
    using Cfg = GemmCfg<BLOCK_N, BLOCK_K, TERMS, STAGES>;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_map(ctx, &ma_hi, a_hi, p.M, p.K, lda, BLOCK_K, BLOCK_M)) != SL_OK) return rc;
    if ((rc = make_map(ctx, &mb_hi, b_hi, p.N, p.K, ldb, BLOCK_K, BLOCK_N)) != SL_OK) return rc;
    ma_lo = ma_hi;
    mb_lo = mb_hi;
    if (TERMS == 3) {
        if ((rc = make_map(ctx, &ma_lo, a_lo, p.M, p.K, lda, BLOCK_K, BLOCK_M)) != SL_OK) return rc;
        if ((rc = make_map(ctx, &mb_lo, b_lo, p.N, p.K, ldb, BLOCK_K, BLOCK_N)) != SL_OK) return rc;
    }
    auto kern = gemm_tf32_kernel<BLOCK_N, BLOCK_K, TERMS, STAGES>;
    SL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int num_tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BLOCK_N - 1) / BLOCK_N);
    const int grid = num_tiles < ctx->num_sms ? num_tiles : ctx->num_sms;
    sl_ctx::ProfRec rec{};
    if (ctx->profiling) {
        cudaEventCreate(&rec.a);
        cudaEventCreate(&rec.b);
        rec.flops = 2.0 * p.M * (double)p.N * p.K;
        cudaEventRecord(rec.a, ctx->stream);
    }
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
    ctx->launches++;
    if (ctx->profiling) {
        cudaEventRecord(rec.b, ctx->stream);
        ctx->prof.push_back(rec);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sl_set_error(ctx, SL_ERR_CUDA, "gemm_tf32_kernel launch: %s", cudaGetErrorString(e));
    return SL_OK;
}

// 2-CTA (cta_group::2) launch: clusters of 2 CTAs, one 256x256 tile per cluster; each CTA's TMA box is 128 rows of A / of B
// MN-major plane: stored [K x rows] with leading dimension ld (rows contiguous); box = 32 rows(mn) x BLOCK_K k
//                 (fp16: box = 64 rows(mn) x BLOCK_K k, standard 128-byte swizzle)
static int make_map_mn(sl_ctx* ctx, CUtensorMap* map, const void* base, size_t rows, size_t K, size_t ld, int block_k, int elem = 4) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return sl_set_error(ctx, SL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)K};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * elem};
    cuuint32_t box[2] = {elem == 2 ? 64u : 32u, (cuuint32_t)block_k};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, elem == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sl_set_error(ctx, SL_ERR_CUDA, "cuTensorMapEncodeTiled (MN-major) failed (%d) rows=%zu K=%zu ld=%zu", (int)r, rows, K, ld);
    return SL_OK;
}

static int env_int(const char* name, int dflt);

template <int TERMS, int STAGES, bool A_MN, bool B_MN, int KIND = 0>
static int launch_cfg_2cta(sl_ctx* ctx, const GemmParams& p, const void* a_hi, const void* a_lo, size_t lda, const void* b_hi,
                           const void* b_lo, size_t ldb) {
    using Cfg = Gemm2Cfg<TERMS, STAGES, KIND>;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    auto mk = [&](CUtensorMap* m, const void* base, size_t rows, size_t ld, bool mn) {
        return mn ? make_map_mn(ctx, m, base, rows, p.K, ld, Cfg::BLOCK_K, Cfg::ELEM)
                  : make_map(ctx, m, base, rows, p.K, ld, Cfg::BLOCK_K, 128, Cfg::ELEM);
    };
    if ((rc = mk(&ma_hi, a_hi, p.M, lda, A_MN)) != SL_OK) return rc;
    if ((rc = mk(&mb_hi, b_hi, p.N, ldb, B_MN)) != SL_OK) return rc;
    ma_lo = ma_hi;
    mb_lo = mb_hi;
    if (TERMS == 3) {
        if ((rc = mk(&ma_lo, a_lo, p.M, lda, A_MN)) != SL_OK) return rc;
        if ((rc = mk(&mb_lo, b_lo, p.N, ldb, B_MN)) != SL_OK) return rc;
    }
    auto kern = gemm_tf32_2cta_kernel<TERMS, STAGES, A_MN, B_MN, KIND>;
    SL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int num_tiles = ((p.M + Cfg::TILE_M - 1) / Cfg::TILE_M) * ((p.N + Cfg::TILE_N - 1) / Cfg::TILE_N);
    // data-parallel runs may leave a few SMs to the communication kernels (SLICED_GEMM_RESERVE_SMS, default 0): the persistent gemm
    // CTAs own all of an SM's shared memory, so an NCCL kernel cannot co-reside with them
    const int reserve = ctx->nccl_comm ? env_int("SLICED_GEMM_RESERVE_SMS", 0) : 0;
    const int max_clusters = (ctx->num_sms - (reserve > 0 && reserve < ctx->num_sms - 2 ? reserve : 0)) / 2;
    // dynamic scheduling: one cluster per work unit, handed out by cluster launch control (see the kernel)
    const int grid = p.dynamic ? 2 * num_tiles * (p.splits > 0 ? p.splits : 1) : 2 * (num_tiles < max_clusters ? num_tiles : max_clusters);
    sl_ctx::ProfRec rec{};
    if (ctx->profiling) {
        cudaEventCreate(&rec.a);
        cudaEventCreate(&rec.b);
        rec.flops = 2.0 * p.M * (double)p.N * p.K;
        cudaEventRecord(rec.a, ctx->stream);
    }
    static unsigned long long* stats_dev = nullptr;
    const bool want_stats = env_int("SLICED_GEMM_STATS", 0) != 0;
    GemmParams ps = p;
    if (want_stats) {   // diagnosis only: synchronous, prints where the MMA issuer's clocks went
        if (!stats_dev) cudaMalloc(&stats_dev, 8 * sizeof(unsigned long long) * 65536);
        cudaMemsetAsync(stats_dev, 0, 8 * sizeof(unsigned long long) * (size_t)(grid / 2 < 65536 ? grid / 2 : 65536), ctx->stream);
        if (grid / 2 <= 65536) ps.stats = stats_dev;
    }
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, ps);
    if (ps.stats) {
        cudaStreamSynchronize(ctx->stream);
        std::vector<unsigned long long> h((size_t)(grid / 2) * 8);
        cudaMemcpy(h.data(), stats_dev, h.size() * 8, cudaMemcpyDeviceToHost);
        double tot = 0, full = 0, empty = 0, store = 0, units = 0, active = 0, tmax = 0;
        for (int c = 0; c < grid / 2; ++c) {
            if (!h[(size_t)c * 8 + 4]) continue;   // cluster whose unit was taken over before it launched
            active += 1; tot += h[(size_t)c * 8]; full += h[(size_t)c * 8 + 1]; empty += h[(size_t)c * 8 + 2]; store += h[(size_t)c * 8 + 3];
            units += h[(size_t)c * 8 + 4];
            if (h[(size_t)c * 8] > tmax) tmax = (double)h[(size_t)c * 8];
        }
        printf("gemm stats M %d N %d K %d splits %d A_MN %d B_MN %d mask %d relu %d: clusters %.0f units %.0f | MMA issuer clocks avg %.0f max %.0f | wait operands %.1f%% | "
               "wait tmem-empty %.1f%% | store phase per unit %.0f clocks (budget %d k-blocks)\n",
               p.M, p.N, p.K, p.splits, (int)A_MN, (int)B_MN, p.mask_src != nullptr, p.relu, active, units, tot / active, tmax, 100. * full / tot, 100. * empty / tot,
               store / units, 2 * (p.kc_first > p.kc_blocks ? p.kc_first : p.kc_blocks));
        fflush(stdout);
    }
    ctx->launches++;
    if (ctx->profiling) {
        cudaEventRecord(rec.b, ctx->stream);
        ctx->prof.push_back(rec);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sl_set_error(ctx, SL_ERR_CUDA, "gemm_tf32_2cta_kernel launch: %s", cudaGetErrorString(e));
    return SL_OK;
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

}  // namespace

// tile configuration: SLICED_GEMM_CFG = 0 auto | 1: 128x256x32 | 2: 128x256x16 | 3: 128x128x32 | 4: 2-CTA pairs, 256x256x32 per pair
static int sl_gemm_pick_cfg(sl_ctx* ctx, size_t M, size_t N) {
    int cfg = env_int("SLICED_GEMM_CFG", 0);
    if (cfg == 0) {
        const long pair_tiles = (long)((M + 255) / 256) * ((N + 255) / 256);
        cfg = (M >= 256 && N >= 256 && pair_tiles >= ctx->num_sms / 4) ? 4 : 3;   // enough 256x256 tiles to occupy half the CTA pairs
    }
    return cfg;
}

// folds split-K partials in split order and applies the epilogue (see sl_gemm_tc_planes)
// (mask bits: n % 32 == 0, so the 32 consecutive elements of a warp are the 32 bits of one word)
__global__ void __launch_bounds__(256) splitk_fold_kernel(size_t total, size_t n, int splits, const float* __restrict__ partial, float* C, int accumulate,
                                                          const float* __restrict__ bias, int relu, float* C2, const float* __restrict__ mask_src,
                                                          const uint32_t* __restrict__ bits_in, uint32_t* bits_out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float v = partial[i];
        for (int s = 1; s < splits; ++s) v += partial[(size_t)s * total + i];
        if (accumulate) v += C[i];
        if (bias) v += __ldg(bias + i % n);
        if (mask_src) v = (__ldg(mask_src + i) >= 0.f ? 1.f : 0.f) * v;
        if (bits_in) v = ((__ldg(bits_in + (i >> 5)) >> (i & 31)) & 1u ? 1.f : 0.f) * v;
        if (bits_out) {
            const uint32_t w = __ballot_sync(0xffffffffu, v >= 0.f);
            if ((threadIdx.x & 31) == 0) bits_out[i >> 5] = w;
        }
        if (relu) v = (v >= 0.f ? 1.f : 0.f) * v;
        C[i] = v;
        if (C2) C2[i] = (v >= 0.f ? 1.f : 0.f) * v;
    }
}

// 128-bit variant (n % 4 == 0, 16-byte aligned pointers; with mask bits n % 32 == 0): a thread folds 4 consecutive elements; the 8
// threads covering one mask word combine their 4-bit nibbles with three shuffles.  Same per-element arithmetic and order as above.
__global__ void __launch_bounds__(256) splitk_fold_vec_kernel(size_t total4, size_t n, int splits, const float* __restrict__ partial, float* C, int accumulate,
                                                              const float* __restrict__ bias, int relu, float* C2, const float* __restrict__ mask_src,
                                                              const uint32_t* __restrict__ bits_in, uint32_t* bits_out) {
    const size_t total = total4 * 4;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (size_t)gridDim.x * blockDim.x) {
        const size_t i = q * 4;
        const Pack<float> p0 = ld_stream(partial + i);
        float v[4] = {p0.v[0], p0.v[1], p0.v[2], p0.v[3]};
        for (int s = 1; s < splits; ++s) {
            const Pack<float> ps = ld_stream(partial + (size_t)s * total + i);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += ps.v[e];
        }
        if (accumulate) {
            const Pack<float> c = ld_pack(C + i);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += c.v[e];
        }
        if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + i % n));
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        if (mask_src) {
            const Pack<float> m = ld_stream(mask_src + i);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (m.v[e] >= 0.f ? 1.f : 0.f) * v[e];
        }
        if (bits_in) {
            const uint32_t w = __ldg(bits_in + (i >> 5)) >> (i & 31);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = ((w >> e) & 1u ? 1.f : 0.f) * v[e];
        }
        if (bits_out) {   // (total4 % 8 == 0 because n % 32 == 0: the 8 threads of a word are all inside the loop together)
            uint32_t w = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) w |= (v[e] >= 0.f ? 1u : 0u) << e;
            w <<= 4 * (threadIdx.x & 7);
            const uint32_t grp = 0xFFu << (threadIdx.x & 24);   // the 8 lanes of this word (they enter and leave the loop together)
            w |= __shfl_xor_sync(grp, w, 1);
            w |= __shfl_xor_sync(grp, w, 2);
            w |= __shfl_xor_sync(grp, w, 4);
            if ((threadIdx.x & 7) == 0) bits_out[i >> 5] = w;
        }
        Pack<float> o, o2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (relu) v[e] = (v[e] >= 0.f ? 1.f : 0.f) * v[e];
            o.v[e] = v[e];
            o2.v[e] = (v[e] >= 0.f ? 1.f : 0.f) * v[e];
        }
        st_pack(C + i, o);
        if (C2) st_pack(C2 + i, o2);
    }
}

// Runs the tensor-core kernel on prepared K-major planes: C[M x N] (=|+=) A[M x K] * B[N x K]^T (+ bias, relu).
// a_lo / b_lo are NULL in TF32 mode.  lda / ldb in floats, multiples of 4, planes 16-byte aligned.
// kind 1: the planes are fp16 (3xFP16 mode, 2-CTA kernel only; lda / ldb in halves, multiples of 8) and row_scale [M] / col_scale [N]
// hold the inverse operand scales applied by the epilogue.
int sl_gemm_tc_planes(sl_ctx* ctx, int M, int N, int K, const void* a_hi, const void* a_lo, size_t lda, const void* b_hi, const void* b_lo,
                      size_t ldb, float* C, const float* bias, int accumulate, int relu, float* c2, const float* mask_src, int a_mn, int b_mn,
                      int kind = 0, const float* row_scale = nullptr, const float* col_scale = nullptr, const uint32_t* bits_in = nullptr,
                      uint32_t* bits_out = nullptr) {
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.C = C; p.bias = bias; p.accumulate = accumulate; p.relu = relu; p.C2 = c2; p.mask_src = mask_src;
    p.bits_in = bits_in; p.bits_out = bits_out; p.bits_words = (N + 31) / 32; p.stats = nullptr;
    p.row_scale = row_scale; p.col_scale = col_scale;
    p.debug = env_int("SLICED_GEMM_DEBUG", 0);
    p.dynamic = env_int("SLICED_GEMM_SCHED", 1) != 0;   // 0: static persistent schedule | 1 (default): cluster launch control
    // L2 residency: when one operand is much smaller than the other (weights vs a batch of activations) and fits in a fraction of
    // L2, its loads are marked evict_last and the big, streamed operand's loads evict_first, so the stream does not flush it.
    p.hint_a = p.hint_b = 0;
    {
        const int hint_mode = env_int("SLICED_GEMM_L2HINT", 2);   // 0 off | 1 / 2 see below (1 measured harmful: 12 GB, the streamed
                                                                   // operand is still shared by the 16 tiles of its row band)
        const double a_bytes = (double)M * K * 4, b_bytes = (double)N * K * 4;   // hi + lo planes
        constexpr unsigned long long EVICT_FIRST = 0x12F0000000000000ull, EVICT_LAST = 0x14F0000000000000ull;
        if (hint_mode >= 1) {   // 1: small operand evict_last + streamed operand evict_first; 2: small operand evict_last only
            const unsigned long long stream = hint_mode == 1 ? EVICT_FIRST : 0ull;
            if (b_bytes * 4 <= a_bytes && b_bytes <= 80e6) { p.hint_b = EVICT_LAST; p.hint_a = stream; }
            else if (a_bytes * 4 <= b_bytes && a_bytes <= 80e6) { p.hint_a = EVICT_LAST; p.hint_b = stream; }
        }
    }
    p.c_vec_ok = (N % 4 == 0) && sl_aligned16(C) && (!bias || sl_aligned16(bias)) && (!c2 || sl_aligned16(c2)) && (!mask_src || sl_aligned16(mask_src)) &&
                 (!col_scale || sl_aligned16(col_scale));
    const bool three = a_lo != nullptr;
    // K-chunk (in k-blocks of 32) accumulated inside TMEM before promotion to fp32 registers; 0 = whole K (TF32 fast mode)
    p.kc_blocks = three ? env_int("SLICED_GEMM_KC", 4) : env_int("SLICED_GEMM_KC_TF32", 0);
    p.kc_first = three ? env_int("SLICED_GEMM_KC_FIRST", 8) : 0;
    // tile configuration: SLICED_GEMM_CFG = 0 auto | 1: 128x256x32 | 2: 128x256x16 (swizzle 64B, deeper ring) | 3: 128x128x32
    //                                       | 4: 2-CTA pairs (cta_group::2), 256x256x32 per pair
    const int cfg = sl_gemm_pick_cfg(ctx, M, N);
    if (cfg != 4 && (a_mn || b_mn || kind)) return sl_set_error(ctx, SL_ERR_INVALID_ARG, "MN-major / fp16 planes need the 2-CTA kernel");
    if ((bits_in || bits_out) && (cfg != 4 || !p.c_vec_ok || N % 32 != 0))
        return sl_set_error(ctx, SL_ERR_INVALID_ARG, "mask bits need the 2-CTA kernel, N % 32 == 0 and aligned pointers");
    p.splits = 1;
    // Tile raster.  Measured on the MLP shapes (tools/raster_sweep.py): narrow groups win — with 2 blocks of the LONG dimension per
    // group the tiles running concurrently span the whole short dimension, so the smaller operand (the weights: 134 MB of hi/lo
    // planes, about one L2) is shared by every CTA pair while the big operand streams through exactly once.
    // Re-measured with DRAM counters for the 3xFP16 kernel (tools/raster_traffic.py, profiles/r1_raster_traffic.txt): a long
    // M with a short N (activations x weights) is best swept in bands of 8 n-blocks (half of the weights stay L2-resident, marked
    // evict_last, while the activations stream through twice): 6.9 -> 5.6 GB and 4.54 -> 4.33 ms; a deep K (weight gradient) wants
    // wider groups (8): 15.3 -> 11.9 GB, 4.73 -> 4.53 ms.
    int def_group = 2, def_along_n = N > M ? 1 : 0;
    if ((long)K >= 4l * (M > N ? M : N)) { def_group = 8; def_along_n = 0; }
    else if ((long)M >= 4l * N) { def_group = 8; def_along_n = 1; }
    else if ((long)N >= 4l * M) { def_group = 8; def_along_n = 0; }
    p.group = env_int("SLICED_GEMM_GROUP", def_group);
    p.group_along_n = env_int("SLICED_GEMM_GROUP_N", def_along_n);
    if (p.group < 1) p.group = 1;
    if (cfg == 4) {  // cta_group::2, 256x256 tile per CTA pair
        // Wave quantisation: with T tiles on 74 CTA pairs the last wave is T mod 74 wide (dW of the MLP: 256 tiles = 3.46 waves
        // -> 4).  When that wastes > 8 % and K is long, split K so that tiles x splits fills whole waves; the partial tiles go
        // to scratch and are folded in split order (deterministic), the fold also applies the epilogue.
        const long tiles = (long)((M + 255) / 256) * ((N + 255) / 256);
        const long pairs = ctx->num_sms / 2;
        auto eff = [&](long units) { return (double)units / (double)(((units + pairs - 1) / pairs) * pairs); };
        int best = 1;
        const int block_k = kind == 1 ? 64 : 32;
        const int num_kb = (K + block_k - 1) / block_k;
        const int max_splits = env_int("SLICED_GEMM_MAX_SPLITS", 4);
        // cost model (microseconds): waves x k-blocks per unit at the kernel's measured rate, plus what the split costs in HBM
        // traffic (sp partial tiles written instead of one, then read back and folded).  At K = 65536 (the MLP's weight gradient on
        // one GPU) the split wins 0.3 ms; at K = 8192 (the same gemm on 8 GPUs) the fold costs more than the partial last wave.
        const double rate = kind == 1 ? 470e12 : (three ? 240e12 : 600e12);   // algorithmic flop/s of the whole GPU in this mode
        const double us_wave_kb = 2.0 * 256 * 256 * block_k / (rate / (double)pairs) * 1e6;
        auto cost = [&](int sp) {
            const long units = tiles * sp;
            const long waves = (units + pairs - 1) / pairs;
            const int kbs = (num_kb + sp - 1) / sp;
            // (a deliberately pessimistic fold cost: with the vectorised fold kernel a 2-way split at K = 8192 — the weight gradient
            // of an 8-GPU run — measures the same as no split, 3.28 vs 3.29 ms per step, so the variant with fewer launches is kept)
            const double fold_us = sp > 1 ? 2.0 * sp * (double)M * N * 4.0 / 4.0e12 * 1e6 + 4.0 : 0.0;
            return (double)waves * kbs * us_wave_kb + fold_us;
        };
        if (eff(tiles) < 0.92)
            for (int sp = 2; sp <= max_splits; ++sp) {
                const int kbs = (num_kb + sp - 1) / sp;
                if (kbs * block_k < 2048 || (long)(sp - 1) * kbs >= num_kb) continue;   // keep every split >= 2048 deep and non-empty
                if (cost(sp) < 0.97 * cost(best)) best = sp;
            }
        GemmParams q = p;
        float* final_c = C;
        if (best > 1) {
            void* ws = nullptr;
            int rc = sl_ws_reserve(ctx, (size_t)best * M * N * sizeof(float), &ws);
            if (rc != SL_OK) return rc;
            q.splits = best;
            q.C = (float*)ws;
            q.bias = nullptr; q.accumulate = 0; q.relu = 0; q.C2 = nullptr; q.mask_src = nullptr; q.bits_in = nullptr; q.bits_out = nullptr;
            q.c_vec_ok = (N % 4 == 0);
        }
        int rc;
#define SL_2CTA(AM, BM) (kind == 1 ? launch_cfg_2cta<3, 3, AM, BM, 1>(ctx, q, a_hi, a_lo, lda, b_hi, b_lo, ldb) \
                         : three   ? launch_cfg_2cta<3, 3, AM, BM>(ctx, q, a_hi, a_lo, lda, b_hi, b_lo, ldb)     \
                                   : launch_cfg_2cta<1, 6, AM, BM>(ctx, q, a_hi, a_lo, lda, b_hi, b_lo, ldb))
        if (a_mn && b_mn) rc = SL_2CTA(true, true);
        else if (a_mn) rc = SL_2CTA(true, false);
        else if (b_mn) rc = SL_2CTA(false, true);
        else rc = SL_2CTA(false, false);
#undef SL_2CTA
        if (rc != SL_OK || best == 1) return rc;
        const size_t total = (size_t)M * N;
        const size_t cap = (size_t)ctx->num_sms * 8;
        size_t blocks = (total + 255) / 256;
        if (p.c_vec_ok && total % 32 == 0 && !env_int("SLICED_GEMM_FOLD_SCALAR", 0)) {
            blocks = (total / 4 + 255) / 256;
            SL_LAUNCH(ctx, splitk_fold_vec_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, total / 4, (size_t)N, best, (const float*)q.C, final_c,
                      p.accumulate, p.bias, p.relu, p.C2, p.mask_src, p.bits_in, p.bits_out);
            return SL_OK;
        }
        SL_LAUNCH(ctx, splitk_fold_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, total, (size_t)N, best, (const float*)q.C, final_c, p.accumulate,
                  p.bias, p.relu, p.C2, p.mask_src, p.bits_in, p.bits_out);
        return SL_OK;
    }
    const float *fa_hi = (const float*)a_hi, *fa_lo = (const float*)a_lo, *fb_hi = (const float*)b_hi, *fb_lo = (const float*)b_lo;
    if (three) {
        if (cfg == 1) return launch_cfg<256, 32, 3, 2>(ctx, p, fa_hi, fa_lo, lda, fb_hi, fb_lo, ldb);
        if (cfg == 2) return launch_cfg<256, 16, 3, 4>(ctx, p, fa_hi, fa_lo, lda, fb_hi, fb_lo, ldb);
        return launch_cfg<128, 32, 3, 3>(ctx, p, fa_hi, fa_lo, lda, fb_hi, fb_lo, ldb);
    } else {
        if (cfg == 1) return launch_cfg<256, 32, 1, 4>(ctx, p, fa_hi, fa_lo, lda, fb_hi, fb_lo, ldb);
        if (cfg == 2) return launch_cfg<256, 16, 1, 8>(ctx, p, fa_hi, fa_lo, lda, fb_hi, fb_lo, ldb);
        return launch_cfg<128, 32, 1, 6>(ctx, p, fa_hi, fa_lo, lda, fb_hi, fb_lo, ldb);
    }
}

// Splits (and if needed transposes) one operand into K-major planes.  `rows` = M or N, `K` = contraction length.
// k_contiguous: src is [rows x K]; otherwise src is [K x rows].
int sl_gemm_prep_operand(sl_ctx* ctx, const float* src, size_t rows, size_t K, int k_contiguous, float* hi, float* lo, size_t ldp) {
    const size_t cap = (size_t)ctx->num_sms * 8;
    if (k_contiguous) {
        const int vec = (K % 4 == 0) && sl_aligned16(src) && (ldp % 4 == 0);
        const size_t total = vec ? rows * (K / 4) : rows * K;
        size_t blocks = (total + 255) / 256;
        SL_LAUNCH(ctx, prep_kmajor_kernel, (unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap), 256, 0, rows, K, ldp, src, hi, lo, vec);
    } else {
        const size_t ntiles = ((rows + 63) / 64) * ((K + 63) / 64);
        SL_LAUNCH(ctx, prep_transpose_kernel, (unsigned)(ntiles < cap ? (ntiles ? ntiles : 1) : cap), dim3(32, 8, 1), 0, K, rows, ldp, src, hi, lo);
    }
    return SL_OK;
}

static bool tc_eligible(size_t m, size_t n, size_t k) {
    if (m > 0x7fffffff || n > 0x7fffffff || k > 0x7fffffff) return false;
    if (env_int("SLICED_GEMM_TC_FORCE", 0)) return true;  // tests: push even tiny / ragged shapes through the tcgen05 kernel
    // 128-row tcgen05 tiles: below this the CUDA-core kernel wins (and skinny heads like N = 10 stay exact fp32)
    const size_t min_dim = (size_t)env_int("SLICED_GEMM_TC_MIN_DIM", 64);
    return m >= min_dim && n >= min_dim && k >= 32 && (double)m * (double)n * (double)k >= (double)(1 << 22);
}

// ---- operand-plane reuse.  Inside a scope (sl_gemm_scope_begin/_end) the planes of an operand that is split "flat" (same
// layout as the source: K-major with K % 4 == 0, or MN-major) are kept in their own allocation and reused by every later gemm of
// the scope that reads the same buffer: out_grad in both halves of gemm_grad, an activation in its forward gemm and in the
// weight-gradient gemm, a weight in forward and in the input-gradient gemm.  Slots are handed out in call order, so a loop
// that issues the same sequence every iteration never reallocates.
static int plane_cache_get(sl_ctx* ctx, const float* src, size_t elems, bool three, float** hi, float** lo, bool* hit) {
    for (auto& e : ctx->plane_cache)
        if (e.valid && e.src == src && e.elems == elems && (!three || e.lo)) {
            *hi = e.hi; *lo = three ? e.lo : nullptr; *hit = true;
            return SL_OK;
        }
    if (ctx->plane_cursor >= ctx->plane_cache.size()) ctx->plane_cache.push_back(sl_ctx::PlaneEntry{nullptr, 0, nullptr, nullptr, 0, false});
    sl_ctx::PlaneEntry& e = ctx->plane_cache[ctx->plane_cursor++];
    const size_t bytes = ((elems * 4) + 255) & ~size_t(255);
    if (e.cap_bytes < bytes || (three && !e.lo)) {
        SL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (e.hi) cudaFree(e.hi);
        if (e.lo) cudaFree(e.lo);
        e.hi = e.lo = nullptr;
        SL_CUDA(ctx, cudaMalloc((void**)&e.hi, bytes));
        if (three) SL_CUDA(ctx, cudaMalloc((void**)&e.lo, bytes));
        e.cap_bytes = bytes;
    }
    e.src = src; e.elems = elems; e.valid = true;
    *hi = e.hi; *lo = three ? e.lo : nullptr; *hit = false;
    return SL_OK;
}

// fused epilogue request (f32): v = acc (+ C) (+ bias[n]) ; v *= (mask_src >= 0) ; [relu in place] ; C = v ; C2 = relu(v)
struct Epi {
    const float* bias = nullptr;
    int relu = 0;
    float* c2 = nullptr;
    const float* mask_src = nullptr;
    // the relu mask as one bit per element (see GemmParams): bits_out is taken before the in-place relu, bits_in multiplies like mask_src.
    // Only the 2-CTA tensor-core path (and the skinny nt kernel for bits_in) fuses them: sl_linear_*_bits pick the path.
    const uint32_t* bits_in = nullptr;
    uint32_t* bits_out = nullptr;
    bool any() const { return bias || relu || c2 || mask_src || bits_in || bits_out; }
};

// one thread per mask word: bits of z (row-major [rows x cols]) -> word, then relu in place.  Generic fallback of sl_linear_fwd_bits.
__global__ void __launch_bounds__(256) bits_relu_kernel(size_t rows, size_t cols, size_t nw, float* z, uint32_t* bits) {
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < rows * nw; w += (size_t)gridDim.x * blockDim.x) {
        const size_t r = w / nw, c0 = (w % nw) * 32;
        uint32_t m = 0;
        for (int j = 0; j < 32 && c0 + j < cols; ++j) {
            const float v = z[r * cols + c0 + j];
            const bool pos = v >= 0.f;
            m |= (pos ? 1u : 0u) << j;
            z[r * cols + c0 + j] = (pos ? 1.f : 0.f) * v;
        }
        bits[w] = m;
    }
}
// x[i] *= bit(i): generic fallback of sl_linear_bwd_input_relu_bits
__global__ void __launch_bounds__(256) bits_apply_kernel(size_t rows, size_t cols, size_t nw, float* x, const uint32_t* __restrict__ bits) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / cols, c = i % cols;
        x[i] = ((__ldg(bits + r * nw + (c >> 5)) >> (c & 31)) & 1u ? 1.f : 0.f) * x[i];
    }
}

// the same epilogue as a separate pass, for shapes that do not run on the tensor-core kernel (tiny / skinny)
__global__ void __launch_bounds__(256) epilogue_pass_kernel(size_t total, size_t n, float* C, const float* __restrict__ bias, int relu, float* C2,
                                                            const float* __restrict__ mask_src) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float v = C[i];
        if (bias) v += __ldg(bias + i % n);
        if (mask_src) v = (__ldg(mask_src + i) >= 0.f ? 1.f : 0.f) * v;
        if (relu) v = (v >= 0.f ? 1.f : 0.f) * v;
        C[i] = v;
        if (C2) C2[i] = (v >= 0.f ? 1.f : 0.f) * v;
    }
}

// 3xFP16 path: scale + split both operands into fp16 hi / lo planes IN THEIR OWN LAYOUT (K-major or MN-major, never transposed)
// and run the kind::f16 2-CTA kernel; the epilogue undoes the scaling.
// *inv_out receives the [mn] vector of inverse scales the epilogue needs (scale_inv, or a cached one).
// colsum_acc (MN-major operand only): colsum_acc[j] += sum over k of src[k][j], computed in the same pass as the column maxima.
static int prep16_operand(sl_ctx* ctx, const float* src, size_t mn, size_t k, bool k_contiguous, __half* hi, __half* lo, float* scale,
                          float* scale_inv, const float** inv_out, float* colsum_acc, const float* row_factor = nullptr) {
    const size_t cap = (size_t)ctx->num_sms * 8;
    *inv_out = scale_inv;
    if (k_contiguous) {  // [mn x k]
        const unsigned grid = (unsigned)(mn < cap * 8 ? mn : cap * 8);
#define SL_ROWS(VPTV) SL_LAUNCH(ctx, (prep16_rows_kernel<VPTV>), grid, 128, 0, mn, k, src, hi, lo, scale_inv)
        if (k <= 2048) SL_ROWS(4);
        else if (k <= 4096) SL_ROWS(8);
        else if (k <= 8192) SL_ROWS(16);
        else SL_ROWS(0);
#undef SL_ROWS
        return SL_OK;
    }
    // [k x mn]: column maxima in two deterministic passes, then the element-wise split.
    // row_factor: the buffer is split as src[k][mn] * row_factor[k] (the other operand's row scales folded in, see gemm_f16x3).
    {
        const size_t col_blocks = (mn / 4 + 255) / 256;
        size_t slabs = cap / col_blocks;
        if (slabs < 1) slabs = 1;
        if (slabs > (k + 31) / 32) slabs = (k + 31) / 32;   // at least 32 rows per slab
        if (slabs > 65535) slabs = 65535;
        const size_t rows_per_slab = (k + slabs - 1) / slabs;
        slabs = (k + rows_per_slab - 1) / rows_per_slab;
        void* partial = nullptr;
        int rc = sl_ws_reserve(ctx, (colsum_acc ? 2 : 1) * slabs * mn * sizeof(float), &partial);
        if (rc != SL_OK) return rc;
        float* psum = colsum_acc ? (float*)partial + slabs * mn : nullptr;
        SL_LAUNCH(ctx, prep16_colmax_kernel, dim3((unsigned)col_blocks, (unsigned)slabs, 1), 256, 0, k, mn, rows_per_slab, src, (float*)partial, psum, row_factor);
        SL_LAUNCH(ctx, prep16_colscale_kernel, (unsigned)((mn + 31) / 32), dim3(32, 8, 1), 0, mn, (int)slabs, (const float*)partial, scale, scale_inv,
                  (const float*)psum, colsum_acc);
    }
    const size_t total = k * (mn / 4);
    size_t blocks = (total + 255) / 256;
    SL_LAUNCH(ctx, prep16_cols_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, k, mn, src, (const float*)scale, hi, lo, row_factor);
    return SL_OK;
}

// Row-scaled fp16 planes of a K-contiguous operand, kept for the rest of the scope (slots handed out in call order, grow-only)
static sl_ctx::RowPlanes* rowplanes_find(sl_ctx* ctx, const void* src, size_t rows, size_t cols) {
    for (auto& e : ctx->rowplane_cache)
        if (e.valid && e.src == src && e.rows == rows && e.cols == cols) return &e;
    return nullptr;
}
static int rowplanes_new(sl_ctx* ctx, const void* src, size_t rows, size_t cols, sl_ctx::RowPlanes** out) {
    if (ctx->rowplane_cursor >= ctx->rowplane_cache.size())
        ctx->rowplane_cache.push_back(sl_ctx::RowPlanes{nullptr, 0, 0, nullptr, nullptr, nullptr, 0, 0, false});
    sl_ctx::RowPlanes& e = ctx->rowplane_cache[ctx->rowplane_cursor++];
    const size_t bytes = (rows * cols * 2 + 255) & ~size_t(255);
    if (e.cap_bytes < bytes || e.cap_rows < rows) {
        SL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (e.hi) cudaFree(e.hi);
        if (e.lo) cudaFree(e.lo);
        if (e.inv) cudaFree(e.inv);
        e.hi = e.lo = nullptr; e.inv = nullptr; e.cap_bytes = e.cap_rows = 0;
        SL_CUDA(ctx, cudaMalloc(&e.hi, bytes));
        SL_CUDA(ctx, cudaMalloc(&e.lo, bytes));
        SL_CUDA(ctx, cudaMalloc((void**)&e.inv, rows * sizeof(float)));
        e.cap_bytes = bytes; e.cap_rows = rows;
    }
    e.src = src; e.rows = rows; e.cols = cols; e.valid = true;
    *out = &e;
    return SL_OK;
}

int sl_allreduce_sum_async(sl_ctx* ctx, int dtype, void* buf, size_t n);

// m_chunks > 1 (MN-major A only): the product is computed in m_chunks row blocks of C, one kernel launch each, from the same planes;
// with exchange_chunks every finished block is handed to sl_allreduce_sum_async so that its exchange overlaps the next block's gemm.
static int gemm_f16x3(sl_ctx* ctx, int trans_a, int trans_b, size_t m, size_t n, size_t k, const float* a, const float* b, float* c, int accumulate,
                      const Epi& epi, float* b_colsum_acc = nullptr, int m_chunks = 1, bool exchange_chunks = false) {
    const bool a_kc = !trans_a, b_kc = trans_b != 0;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t a_plane = up(m * k * 2), b_plane = up(n * k * 2), sm = up(m * 4), sn = up(n * 4);
    char* ws = nullptr;
    int rc = sl_ws2_reserve(ctx, 2 * a_plane + 2 * b_plane + 2 * sm + 2 * sn, (void**)&ws);
    if (rc != SL_OK) return rc;
    __half* a_hi = (__half*)ws;
    __half* a_lo = (__half*)(ws + a_plane);
    __half* b_hi = (__half*)(ws + 2 * a_plane);
    __half* b_lo = (__half*)(ws + 2 * a_plane + b_plane);
    float* a_sc = (float*)(ws + 2 * a_plane + 2 * b_plane);
    float* a_inv = (float*)((char*)a_sc + sm);
    float* b_sc = (float*)((char*)a_inv + sm);
    float* b_inv = (float*)((char*)b_sc + sn);
    const float *a_inv_use = nullptr, *b_inv_use = nullptr;
    // Inside a scope a K-contiguous operand's row-scaled planes are kept (and found again): the forward gemm's split of an
    // activation [batch x width] serves the weight-gradient gemm of the same step, which contracts over the batch and reads the very
    // same planes as its MN-major operand.  Their scaling is per ROW (= per k of that product), so it is undone inside the
    // contraction: the OTHER operand is split as other[k][.] * (1 / s_k) — one exact power-of-two factor per row folded into its own
    // column split — and the reused operand contributes no epilogue scale.  (Before: every activation was split twice per step,
    // once per row and once per column: 4 of the 7 big operand passes of the MLP step were second splits.)
    sl_ctx::RowPlanes *ra = nullptr, *rb = nullptr;
    const bool scope = ctx->plane_scope && !env_int("SLICED_GEMM_NO_ROWPLANE_REUSE", 0);
    if (scope && !a_kc && !b_kc) {
        ra = rowplanes_find(ctx, a, k, m);
        if (!ra) rb = rowplanes_find(ctx, b, k, n);
    }
    auto kmajor = [&](const float* src, size_t mn, __half*& hi, __half*& lo, float* inv_scratch, float* sc_scratch, const float** inv_use) -> int {
        if (!scope) return prep16_operand(ctx, src, mn, k, true, hi, lo, sc_scratch, inv_scratch, inv_use, nullptr);
        sl_ctx::RowPlanes* e = rowplanes_find(ctx, src, mn, k);
        if (!e) {
            int r = rowplanes_new(ctx, src, mn, k, &e);
            if (r != SL_OK) return r;
            const float* dummy = nullptr;
            if ((r = prep16_operand(ctx, src, mn, k, true, (__half*)e->hi, (__half*)e->lo, sc_scratch, e->inv, &dummy, nullptr)) != SL_OK) return r;
        }
        hi = (__half*)e->hi; lo = (__half*)e->lo; *inv_use = e->inv;
        return SL_OK;
    };
    if (ra) {          // A's planes exist (row-scaled over k): B is split with A's inverse row scales folded in
        a_hi = (__half*)ra->hi; a_lo = (__half*)ra->lo; a_inv_use = nullptr;
        if ((rc = prep16_operand(ctx, b, n, k, false, b_hi, b_lo, b_sc, b_inv, &b_inv_use, b_colsum_acc, ra->inv)) != SL_OK) return rc;
    } else if (rb) {   // the mirror image
        b_hi = (__half*)rb->hi; b_lo = (__half*)rb->lo; b_inv_use = nullptr;
        if ((rc = prep16_operand(ctx, a, m, k, false, a_hi, a_lo, a_sc, a_inv, &a_inv_use, nullptr, rb->inv)) != SL_OK) return rc;
        if (b_colsum_acc) {   // (the bias gradient normally rides B's column-maxima pass, which did not run)
            if ((rc = sl_add_row_mut_grad(ctx, SL_F32, k, n, b_colsum_acc, b)) != SL_OK) return rc;
        }
    } else {
        if (a_kc) { if ((rc = kmajor(a, m, a_hi, a_lo, a_inv, a_sc, &a_inv_use)) != SL_OK) return rc; }
        else if ((rc = prep16_operand(ctx, a, m, k, false, a_hi, a_lo, a_sc, a_inv, &a_inv_use, nullptr)) != SL_OK) return rc;
        if (b_kc) { if ((rc = kmajor(b, n, b_hi, b_lo, b_inv, b_sc, &b_inv_use)) != SL_OK) return rc; }
        else if ((rc = prep16_operand(ctx, b, n, k, false, b_hi, b_lo, b_sc, b_inv, &b_inv_use, b_colsum_acc)) != SL_OK) return rc;
    }
    if (m_chunks > 1 && !a_kc && !epi.any() && m % ((size_t)256 * m_chunks) == 0 && sl_gemm_pick_cfg(ctx, m / m_chunks, n) == 4) {
        const size_t mc = m / m_chunks;
        for (int ci = 0; ci < m_chunks; ++ci) {
            const size_t m0 = (size_t)ci * mc;
            rc = sl_gemm_tc_planes(ctx, (int)mc, (int)n, (int)k, a_hi + m0, a_lo + m0, m, b_hi, b_lo, b_kc ? k : n, c + m0 * n, nullptr, accumulate, 0,
                                   nullptr, nullptr, 1, b_kc ? 0 : 1, 1, a_inv_use ? a_inv_use + m0 : nullptr, b_inv_use);
            if (rc != SL_OK) return rc;
            if (exchange_chunks && (rc = sl_allreduce_sum_async(ctx, SL_F32, c + m0 * n, mc * n)) != SL_OK) return rc;
        }
        return SL_OK;
    }
    rc = sl_gemm_tc_planes(ctx, (int)m, (int)n, (int)k, a_hi, a_lo, a_kc ? k : m, b_hi, b_lo, b_kc ? k : n, c, epi.bias, accumulate, epi.relu, epi.c2,
                           epi.mask_src, a_kc ? 0 : 1, b_kc ? 0 : 1, 1, a_inv_use, b_inv_use, epi.bits_in, epi.bits_out);
    if (rc == SL_OK && exchange_chunks) rc = sl_allreduce_sum_async(ctx, SL_F32, c, m * n);
    return rc;
}

static bool f16x3_eligible(sl_ctx* ctx, int trans_a, int trans_b, size_t m, size_t n, size_t k, const void* a, const void* b) {
    return sl_gemm_pick_cfg(ctx, m, n) == 4 && ((trans_a ? m : k) % 8 == 0) && ((trans_b ? k : n) % 8 == 0) && sl_aligned16(a) && sl_aligned16(b);
}

static int gemm_ex_impl(sl_ctx* ctx, int dtype, int trans_a, int trans_b, size_t m, size_t n, size_t k, const void* a, const void* b, void* c,
                        int accumulate, int mode, const Epi& epi) {
    const float* bias = epi.bias;
    const int relu = epi.relu;
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (m == 0 || n == 0) return SL_OK;
    SL_REQUIRE(ctx, c != nullptr, "NULL output");
    if (k == 0) return accumulate ? SL_OK : sl_clear(ctx, c, m * n * sl_dtype_size(dtype));
    SL_REQUIRE(ctx, a && b, "NULL operand");
    if (mode < 0) mode = ctx->gemm_mode;
    // a gemm that overwrites a buffer whose planes are cached makes them stale
    sl_note_writes(ctx, c, epi.c2);
    // a few output tiles over a deep contraction (weight gradients of small layers): one tensor-core tile per SM would leave the
    // machine idle; the split-K CUDA-core kernel (gemm_skinny.cu smallout_tn_kernel) spreads K over all SMs instead
    const bool small_tn = dtype == SL_F32 && trans_a && !trans_b && k >= 256 && ((m + 63) / 64) * ((n + 63) / 64) <= (size_t)ctx->num_sms / 2 &&
                          !env_int("SLICED_GEMM_TC_FORCE", 0);
    if (dtype != SL_F32 || mode == SL_GEMM_SIMT || !tc_eligible(m, n, k) || small_tn) {
        if (epi.any() && dtype != SL_F32) return sl_set_error(ctx, SL_ERR_UNSUPPORTED, "fused epilogue is f32 only");
        int rc = 1;
        int mask_done = 0;
        if (dtype == SL_F32 && mode != SL_GEMM_SIMT)  // HBM-bound skinny shapes (n <= 16 or k <= 16) have their own kernels
            rc = sl_gemm_skinny_f32(ctx, trans_a, trans_b, m, n, k, (const float*)a, (const float*)b, (float*)c, accumulate, epi.mask_src, &mask_done,
                                    epi.bits_in);
        if (rc > 0) rc = sl_gemm_simt(ctx, dtype, trans_a, trans_b, m, n, k, a, b, c, accumulate);
        if (rc != SL_OK || !epi.any()) return rc;
        if (epi.bits_out || (epi.bits_in && !mask_done))
            return sl_set_error(ctx, SL_ERR_INVALID_ARG, "mask bits reached a kernel that cannot fuse them (internal: sl_linear_*_bits pick the path)");
        if (mask_done && !bias && !relu && !epi.c2) return SL_OK;   // the skinny kernel already applied the only epilogue term
        const size_t total = m * n;
        const size_t cap = (size_t)ctx->num_sms * 8;
        size_t blocks = (total + 255) / 256;
        SL_LAUNCH(ctx, epilogue_pass_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, total, n, (float*)c, bias, relu, epi.c2, mask_done ? nullptr : epi.mask_src);
        return SL_OK;
    }
    // 3xFP16: needs the 2-CTA kernel and 16-byte aligned fp16 rows in whatever layout the operand already has; anything else
    // takes the 3xTF32 path below (same accuracy class, half the tensor-pipe rate).
    if (mode == SL_GEMM_3XF16) {
        const bool ok = f16x3_eligible(ctx, trans_a, trans_b, m, n, k, a, b);
        if (ok) return gemm_f16x3(ctx, trans_a, trans_b, m, n, k, (const float*)a, (const float*)b, (float*)c, accumulate, epi);
        if (env_int("SLICED_GEMM_F16_STRICT", 0))   // tests: make the silent 3xTF32 substitution visible
            return sl_set_error(ctx, SL_ERR_UNSUPPORTED, "3xFP16 path cannot take %zux%zux%zu (trans %d,%d)", m, n, k, trans_a, trans_b);
        mode = SL_GEMM_3XTF32;
    }
    const bool three = mode == SL_GEMM_3XTF32;
    // Operand forms.  K-major = contraction index contiguous (A: !trans_a, B: trans_b).  The 2-CTA kernel also consumes MN-major
    // operands directly (UMMA MN-major descriptors), so a non-K-contiguous operand is only SPLIT in place there; the single-CTA
    // kernels want K-major tiles and get a transposing split.
    const bool mn_ok = sl_gemm_pick_cfg(ctx, m, n) == 4 && !env_int("SLICED_GEMM_NO_MN", 0);
    const bool a_kc = !trans_a, b_kc = trans_b != 0;
    const bool a_mn = !a_kc && mn_ok && (m % 4 == 0), b_mn = !b_kc && mn_ok && (n % 4 == 0);
    // leading dimensions of the planes: K-major [rows x ldk], MN-major [k x ldm]
    const size_t ldk = (k + 3) & ~size_t(3);
    const size_t a_ld = a_mn ? m : ldk, b_ld = b_mn ? n : ldk;
    const size_t a_elems = a_mn ? k * m : m * ldk, b_elems = b_mn ? k * n : n * ldk;
    const bool a_direct = !three && (a_kc ? (k % 4 == 0) : a_mn) && sl_aligned16(a);
    const bool b_direct = !three && (b_kc ? (k % 4 == 0) : b_mn) && sl_aligned16(b);
    const size_t a_plane = ((a_elems * 4) + 255) & ~size_t(255);
    const size_t b_plane = ((b_elems * 4) + 255) & ~size_t(255);
    // flat split (planes laid out exactly like the source buffer) -> eligible for reuse inside a scope
    const bool a_cacheable = ctx->plane_scope && !a_direct && (a_mn || (a_kc && ldk == k));
    const bool b_cacheable = ctx->plane_scope && !b_direct && (b_mn || (b_kc && ldk == k));
    const size_t need = ((a_direct || a_cacheable) ? 0 : a_plane * (three ? 2 : 1)) + ((b_direct || b_cacheable) ? 0 : b_plane * (three ? 2 : 1));
    char* ws = nullptr;
    if (need) {
        int rc = sl_ws2_reserve(ctx, need, (void**)&ws);
        if (rc != SL_OK) return rc;
    }
    const float *a_hi = (const float*)a, *a_lo = nullptr, *b_hi = (const float*)b, *b_lo = nullptr;
    size_t lda = a_kc ? k : m, ldb = b_kc ? k : n;
    char* cur = ws;
    if (!a_direct) {
        float *h, *l = nullptr;
        bool hit = false;
        if (a_cacheable) {
            int rc = plane_cache_get(ctx, (const float*)a, m * k, three, &h, &l, &hit);
            if (rc != SL_OK) return rc;
        } else {
            h = (float*)cur; cur += a_plane;
            if (three) { l = (float*)cur; cur += a_plane; }
        }
        // MN-major: the source [k x m] is split element-wise as a [k x m] matrix; K-major: split (a_kc) or transpose-split
        if (!hit) {
            int rc = a_mn ? sl_gemm_prep_operand(ctx, (const float*)a, k, m, 1, h, l, a_ld) : sl_gemm_prep_operand(ctx, (const float*)a, m, k, a_kc, h, l, a_ld);
            if (rc != SL_OK) return rc;
        }
        a_hi = h; a_lo = l; lda = a_ld;
    }
    if (!b_direct) {
        float *h, *l = nullptr;
        bool hit = false;
        if (b_cacheable) {
            int rc = plane_cache_get(ctx, (const float*)b, n * k, three, &h, &l, &hit);
            if (rc != SL_OK) return rc;
        } else {
            h = (float*)cur; cur += b_plane;
            if (three) { l = (float*)cur; cur += b_plane; }
        }
        if (!hit) {
            int rc = b_mn ? sl_gemm_prep_operand(ctx, (const float*)b, k, n, 1, h, l, b_ld) : sl_gemm_prep_operand(ctx, (const float*)b, n, k, b_kc, h, l, b_ld);
            if (rc != SL_OK) return rc;
        }
        b_hi = h; b_lo = l; ldb = b_ld;
    }
    return sl_gemm_tc_planes(ctx, (int)m, (int)n, (int)k, a_hi, a_lo, lda, b_hi, b_lo, ldb, (float*)c, bias, accumulate, relu, epi.c2, epi.mask_src,
                             a_mn ? 1 : 0, b_mn ? 1 : 0, 0, nullptr, nullptr, epi.bits_in, epi.bits_out);
}

// does gemm_ex_impl route this f32 product to the 2-CTA tensor-core kernel with a vectorised epilogue (the only sink that fuses mask bits)?
static bool bits_fused_path(sl_ctx* ctx, int mode, size_t m, size_t n, size_t k, const void* c, const void* bias, const void* bits) {
    if (mode < 0) mode = ctx->gemm_mode;
    return mode != SL_GEMM_SIMT && tc_eligible(m, n, k) && sl_gemm_pick_cfg(ctx, m, n) == 4 && n % 32 == 0 && sl_aligned16(c) &&
           (!bias || sl_aligned16(bias)) && bits && (reinterpret_cast<uintptr_t>(bits) & 3u) == 0;
}

extern "C" {

int sl_gemm_ex(sl_ctx* ctx, int dtype, int trans_a, int trans_b, size_t m, size_t n, size_t k, const void* a, const void* b, void* c,
               int accumulate, int mode) {
    return gemm_ex_impl(ctx, dtype, trans_a, trans_b, m, n, k, a, b, c, accumulate, mode, Epi{});
}

int sl_gemm(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* rhs, void* out, int mode) {
    return gemm_ex_impl(ctx, dtype, 0, 0, m, n, k, lhs, rhs, out, 0, mode, Epi{});
}

int sl_gemm_nt(sl_ctx* ctx, int dtype, size_t m, size_t n, size_t k, const void* a, const void* b, void* c, int mode) {
    return gemm_ex_impl(ctx, dtype, 0, 1, m, n, k, a, b, c, 0, mode, Epi{});
}

int sl_gemm_tn(sl_ctx* ctx, int dtype, size_t m, size_t n, size_t k, const void* a, const void* b, void* c, int mode) {
    return gemm_ex_impl(ctx, dtype, 1, 0, m, n, k, a, b, c, 0, mode, Epi{});
}

int sl_linear_fwd(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* rhs, const void* bias, void* z_out,
                  void* act_out, int mode) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_REQUIRE(ctx, dtype == SL_F32, "sl_linear_fwd is f32 only");
    SL_REQUIRE(ctx, z_out != nullptr, "z_out is NULL");
    Epi e;
    e.bias = (const float*)bias;
    e.c2 = (float*)act_out;
    return gemm_ex_impl(ctx, dtype, 0, 0, m, n, k, lhs, rhs, z_out, 0, mode, e);
}

int sl_linear_bwd_input_relu(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* rhs, const void* out_grad, const void* z_prev,
                             void* x_grad, int mode) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_REQUIRE(ctx, dtype == SL_F32, "sl_linear_bwd_input_relu is f32 only");
    SL_REQUIRE(ctx, x_grad != nullptr, "x_grad is NULL");
    Epi e;
    e.mask_src = (const float*)z_prev;
    // gemmT(m, k, n, out_grad, rhs, x_grad) (gemm/grad/cpu_stack.rs:36) with the relu gradient (src/matrix.rs:186) in the epilogue
    return gemm_ex_impl(ctx, dtype, 0, 1, m, k, n, out_grad, rhs, x_grad, 0, mode, e);
}

int sl_linear_fwd_bits(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* rhs, const void* bias, void* act_out,
                       uint32_t* mask_bits, int mode) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_REQUIRE(ctx, dtype == SL_F32, "sl_linear_fwd_bits is f32 only");
    SL_REQUIRE(ctx, act_out != nullptr && mask_bits != nullptr, "NULL output");
    if (m == 0 || n == 0) return SL_OK;
    sl_note_writes(ctx, act_out, mask_bits);
    Epi e;
    e.bias = (const float*)bias;
    if (k > 0 && bits_fused_path(ctx, mode, m, n, k, act_out, bias, mask_bits)) {
        e.relu = 1;
        e.bits_out = mask_bits;
        return gemm_ex_impl(ctx, dtype, 0, 0, m, n, k, lhs, rhs, act_out, 0, mode, e);
    }
    // any other shape: z = lhs * rhs + bias into act_out, then one pass takes the bits and applies the relu in place
    int rc = gemm_ex_impl(ctx, dtype, 0, 0, m, n, k, lhs, rhs, act_out, 0, mode, e);
    if (rc != SL_OK) return rc;
    if (k == 0 && bias) {   // (gemm of an empty contraction clears the output and applies no epilogue)
        const size_t total = m * n, cap = (size_t)ctx->num_sms * 8, blocks = (total + 255) / 256;
        SL_LAUNCH(ctx, epilogue_pass_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, total, n, (float*)act_out, (const float*)bias, 0, (float*)nullptr,
                  (const float*)nullptr);
    }
    const size_t nw = (n + 31) / 32, words = m * nw, cap = (size_t)ctx->num_sms * 8, blocks = (words + 255) / 256;
    SL_LAUNCH(ctx, bits_relu_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, m, n, nw, (float*)act_out, mask_bits);
    return SL_OK;
}

int sl_linear_bwd_input_relu_bits(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* rhs, const void* out_grad, const uint32_t* mask_bits,
                                  void* x_grad, int mode) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_REQUIRE(ctx, dtype == SL_F32, "sl_linear_bwd_input_relu_bits is f32 only");
    SL_REQUIRE(ctx, x_grad != nullptr && mask_bits != nullptr, "NULL argument");
    if (m == 0 || k == 0) return SL_OK;
    sl_note_writes(ctx, x_grad);
    Epi e;
    // x_grad[m x k] = out_grad[m x n] * rhs[k x n]^T: the product's output extent is k, its contraction n
    const bool al4 = (reinterpret_cast<uintptr_t>(mask_bits) & 3u) == 0;
    const bool skinny = n >= 1 && n <= 16 && k % 32 == 0 && al4 && (mode < 0 ? ctx->gemm_mode : mode) != SL_GEMM_SIMT;   // the kernel decides; checked below
    if (n > 0 && bits_fused_path(ctx, mode, m, k, n, x_grad, nullptr, mask_bits)) {
        e.bits_in = mask_bits;
        return gemm_ex_impl(ctx, dtype, 0, 1, m, k, n, out_grad, rhs, x_grad, 0, mode, e);
    }
    if (skinny && !tc_eligible(m, k, n)) {
        int done = 0;
        int rc = sl_gemm_skinny_f32(ctx, 0, 1, m, k, n, (const float*)out_grad, (const float*)rhs, (float*)x_grad, 0, nullptr, &done, mask_bits);
        if (rc < 0 || (rc == SL_OK && done)) return rc;
        if (rc > 0) rc = gemm_ex_impl(ctx, dtype, 0, 1, m, k, n, out_grad, rhs, x_grad, 0, mode, e);   // not a skinny shape after all
        if (rc != SL_OK) return rc;
    } else {
        int rc = gemm_ex_impl(ctx, dtype, 0, 1, m, k, n, out_grad, rhs, x_grad, 0, mode, e);
        if (rc != SL_OK) return rc;
    }
    const size_t nw = (k + 31) / 32, total = m * k, cap = (size_t)ctx->num_sms * 8, blocks = (total + 255) / 256;
    SL_LAUNCH(ctx, bits_apply_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, m, k, nw, (float*)x_grad, mask_bits);
    return SL_OK;
}

int sl_add_row_mut_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* rhs_grad, const void* out_grad);

static int linear_bwd_params_impl(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* out_grad, void* w_grad,
                                  void* b_grad, int mode, int chunks, bool exchange);

int sl_linear_bwd_params(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* out_grad, void* w_grad, void* b_grad,
                         int mode) {
    return linear_bwd_params_impl(ctx, dtype, m, k, n, lhs, out_grad, w_grad, b_grad, mode, 1, false);
}

int sl_linear_bwd_params_exchange(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* out_grad, void* w_grad,
                                  void* b_grad, int chunks, int mode) {
    return linear_bwd_params_impl(ctx, dtype, m, k, n, lhs, out_grad, w_grad, b_grad, mode, chunks < 1 ? 1 : chunks, true);
}

static int linear_bwd_params_impl(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* out_grad, void* w_grad,
                                  void* b_grad, int mode, int chunks, bool exchange) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (m == 0 || k == 0 || n == 0) return SL_OK;
    SL_REQUIRE(ctx, lhs && out_grad && w_grad, "NULL argument");
    if (mode < 0) mode = ctx->gemm_mode;
    // the weight-gradient gemm contracts over the batch: out_grad is its MN-major B operand, whose column maxima need one full
    // pass over out_grad anyway — that pass also produces the bias gradient's column sums
    if (b_grad && dtype == SL_F32 && mode == SL_GEMM_3XF16 && tc_eligible(k, n, m) && f16x3_eligible(ctx, 1, 0, k, n, m, lhs, out_grad)) {
        sl_note_writes(ctx, w_grad, b_grad);
        int rc = gemm_f16x3(ctx, 1, 0, k, n, m, (const float*)lhs, (const float*)out_grad, (float*)w_grad, 0, Epi{}, (float*)b_grad, chunks, exchange);
        if (rc == SL_OK && exchange) rc = sl_allreduce_sum_async(ctx, dtype, b_grad, n);
        return rc;
    }
    if (b_grad) {
        int rc = sl_add_row_mut_grad(ctx, dtype, m, n, b_grad, out_grad);
        if (rc != SL_OK) return rc;
    }
    int rc = gemm_ex_impl(ctx, dtype, 1, 0, k, n, m, lhs, out_grad, w_grad, 0, mode, Epi{});
    if (rc == SL_OK && exchange) {
        rc = sl_allreduce_sum_async(ctx, dtype, w_grad, k * n);
        if (rc == SL_OK && b_grad) rc = sl_allreduce_sum_async(ctx, dtype, b_grad, n);
    }
    return rc;
}

int sl_gemm_scope_begin(sl_ctx* ctx) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    for (auto& e : ctx->plane_cache) e.valid = false;
    ctx->plane_cursor = 0;
    for (auto& e : ctx->rowplane_cache) e.valid = false;
    ctx->rowplane_cursor = 0;
    ctx->plane_scope = true;
    return SL_OK;
}

int sl_gemm_scope_end(sl_ctx* ctx) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    for (auto& e : ctx->plane_cache) e.valid = false;
    for (auto& e : ctx->rowplane_cache) e.valid = false;
    ctx->plane_scope = false;
    return SL_OK;
}

int sl_gemm_grad(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* rhs, void* lhs_grad, void* rhs_grad,
                 const void* out_grad, int accumulate, int mode) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_REQUIRE(ctx, out_grad != nullptr || m == 0 || n == 0, "NULL out_grad");
    // both halves read out_grad: split it once (implicit scope unless the caller already opened one)
    struct ScopeGuard {
        sl_ctx* c; bool own;
        ScopeGuard(sl_ctx* ctx, bool want) : c(ctx), own(want && !ctx->plane_scope) { if (own) sl_gemm_scope_begin(c); }
        ~ScopeGuard() { if (own) sl_gemm_scope_end(c); }
    } guard(ctx, lhs_grad && rhs_grad && dtype == SL_F32);
    if (lhs_grad) {  // gemmT(m, k, n, out_grad, rhs, lhs_grad): lhs_grad[m x k] = out_grad[m x n] * rhs[k x n]^T
        int rc = gemm_ex_impl(ctx, dtype, 0, 1, m, k, n, out_grad, rhs, lhs_grad, accumulate, mode, Epi{});
        if (rc != SL_OK) return rc;
    }
    if (rhs_grad) {  // Tgemm(k, n, m, lhs, out_grad, rhs_grad): rhs_grad[k x n] = lhs[m x k]^T * out_grad[m x n]
        int rc = gemm_ex_impl(ctx, dtype, 1, 0, k, n, m, lhs, out_grad, rhs_grad, accumulate, mode, Epi{});
        if (rc != SL_OK) return rc;
    }
    return SL_OK;
}

}  // extern "C"
