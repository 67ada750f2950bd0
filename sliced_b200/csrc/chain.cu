// chain.cu — sl_fused_chain: a micro-op interpreter for chains of element-wise ops (the general form of the hand-fused
// sl_chained_fwd / sl_chained_bwd; SURVEY 8f row 2, north_star "fuse chained ops from the tape where the graph allows").
//
// One generic, HBM-bound kernel: every thread owns one 128-bit pack per input, the program (<= 32 instructions, passed by value:
// it sits in the constant bank, decoding is warp-uniform) runs once per pack.  The virtual registers live in SHARED memory,
// laid out [register][thread] as 16-byte packs (conflict-free LDS.128 / STS.128) — indexing a register file held in real
// registers with run-time indices would force it into local memory anyway.  Two host-computed hints per instruction keep the
// shared-memory traffic near one access per micro-op on linear chains: operand forwarding (an operand that is the previous
// instruction's result is taken from the live register) and dead-store elimination (a result that is only ever consumed by the
// next instruction, and is no output, is never written back).
// Shared memory per block = n_regs x 256 threads x 16 B, so short programs keep 8 blocks resident per SM.
#include "common.cuh"
#include "ew_math.cuh"

namespace {

constexpr int CH_THREADS = 256;
constexpr int FWD_A = 1, FWD_B = 2, NO_STORE = 4;

struct ChainArgs {
    const void* in[SL_CHAIN_MAX_INPUTS];
    void* out[SL_CHAIN_MAX_OUTPUTS];
    uint32_t in_coherent;   // bit i: input i aliases an output -> plain (coherent) loads instead of the read-only path
};

template <typename T, int V>
struct Vec {
    T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ch_load(const T* p, bool coherent) {
    Vec<T, V> r;
    if (V == 1) {
        r.v[0] = coherent ? *p : __ldg(p);
    } else {
        const Pack<T> q = coherent ? ld_pack(p) : ld_stream(p);
#pragma unroll
        for (int e = 0; e < V; ++e) r.v[e] = q.v[e];
    }
    return r;
}
template <typename T, int V>
__device__ __forceinline__ void ch_store(T* p, const Vec<T, V>& x) {
    if (V == 1) {
        *p = x.v[0];
    } else {
        Pack<T> q;
#pragma unroll
        for (int e = 0; e < V; ++e) q.v[e] = x.v[e];
        st_pack(p, q);
    }
}

#define CH_UNARY_CASES(KIND, FN)                                                                                            \
    case KIND + SL_UN_SQUARE: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_SQUARE>(a.v[e], p0, p1); break;   \
    case KIND + SL_UN_POW: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_POW>(a.v[e], p0, p1); break;         \
    case KIND + SL_UN_RELU: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_RELU>(a.v[e], p0, p1); break;       \
    case KIND + SL_UN_TANH: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_TANH>(a.v[e], p0, p1); break;       \
    case KIND + SL_UN_SIGMOID: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_SIGMOID>(a.v[e], p0, p1); break; \
    case KIND + SL_UN_EXP: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_EXP>(a.v[e], p0, p1); break;         \
    case KIND + SL_UN_LN: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_LN>(a.v[e], p0, p1); break;           \
    case KIND + SL_UN_NEG_LN: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_NEG_LN>(a.v[e], p0, p1); break;   \
    case KIND + SL_UN_CLIP: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_CLIP>(a.v[e], p0, p1); break;       \
    case KIND + SL_UN_NEG: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_NEG>(a.v[e], p0, p1); break;         \
    case KIND + SL_UN_MUL_SCALAR: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_MUL_SCALAR>(a.v[e], p0, p1); break; \
    case KIND + SL_UN_NEG_DIV_SCALAR: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_NEG_DIV_SCALAR>(a.v[e], p0, p1); break; \
    case KIND + SL_UN_ADD_SCALAR: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = FN<SL_UN_ADD_SCALAR>(a.v[e], p0, p1); break;

// V elements per thread per iteration (V = pack width for the aligned body, 1 for tails / misaligned pointers)
template <typename T, int V>
__global__ void __launch_bounds__(CH_THREADS) chain_kernel(const sl_chain_prog prog, const ChainArgs args, size_t begin, size_t count) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    Vec<T, V>* regs = reinterpret_cast<Vec<T, V>*>(ch_smem);   // [reg][thread]
    const int t = threadIdx.x;
    auto R = [&](int r) -> Vec<T, V>& { return regs[r * CH_THREADS + t]; };
    for (size_t i = (size_t)blockIdx.x * CH_THREADS + t; i < count; i += (size_t)gridDim.x * CH_THREADS) {
        const size_t off = begin + i * V;
#pragma unroll 1
        for (int r = 0; r < prog.n_in; ++r) R(r) = ch_load<T, V>((const T*)args.in[r] + off, (args.in_coherent >> r) & 1u);
        Vec<T, V> d;   // the previous instruction's result (operand forwarding)
#pragma unroll 1
        for (int k = 0; k < prog.n_instr; ++k) {
            const sl_chain_instr ins = prog.instr[k];
            const Vec<T, V> a = (ins.flags & FWD_A) ? d : R(ins.a);
            Vec<T, V> b = a;
            if (ins.op <= SL_CH_DIV) b = (ins.flags & FWD_B) ? d : R(ins.b);
            const T p0 = (T)ins.imm0, p1 = (T)ins.imm1;
            switch (ins.op) {
            case SL_CH_ADD: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = a.v[e] + b.v[e]; break;
            case SL_CH_SUB: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = a.v[e] - b.v[e]; break;
            case SL_CH_MUL: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = a.v[e] * b.v[e]; break;
            case SL_CH_DIV: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = a.v[e] / b.v[e]; break;
            case SL_CH_RDIV_IMM: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = p0 / a.v[e]; break;
            case SL_CH_CONST: _Pragma("unroll") for (int e = 0; e < V; ++e) d.v[e] = p0; break;
            case SL_CH_COPY: d = a; break;
            CH_UNARY_CASES(SL_CH_UNARY_F, unary_f)
            CH_UNARY_CASES(SL_CH_UNARY_D, unary_d)
            default: break;
            }
            if (!(ins.flags & NO_STORE)) R(ins.dst) = d;
        }
#pragma unroll 1
        for (int j = 0; j < prog.n_out; ++j) {
            Vec<T, V> v = R(prog.out_reg[j]);
            T* dst = (T*)args.out[j] + off;
            if (prog.out_acc[j]) {
                const Vec<T, V> o = ch_load<T, V>(dst, true);
#pragma unroll
                for (int e = 0; e < V; ++e) v.v[e] = o.v[e] + v.v[e];
            }
            ch_store<T, V>(dst, v);
        }
    }
}

template <typename T>
int chain_t(sl_ctx* ctx, const sl_chain_prog& prog, const ChainArgs& args, size_t n, bool aligned) {
    constexpr int V = Pack<T>::N;
    const size_t smem_v = (size_t)prog.n_regs * CH_THREADS * sizeof(Vec<T, V>);
    const size_t smem_1 = (size_t)prog.n_regs * CH_THREADS * sizeof(Vec<T, 1>);
    const size_t npacks = aligned ? n / V : 0;
    if (npacks) {
        auto kern = chain_kernel<T, V>;
        if (smem_v > 48 * 1024) SL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v));
        size_t per_sm = (size_t)(220 * 1024) / (smem_v ? smem_v : 1);
        per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
        const size_t blocks = (npacks + CH_THREADS - 1) / CH_THREADS, cap = (size_t)ctx->num_sms * per_sm;
        SL_LAUNCH(ctx, kern, (unsigned)(blocks < cap ? blocks : cap), CH_THREADS, smem_v, prog, args, (size_t)0, npacks);
    }
    const size_t done = npacks * V;
    if (done < n) {
        auto kern = chain_kernel<T, 1>;
        if (smem_1 > 48 * 1024) SL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_1));
        const size_t rem = n - done, blocks = (rem + CH_THREADS - 1) / CH_THREADS, cap = (size_t)ctx->num_sms * 8;
        SL_LAUNCH(ctx, kern, (unsigned)(blocks < cap ? blocks : cap), CH_THREADS, smem_1, prog, args, done, rem);
    }
    return SL_OK;
}

bool float_only(int op) {
    const int u = op >= SL_CH_UNARY_D ? op - SL_CH_UNARY_D : op - SL_CH_UNARY_F;
    return op >= SL_CH_UNARY_F &&
           (u == SL_UN_POW || u == SL_UN_TANH || u == SL_UN_SIGMOID || u == SL_UN_EXP || u == SL_UN_LN || u == SL_UN_NEG_LN);
}

}  // namespace

extern "C" int sl_fused_chain(sl_ctx* ctx, int dtype, const sl_chain_prog* prog_in, const void* const* inputs, void* const* outputs, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_REQUIRE(ctx, prog_in && outputs, "NULL argument");
    sl_chain_prog prog = *prog_in;
    SL_REQUIRE(ctx, prog.n_instr >= 0 && prog.n_instr <= SL_CHAIN_MAX_INSTRS, "bad n_instr");
    SL_REQUIRE(ctx, prog.n_in >= 0 && prog.n_in <= SL_CHAIN_MAX_INPUTS && (prog.n_in == 0 || inputs), "bad n_in");
    SL_REQUIRE(ctx, prog.n_out >= 1 && prog.n_out <= SL_CHAIN_MAX_OUTPUTS, "bad n_out");
    SL_REQUIRE(ctx, prog.n_regs >= prog.n_in && prog.n_regs >= 1 && prog.n_regs <= SL_CHAIN_MAX_REGS, "bad n_regs");
    // validate: every register is written before it is read, codes in range, integer programs stay integer
    bool defined[SL_CHAIN_MAX_REGS] = {};
    for (int r = 0; r < prog.n_in; ++r) defined[r] = true;
    for (int k = 0; k < prog.n_instr; ++k) {
        sl_chain_instr& in = prog.instr[k];
        const int op = in.op;
        const bool binary = op <= SL_CH_DIV;
        const bool unary_f = op >= SL_CH_UNARY_F && op < SL_CH_UNARY_F + SL_UN_COUNT_;
        const bool unary_d = op >= SL_CH_UNARY_D && op < SL_CH_UNARY_D + SL_UN_COUNT_;
        SL_REQUIRE(ctx, binary || op == SL_CH_RDIV_IMM || op == SL_CH_CONST || op == SL_CH_COPY || unary_f || unary_d, "bad opcode");
        SL_REQUIRE(ctx, in.dst < prog.n_regs, "dst out of range");
        if (op != SL_CH_CONST) SL_REQUIRE(ctx, in.a < prog.n_regs && defined[in.a], "operand a read before written");
        if (binary) SL_REQUIRE(ctx, in.b < prog.n_regs && defined[in.b], "operand b read before written");
        if (dtype == SL_I32 && float_only(op)) return sl_set_error(ctx, SL_ERR_UNSUPPORTED, "sl_fused_chain: opcode %d needs a float dtype", op);
        defined[in.dst] = true;
        in.flags = 0;
    }
    for (int j = 0; j < prog.n_out; ++j) SL_REQUIRE(ctx, prog.out_reg[j] < prog.n_regs && defined[prog.out_reg[j]], "output register never written");
    // hints: operand forwarding from the previous instruction, and results nobody reads back from shared memory
    for (int k = 0; k < prog.n_instr; ++k) {
        sl_chain_instr& in = prog.instr[k];
        if (k > 0) {
            const int pd = prog.instr[k - 1].dst;
            if (in.op != SL_CH_CONST && in.a == pd) in.flags |= FWD_A;
            if (in.op <= SL_CH_DIV && in.b == pd) in.flags |= FWD_B;
        }
        bool read_later = false;
        for (int j = 0; j < prog.n_out; ++j) read_later |= prog.out_reg[j] == in.dst;   // outputs are read from shared memory at the end
        for (int q = k + 1; q < prog.n_instr && !read_later; ++q) {
            const sl_chain_instr& nx = prog.instr[q];
            const bool uses = (nx.op != SL_CH_CONST && nx.a == in.dst) || (nx.op <= SL_CH_DIV && nx.b == in.dst);
            if (uses && q > k + 1) read_later = true;   // the next instruction takes it from the live register
            if (nx.dst == in.dst) break;                // overwritten: later reads see the new value
        }
        if (!read_later) in.flags |= NO_STORE;
    }
    // an output register overwritten after the instruction that produced the wanted value is the caller's bug; the value stored is
    // whatever the register holds at the end (documented: outputs read the FINAL register contents)
    if (n == 0) return SL_OK;
    ChainArgs args{};
    bool aligned = true;
    for (int j = 0; j < prog.n_out; ++j) {
        SL_REQUIRE(ctx, outputs[j] != nullptr, "NULL output");
        args.out[j] = outputs[j];
        aligned = aligned && sl_aligned16(outputs[j]);
        sl_note_write(ctx, outputs[j]);
    }
    for (int r = 0; r < prog.n_in; ++r) {
        SL_REQUIRE(ctx, inputs[r] != nullptr, "NULL input");
        args.in[r] = inputs[r];
        aligned = aligned && sl_aligned16(inputs[r]);
        for (int j = 0; j < prog.n_out; ++j)
            if (inputs[r] == outputs[j]) args.in_coherent |= 1u << r;
    }
    SL_DISPATCH_DTYPE(ctx, dtype, T, return chain_t<T>(ctx, prog, args, n, aligned));
    return SL_OK;
}
