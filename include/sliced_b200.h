/*
 * sliced_b200.h — C ABI of libsliced_b200.so, the B200 (sm_100a) implementation of the
 * forward + backward op set of elftausend/sliced.
 *
 * This is the drop-in boundary: every entry point below is what a Rust `sliced-b200-sys`
 * FFI crate binds so that sliced's own op traits (src/ops2/<op>/mod.rs, grad.rs) can be
 * implemented `for CUDA<Mods>` in a new `src/ops2/<op>/cuda.rs` next to `cpu.rs` /
 * `opencl.rs` (hook: reference src/lib.rs:24-25).  Each function cites the reference
 * interface it replaces as `ref: file:line` (paths relative to the reference crate root).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  All matrix arguments are row-major,
 *    contiguous device pointers (the reference has no strides / leading dimensions).
 *  - Every function returns an sl_status (0 = ok, negative = error); nothing throws or
 *    aborts across the boundary.  sl_last_error_string(ctx) describes the last failure.
 *  - A sl_ctx owns one CUDA device + one stream.  All op calls are asynchronous on that
 *    stream and ordered; only sl_read / sl_sync / scalar *_host calls block.
 *  - "SET" entry points fully overwrite their output; "ACC" entry points `+=` into it,
 *    exactly as the reference CPU slice function they replace does.
 *  - dtype is one of sl_dtype.  f32 is supported everywhere; f64 and i32 are supported by
 *    the element-wise / broadcast / reduction / transpose families and by the CUDA-core
 *    gemm (the reference's unit tests are written in i32 / f64).  Unsupported
 *    combinations return SL_ERR_UNSUPPORTED — there is no CPU fallback on this path.
 */
#ifndef SLICED_B200_H
#define SLICED_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SL_ABI_VERSION 1

typedef struct sl_ctx sl_ctx;

typedef enum sl_status {
    SL_OK = 0,
    SL_ERR_INVALID_ARG = -1,
    SL_ERR_CUDA = -2,
    SL_ERR_UNSUPPORTED = -3,
    SL_ERR_NCCL = -4,
    SL_ERR_NO_DEVICE = -5
} sl_status;

typedef enum sl_dtype { SL_F32 = 0, SL_F64 = 1, SL_I32 = 2 } sl_dtype;

/* Binary element-wise operators. ref: src/ops2/binary_ew/mod.rs:67-99 (add/mul/div/sub). */
typedef enum sl_binop { SL_ADD = 0, SL_SUB = 1, SL_MUL = 2, SL_DIV = 3 } sl_binop;

/*
 * Unary functions reachable through custos `apply_fn` / `add_unary_grad` from sliced.
 * Forward f(x) and the derivative g(x) used by the *_grad entry point (x_grad += g(x)*out_grad)
 * are exactly the closures the reference registers:
 *   SQUARE        f = x*x              g = x*2                       ref: src/ops.rs:36,40
 *   POW(p0)       f = x^p0             g = x^(p0-1) * p0             ref: src/ops.rs:65,70-72
 *   RELU          f = (x>=0)*x         g = (x>=0)                    ref: src/matrix.rs:181,186
 *   TANH          f = tanh(x)          g = 1 - tanh(x)^2             ref: src/matrix.rs:218,223-225
 *   SIGMOID       f = 1/(1+exp(-x))    g = exp(-x)/(1+exp(-x))^2     ref: src/matrix.rs:246-250,255-257
 *   EXP           f = exp(x)           g = exp(x)                    ref: src/ops.rs:438 (grad: :403, commented)
 *   LN            f = ln(x)            g = 1/x
 *   NEG_LN        f = -ln(x)           g = -1/x                      ref: examples/nn.rs:137
 *   CLIP(p0,p1)   f = min(max(x,p0),p1) g = (p0<=x && x<=p1)         ref: src/ops.rs:423
 *   NEG           f = -x               g = -1
 *   MUL_SCALAR    f = x*p0             g = p0
 *   NEG_DIV_SCALAR f = (-x)/p0         g = -1/p0                     ref: examples/nn.rs:151
 *   ADD_SCALAR    f = x+p0             g = 1
 */
typedef enum sl_unop {
    SL_UN_SQUARE = 0,
    SL_UN_POW = 1,
    SL_UN_RELU = 2,
    SL_UN_TANH = 3,
    SL_UN_SIGMOID = 4,
    SL_UN_EXP = 5,
    SL_UN_LN = 6,
    SL_UN_NEG_LN = 7,
    SL_UN_CLIP = 8,
    SL_UN_NEG = 9,
    SL_UN_MUL_SCALAR = 10,
    SL_UN_NEG_DIV_SCALAR = 11,
    SL_UN_ADD_SCALAR = 12,
    SL_UN_COUNT_
} sl_unop;

/* gemm arithmetic mode (f32 only; f64/i32 always use the CUDA-core kernel). */
typedef enum sl_gemm_mode {
    SL_GEMM_3XTF32 = 0, /* tcgen05 kind::tf32, 3 MMAs per k-slice (hi*hi + hi*lo + lo*hi), fp32-class accuracy */
    SL_GEMM_TF32 = 1,   /* flagged fast mode: one tcgen05 kind::tf32 MMA per k-slice */
    SL_GEMM_SIMT = 2,   /* exact fp32 FMA accumulation on CUDA cores (any shape / dtype) */
    SL_GEMM_3XF16 = 3   /* default: tcgen05 kind::f16 on fp16 hi/lo planes of the operands after an exact power-of-two scaling per output
                           row / column: the same 22-bit operand split and 3 MMAs per k-slice as 3XTF32 at twice the MMA rate.
                           Shapes the fp16 path cannot take (small, or rows not a multiple of 8) silently use 3XTF32. */
} sl_gemm_mode;

/* ---------------------------------------------------------------- context / memory */

int sl_abi_version(void);
/* Number of visible CUDA devices (0 on a box without a GPU; never fails). */
int sl_device_count(void);

/* ref: custos `CUDA::<Mods>::new(idx)` † (device object; hook src/lib.rs:24-25). */
int sl_ctx_create(int device, sl_ctx** out_ctx);
/* Same, but borrows an existing cudaStream_t (e.g. torch's current stream) instead of creating one. */
int sl_ctx_create_on_stream(int device, void* cuda_stream, sl_ctx** out_ctx);
int sl_ctx_destroy(sl_ctx* ctx);
const char* sl_last_error_string(sl_ctx* ctx);
void* sl_ctx_stream(sl_ctx* ctx);
int sl_ctx_device(sl_ctx* ctx);
/* Default gemm mode used when an entry point is passed mode < 0. Also read from env SLICED_GEMM_MODE={3xf16,3xtf32,tf32,simt}.
 * The built-in default is SL_GEMM_3XF16. */
int sl_ctx_set_gemm_mode(sl_ctx* ctx, int mode);
/* Number of kernels this library launched on ctx since creation (bench.py's gpu_launches). */
uint64_t sl_ctx_launch_count(sl_ctx* ctx);

/* Per-launch CUDA-event timing of the tensor-core gemm kernel (bench.py's live roofline measurement): between begin and
 * end every MMA-kernel launch is bracketed by two events on the ctx stream; end synchronises and returns the number of
 * launches, their summed device time (ms) and their summed algorithmic flops (2*M*N*K each). */
int sl_ctx_profile_begin(sl_ctx* ctx);
int sl_ctx_profile_end(sl_ctx* ctx, uint64_t* n_launches, double* total_ms, double* total_flops);
/* In-situ per-kernel timing (tools/step_breakdown.py): the first call switches the context to timing EVERY launch between
 * sl_ctx_profile_begin and _end; called between them it writes "kernel,launches,total_ms" lines (sorted by time) into out. */
int sl_ctx_profile_report(sl_ctx* ctx, char* out, size_t cap);

/* ref: custos `Alloc<T>::alloc` / `OnDropBuffer` † */
int sl_malloc(sl_ctx* ctx, size_t bytes, void** out_dptr);
int sl_free(sl_ctx* ctx, void* dptr);
/* pinned host staging memory (ref: custos `Buffer::from((&device, slice))` † host source) */
int sl_host_alloc(sl_ctx* ctx, size_t bytes, void** out_hptr);
int sl_host_free(sl_ctx* ctx, void* hptr);
/* ref: custos `WriteBuf::write` † — host -> device, async on the ctx stream */
int sl_write(sl_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
/* Upload pipeline for a training loop that streams batches from (pinned) host memory: sl_write_prefetch copies on the context's
 * second (copy) stream so the next batch uploads while the current step computes; sl_prefetch_wait makes the compute stream wait for
 * every prefetch issued so far; sl_prefetch_release makes the copy stream wait for the compute issued so far (call it before
 * overwriting a staging buffer that earlier compute may still read). */
int sl_write_prefetch(sl_ctx* ctx, void* dst_dev, const void* src_host_pinned, size_t bytes);
int sl_prefetch_wait(sl_ctx* ctx);
int sl_prefetch_release(sl_ctx* ctx);
/* ref: custos `Read::read` † — device -> host, blocks until the data is on the host */
int sl_read(sl_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
/* Non-blocking variant: ordered on the ctx stream, destination must be pinned (sl_host_alloc); the bytes are valid on the host after
 * the next sl_sync / sl_read.  Lets a training loop fetch every step's loss without stalling the launch of the next step. */
int sl_read_async(sl_ctx* ctx, void* dst_host_pinned, const void* src_dev, size_t bytes);
/* ref: custos `CloneBuf` / `WriteBuf::write_buf` † — device -> device */
int sl_copy(sl_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);
/* ref: custos `ClearBuf::clear` † / `Gradients::zero_grad` (examples/nn.rs:186-188) */
int sl_clear(sl_ctx* ctx, void* dst_dev, size_t bytes);
int sl_fill(sl_ctx* ctx, int dtype, void* dst_dev, double value, size_t n);
int sl_sync(sl_ctx* ctx);

/* CUDA-graph capture / replay of a sequence of calls on ctx — the device analogue of custos' `Lazy` module building the op
 * graph once and replaying it with `run()` (ref: examples/sine_net.rs:178-233 `sine_net_lazy2`).  Between begin and end the
 * calls are recorded, not executed; blocking calls (sl_read, sl_sync, sl_malloc/sl_free) are not allowed, and every buffer /
 * scratch the sequence needs must already exist (run it once eagerly first).  sl_graph_launch replays on the ctx stream. */
int sl_graph_begin(sl_ctx* ctx);
int sl_graph_end(sl_ctx* ctx, void** out_graph_exec);
int sl_graph_launch(sl_ctx* ctx, void* graph_exec);
int sl_graph_destroy(sl_ctx* ctx, void* graph_exec);

/* ---------------------------------------------------------------- E: element-wise / broadcast */

/* out[i] = lhs[i] op rhs[i]  (SET).  ref: src/ops2/binary_ew/mod.rs:55-100, cpu_stack.rs:42-54 */
int sl_binary_ew(sl_ctx* ctx, int dtype, int binop, const void* lhs, const void* rhs, void* out, size_t n);
/* lhs_grad[i] += dl(l,r)*og[i]; rhs_grad[i] += dr(l,r)*og[i]  (ACC; either grad may be NULL).
 * (dl,dr): ADD (1,1) SUB (1,-1) MUL (r,l) DIV (1/r, l/-(r*r)).
 * ref: src/ops2/binary_ew/grad.rs:22-36, grad/cpu_stack.rs:40-60; closures src/ops.rs:125-126,144-145,163-164 */
int sl_binary_ew_grad(sl_ctx* ctx, int dtype, int binop, const void* lhs, const void* rhs,
                      void* lhs_grad, void* rhs_grad, const void* out_grad, size_t n);
/* lhs_grad += og; rhs_grad += og.  ref: src/ops2/binary_ew/grad.rs:38-45, grad/cpu_stack.rs:80-88 */
int sl_add_ew_grad(sl_ctx* ctx, int dtype, void* lhs_grad, void* rhs_grad, const void* out_grad, size_t n);

/* out[i] = f(x[i]) (SET).  ref: custos `ApplyFunction::apply_fn` † as called from src/ops.rs:36,65,423,438, src/matrix.rs:181,218,246 */
int sl_unary(sl_ctx* ctx, int dtype, int unop, double p0, double p1, const void* x, void* out, size_t n);
/* x_grad[i] += g(x[i]) * out_grad[i] (ACC).  ref: custos `UnaryGrad::add_unary_grad` † as called from src/ops.rs:38-41,67-73, src/matrix.rs:183-188 */
int sl_unary_grad(sl_ctx* ctx, int dtype, int unop, double p0, double p1, const void* x, void* x_grad,
                  const void* out_grad, size_t n);

/* out[r,c] = lhs[r,c] op rhs[c] (SET, new buffer).  ref: src/ops2/row_op/mod.rs:17-38, cpu.rs:44-65,81-90 */
int sl_row_op(sl_ctx* ctx, int dtype, int binop, size_t rows, size_t cols, const void* lhs, const void* rhs, void* out);
/* = sl_row_op(ADD).  ref: src/ops2/row_op/mod.rs:28-38 */
int sl_add_row(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* lhs, const void* rhs, void* out);
/* lhs[r,c] += rhs[c] in place.  ref: src/ops2/row_op/mod.rs:40-46, cpu.rs:15-30,67-79 */
int sl_add_row_mut(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* lhs, const void* rhs);
/* lhs_grad = out_grad (SET copy); rhs_grad[c] += sum_r out_grad[r,c] (ACC).  ref: src/ops2/row_op/grad.rs:25-32, grad/cpu.rs:54-64 */
int sl_add_row_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* lhs_grad, void* rhs_grad, const void* out_grad);
/* rhs_grad[c] += sum_r out_grad[r,c] (ACC).  ref: src/ops2/row_op/grad.rs:34-40, grad/cpu.rs:42-50 */
int sl_add_row_mut_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* rhs_grad, const void* out_grad);
/* lhs_grad[r,c] += dl(rhs[c])*og[r,c]; rhs_grad[c] += sum_r dr(lhs[r,c])*og[r,c] (ACC), with
 * (dl,dr) = ADD (1,1), SUB (1,-1), MUL (v,v).  ref: src/ops2/row_op/grad.rs:13-23, grad/cpu.rs:66-88 */
int sl_row_op_grad(sl_ctx* ctx, int dtype, int binop, size_t rows, size_t cols, const void* lhs, const void* rhs,
                   void* lhs_grad, void* rhs_grad, const void* out_grad);

/* out[r,c] = lhs[r,c] op rhs[r] (SET).  ref: src/ops2/col_op/mod.rs:20-58 (sub_cols, div_cols), cpu.rs:40-50 */
int sl_col_op(sl_ctx* ctx, int dtype, int binop, size_t rows, size_t cols, const void* lhs, const void* rhs, void* out);
/* lhs_grad[r,c] += dl(l,rhs[r])*og; rhs_grad[r] += sum_c dr(l,rhs[r])*og (ACC), (dl,dr) as sl_binary_ew_grad.
 * ref: src/ops2/col_op/grad/cpu.rs:37-89 */
int sl_col_op_grad(sl_ctx* ctx, int dtype, int binop, size_t rows, size_t cols, const void* lhs, const void* rhs,
                   void* lhs_grad, void* rhs_grad, const void* out_grad);

/* w[i] -= g[i]*lr.  ref: examples/nn.rs:108-119 (SGD::step), examples/sine_net.rs:105-116 */
int sl_sgd_step(sl_ctx* ctx, int dtype, void* w, const void* g, double lr, size_t n);

/* The 5-op chained graph of examples/chained_perf.rs:86-90 in one pass:
 *   out = (x*x)*x + (b+x)*b        (SET; same operation order as the reference's 5 separate passes) */
int sl_chained_fwd(sl_ctx* ctx, int dtype, const void* x, const void* b, void* out, size_t n);
/* ... and its tape (reverse order of the 5 grad closures) in one pass:
 *   x_grad += d out/d x * og ; b_grad += d out/d b * og   (ACC) */
int sl_chained_bwd(sl_ctx* ctx, int dtype, const void* x, const void* b, void* x_grad, void* b_grad,
                   const void* out_grad, size_t n);

/* Fused element-wise chain (SURVEY 8b `sl_fused_chain`): ONE launch evaluates a short straight-line program over n elements — the
 * device analogue of custos' `Lazy` graph + `optimize()` for chains of element-wise ops (ref: examples/chained_perf.rs:86-114,
 * examples/sine_net.rs:178-233) and of the arbitrary expression closures custos compiles for apply_fn / add_unary_grad /
 * binary_ew (ref: src/ops2/binary_ew/mod.rs:55-63, src/ops.rs:36,65,423,438, src/matrix.rs:181,218,246).
 *
 * The program works on registers r[0 .. n_regs): r[i] = inputs[i][e] for i < n_in; every instruction writes one register;
 * output j stores r[out_reg[j]] (SET) or adds it to what the output holds (ACC: out = out + r).  An input may alias an output
 * (in-place; element-wise so every thread reads before it writes).  Every instruction performs exactly the IEEE operation of the
 * stand-alone op it replaces (same templates, no contraction), so a fused chain is bit-identical to the launch-per-op sequence;
 * it moves (n_in + n_out [+ n_out for ACC]) * 4 bytes per element instead of 8-28 bytes per op.  f32 / f64 / i32 (integer
 * programs may not use the transcendental unary codes). */
#define SL_CHAIN_MAX_INSTRS 32
#define SL_CHAIN_MAX_INPUTS 8
#define SL_CHAIN_MAX_OUTPUTS 4
#define SL_CHAIN_MAX_REGS 24
typedef enum sl_chain_op {
    SL_CH_ADD = 0, SL_CH_SUB = 1, SL_CH_MUL = 2, SL_CH_DIV = 3, /* r[dst] = r[a] op r[b]  (codes of sl_binop) */
    SL_CH_RDIV_IMM = 4,                                          /* r[dst] = imm0 / r[a]   (d/dl of DIV: 1 / r) */
    SL_CH_CONST = 5,                                             /* r[dst] = imm0 */
    SL_CH_COPY = 6,                                              /* r[dst] = r[a] */
    SL_CH_UNARY_F = 16, /* + sl_unop u: r[dst] = f_u(r[a]; imm0, imm1)   exactly sl_unary */
    SL_CH_UNARY_D = 48  /* + sl_unop u: r[dst] = g_u(r[a]; imm0, imm1)   the derivative sl_unary_grad multiplies out_grad by */
} sl_chain_op;
typedef struct sl_chain_instr {
    uint8_t op, dst, a, b;
    uint8_t flags, pad_[3]; /* set by the library (operand forwarding / dead-store elimination); callers leave 0 */
    double imm0, imm1;
} sl_chain_instr;
typedef struct sl_chain_prog {
    int32_t n_instr, n_in, n_out, n_regs;
    sl_chain_instr instr[SL_CHAIN_MAX_INSTRS];
    uint8_t out_reg[SL_CHAIN_MAX_OUTPUTS];
    uint8_t out_acc[SL_CHAIN_MAX_OUTPUTS];
} sl_chain_prog;
int sl_fused_chain(sl_ctx* ctx, int dtype, const sl_chain_prog* prog, const void* const* inputs, void* const* outputs, size_t n);

/* ---------------------------------------------------------------- G: gemm / trans_gemm */

/* out[m x n] = lhs[m x k] * rhs[k x n]  (SET).  mode < 0 -> ctx default.
 * ref: src/ops2/gemm/mod.rs:21-32, cpu_stack.rs:27-45 (BLAS call `T::gemm(m,n,k,..)` :39) */
int sl_gemm(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* rhs, void* out, int mode);
/* custos GenericBlas::gemmT(m,n,k,a,b,c): c[m x n] = a[m x k] * b[n x k]^T (SET).
 * ref: src/ops2/gemm/grad/cpu_stack.rs:36, tests/test_trans_gemm.rs:73-105 */
int sl_gemm_nt(sl_ctx* ctx, int dtype, size_t m, size_t n, size_t k, const void* a, const void* b, void* c, int mode);
/* custos GenericBlas::Tgemm(m,n,k,a,b,c): c[m x n] = a[k x m]^T * b[k x n] (SET).
 * ref: src/ops2/gemm/grad/cpu_stack.rs:39, tests/test_trans_gemm.rs:6-69 */
int sl_gemm_tn(sl_ctx* ctx, int dtype, size_t m, size_t n, size_t k, const void* a, const void* b, void* c, int mode);
/* General form: c[m x n] (=|+=) op(a) * op(b); trans_a: a is stored [k x m]; trans_b: b is stored [n x k]. */
int sl_gemm_ex(sl_ctx* ctx, int dtype, int trans_a, int trans_b, size_t m, size_t n, size_t k,
               const void* a, const void* b, void* c, int accumulate, int mode);
/* lhs_grad[m x k] (=|+=) out_grad * rhs^T ; rhs_grad[k x n] (=|+=) lhs^T * out_grad.
 * NULL grad pointer == `!requires_grad()`.  accumulate=0 reproduces the CPU reference (beta = 0),
 * accumulate=1 the OpenCL reference (src/ops2/gemm/grad/opencl.rs:20,33).
 * ref: src/ops2/gemm/grad.rs:14-29, grad/cpu_stack.rs:24-41 */
int sl_gemm_grad(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* rhs,
                 void* lhs_grad, void* rhs_grad, const void* out_grad, int accumulate, int mode);

/* Fused Linear backward w.r.t. the parameters: w_grad[k x n] = lhs[m x k]^T * out_grad[m x n] (SET, like sl_gemm_grad's rhs half) and
 * b_grad[n] += column sums of out_grad (sl_add_row_mut_grad).  In 3xFP16 mode the pass over out_grad that finds its per-column
 * scales also produces the column sums, so the bias gradient costs no extra read of out_grad.  b_grad may be NULL.
 * ref: src/ops2/gemm/grad.rs:29-36 (rhs_grad) + src/ops2/row_op/cpu.rs:41-48 (add_row_mut_grad), i.e. the two tape closures a
 * `Linear::forward` (examples/nn.rs:40-46) leaves behind for its parameters. */
int sl_linear_bwd_params(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* out_grad, void* w_grad, void* b_grad,
                         int mode);

/* Data-parallel form of sl_linear_bwd_params: the same gradients, each handed to sl_allreduce_sum_async as soon as it is final
 * (join with sl_comm_wait).  With chunks > 1 (3xFP16 path, k a multiple of 256*chunks) w_grad is produced in `chunks` row blocks by
 * separate kernel launches over the same operand planes and every block's exchange starts while the next block is being
 * multiplied — the exchange of the step's last weight gradient no longer waits for the whole gemm.  A no-op exchange on a
 * context without a communicator (results identical to sl_linear_bwd_params). */
int sl_linear_bwd_params_exchange(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* out_grad, void* w_grad,
                                  void* b_grad, int chunks, int mode);

/* Operand-plane reuse scope for the tensor-core gemm.  Every f32 gemm first derives half-width hi/lo planes from its operands
 * (fp16 planes scaled per row / column in the default 3xFP16 mode, tf32 planes in 3xTF32 mode); between begin and end those planes
 * are kept and reused by later gemms that read the SAME buffer (same pointer and size) — e.g. an activation in its forward gemm and
 * again in the weight-gradient gemm of the same training step (3xFP16: the row-scaled planes of the forward serve the gemm that
 * contracts over the buffer's rows, with the row scales folded exactly into the other operand's split; results stay within the
 * mode's stated tolerance but are not bit-identical to the same product computed outside a scope).  Every entry point of this
 * library that writes device memory (ops, sl_write, sl_copy, sl_clear, the all-reduce) drops the cached planes / scales of the
 * buffer it writes, so reuse is always coherent with library calls; only writes from OUTSIDE the library (another stream, another
 * library) to a buffer read by a gemm earlier in the scope are the caller's responsibility.  sl_gemm_grad opens an implicit scope
 * around its two gemms (both read out_grad). */
int sl_gemm_scope_begin(sl_ctx* ctx);
int sl_gemm_scope_end(sl_ctx* ctx);

/* Fused Linear forward — examples/nn.rs:38-46 (`gemm` + `add_row_mut`) followed by `Matrix::relu` (src/matrix.rs:169-190) in the gemm
 * epilogue:  z[m x n] = lhs[m x k] * rhs[k x n] + bias[n];  if act_out != NULL: act_out = (z >= 0) * z.   bias may be NULL.
 * Bit-identical to sl_gemm + sl_add_row_mut + sl_unary(SL_UN_RELU): the same per-element operations in the same order. f32. */
int sl_linear_fwd(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* rhs, const void* bias, void* z_out,
                  void* act_out, int mode);
/* Fused input-gradient of a Linear whose input came out of a relu:
 *   x_grad[m x k] = (z_prev[m x k] >= 0) * (out_grad[m x n] * rhs[k x n]^T)        (SET)
 * == sl_gemm_grad(lhs_grad only) followed by sl_unary_grad(SL_UN_RELU) into a zeroed gradient (ops.rs:260-275, matrix.rs:183-188). f32. */
int sl_linear_bwd_input_relu(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* rhs, const void* out_grad, const void* z_prev,
                             void* x_grad, int mode);
/* The same two fused products with the relu mask carried as ONE BIT per element instead of the pre-activation matrix
 * (mask_bits: uint32 [m x ceil(cols / 32)], bit c % 32 of word c / 32 of row r = (z[r][c] >= 0), cols = n for the forward, k for the
 * gradient): the forward stores relu(z) only, the gradient reads 1/32 of the bytes.  Results are bit-identical to sl_linear_fwd's
 * act_out and to sl_linear_bwd_input_relu (the mask of src/matrix.rs:181-188 is exactly this bit).  On the 2-CTA tensor-core path
 * (cols % 32 == 0) and in the skinny head kernel the bits are produced / consumed in the gemm epilogue; other shapes take an extra
 * element-wise pass.  f32. */
int sl_linear_fwd_bits(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* lhs, const void* rhs, const void* bias, void* act_out,
                       uint32_t* mask_bits, int mode);
int sl_linear_bwd_input_relu_bits(sl_ctx* ctx, int dtype, size_t m, size_t k, size_t n, const void* rhs, const void* out_grad,
                                  const uint32_t* mask_bits, void* x_grad, int mode);

/* ---------------------------------------------------------------- R: reductions */

/* Scalar reductions; result is written to a device scalar `out_dev` (1 element of dtype).
 * ref: src/ops2/sum/mod.rs:15-17 + cpu.rs:18-20; mean/mod.rs:15-17 + cpu.rs:34-36; max/mod.rs:16-18 + cpu.rs:11-13 */
int sl_sum(sl_ctx* ctx, int dtype, const void* x, size_t n, void* out_dev);
int sl_mean(sl_ctx* ctx, int dtype, const void* x, size_t n, void* out_dev);
int sl_max(sl_ctx* ctx, int dtype, const void* x, size_t n, void* out_dev);

/* "rows" ops reduce over rows (output length = cols); "cols" ops reduce over columns (output length = rows).
 * All SET: the output is fully overwritten (the reference's sum_rows accumulates into a fresh zeroed buffer).
 * ref: sum_rows src/ops2/sum/cpu.rs:33-41,78-84; sum_cols :54-63,86-90;
 *      mean_rows src/ops2/mean/cpu.rs:16-20,38-43; mean_cols :27-31,45-49;
 *      max_rows src/ops2/max/cpu.rs:37-54; max_cols :64-80.
 * max_*: idx_out (int32, may be NULL) receives the FIRST index attaining the maximum. */
int sl_sum_rows(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out);
int sl_sum_cols(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out);
int sl_mean_rows(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out);
int sl_mean_cols(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out);
int sl_max_rows(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out, int32_t* idx_out);
int sl_max_cols(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out, int32_t* idx_out);

/* Gradients (all ACC into x_grad).
 * sum_rows_grad:  xg[r,c] += og[c]             ref: src/ops2/sum/grad/cpu.rs:41-48
 * sum_cols_grad:  xg[r,c] += og[r]             ref: src/ops2/sum/grad/cpu.rs:50-56
 * mean_rows_grad: xg[r,c] += (T(cols)/T(len))*og[c]   ref: src/ops2/mean/grad/cpu.rs:30-47
 * mean_cols_grad: xg[r,c] += og[r]/T(cols)     ref: src/ops2/mean/grad/cpu.rs:56-63
 * max_rows_grad:  every (r,c) with x[r,c]==out[c] gets += og[c]   ref: src/ops2/max/grad/cpu.rs:40-55
 * max_cols_grad:  the FIRST c with x[r,c]==out[r] gets += og[r]   ref: src/ops2/max/grad/cpu.rs:57-69 */
int sl_sum_rows_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* x_grad, const void* out_grad);
int sl_sum_cols_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* x_grad, const void* out_grad);
int sl_mean_rows_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* x_grad, const void* out_grad);
int sl_mean_cols_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, void* x_grad, const void* out_grad);
int sl_max_rows_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* out, const void* x, void* x_grad,
                     const void* out_grad);
int sl_max_cols_grad(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* out, const void* x, void* x_grad,
                     const void* out_grad);
/* Same result as sl_max_cols_grad, using the argmax saved by sl_max_cols (8 B per row instead of a scan). */
int sl_max_cols_grad_idx(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const int32_t* idx, void* x_grad,
                         const void* out_grad);

/* ---------------------------------------------------------------- T: transpose */

/* out[c*rows + r] (=|+=) x[r*cols + c].  accumulate=0: SET.
 * ref: src/ops2/transpose/mod.rs:17-21, cpu.rs:10-26,41-52; grad: transpose/grad.rs:13-26, grad/cpu.rs:17-25
 * (the CPU reference's grad is effectively SET — stray `b[idx] = row.clone()` at transpose/cpu.rs:23 —
 *  the OpenCL reference's is ACC; both are available through `accumulate`). */
int sl_transpose(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* x, void* out, int accumulate);

/* ---------------------------------------------------------------- S: softmax */

/* out[r,:] = exp(x[r,:] - max_c x[r,c]) / sum_c exp(..)  (SET; f32/f64).
 * ref: src/ops2/softmax/mod.rs:16-18, cpu.rs:11-16 */
int sl_softmax(sl_ctx* ctx, int dtype, size_t samples, size_t features, const void* x, void* out);
/* x_grad[r,:] = (diag(s) - s s^T) g  ==  s * (g - <s,g>)   (SET; f32/f64).
 * ref: src/ops2/softmax/grad.rs:15-24, grad/cpu.rs:14-62 */
int sl_softmax_grad(sl_ctx* ctx, int dtype, size_t samples, size_t features, void* x_grad, const void* out,
                    const void* out_grad);

/* Fused softmax + categorical cross-entropy, forward and backward, for the tail of examples/nn.rs:190-233 in one kernel:
 *   probs_out       = softmax(logits)                                             (SET)   ref: src/ops2/softmax/cpu.rs:11-16
 *   loss_per_sample = -ln(sum_c clip(probs, 1e-7, 1 - 1e-7) * targets)  [samples] (SET)   ref: examples/nn.rs:124-138 (cce)
 *   logits_grad     = softmax_grad(probs, (-(targets / probs)) / grad_rows)       (SET)   ref: examples/nn.rs:140-152 (cce_grad: the
 *                     division uses the UNCLIPPED probabilities) + src/ops2/softmax/grad/cpu.rs:14-62
 *   *correct_dev   += #rows whose first arg-max of probs equals labels[row]  (labels / correct_dev may be NULL)  ref: nn.rs:195-211
 * features <= 32 runs one thread per row and is bit-identical to the chain sl_softmax, sl_unary(CLIP), sl_binary_ew(MUL),
 * sl_sum_cols, sl_unary(NEG_LN), sl_binary_ew(DIV), sl_unary(NEG_DIV_SCALAR), sl_softmax_grad, sl_count_correct; wider rows use
 * fixed-order block reductions (K-scaled tolerance).  f32 / f64. */
int sl_softmax_cce(sl_ctx* ctx, int dtype, size_t samples, size_t features, const void* logits, const void* targets, const int32_t* labels,
                   size_t grad_rows, void* probs_out, void* logits_grad, void* loss_per_sample, int32_t* correct_dev);

/* The WHOLE training step of a small squared-error MLP — forward, loss, backward, SGD — in one launch of one 16-CTA cluster (8 below 256 samples)
 * (weights, activations and partial gradients in shared memory; gradients summed over distributed shared memory in rank order):
 *   a_l = relu(a_{l-1} W_l + b_l) on hidden layers, out = a_{L-1} W_L + b_L          ref: examples/sine_net.rs:135-147 (Linear + relu)
 *   *loss_sum_dev = sum (out - y)^2 ; tape seed 1 -> d out = 2 (out - y)             ref: examples/sine_net.rs:150-156
 *   grads = parameter gradients of that tape (SET), params -= grads * lr             ref: examples/sine_net.rs:108-116,158-160
 * params / grads are flat: layer l's W [dims[l] x dims[l+1]] at float offset seg_off[2l], its bias at seg_off[2l+1].
 * 1..4 layers of width <= 64, any batch, f32; sl_mlp_small_fits tells whether a shape is taken (SL_ERR_UNSUPPORTED otherwise).
 * Same arithmetic per element as the op-by-op tape, different (fixed) summation order: K-scaled tolerance against it. */
int sl_mlp_small_fits(sl_ctx* ctx, int n_layers, const size_t* dims, size_t batch);
int sl_mlp_small_step(sl_ctx* ctx, int dtype, int n_layers, const size_t* dims, const size_t* seg_off, size_t batch, const void* x,
                      const void* y, void* params, void* grads, double lr, void* loss_sum_dev);

/* ---------------------------------------------------------------- next-row ops */

/* out[i*n + i] = x[i] (only the diagonal is written).  ref: src/ops2/diagflat/cpu.rs:42-46 */
int sl_diagflat(sl_ctx* ctx, int dtype, size_t n, const void* x, void* out);
/* x_grad[i] += out_grad[i*n + i].  ref: src/ops2/diagflat/grad/cpu.rs:32-36 */
int sl_diagflat_grad(sl_ctx* ctx, int dtype, size_t n, void* x_grad, const void* out_grad);
/* out[i*highest_class + (size_t)classes[i]] = 1 (only the ones are written).  ref: src/ops2/onehot/cpu.rs:40-44 */
int sl_onehot(sl_ctx* ctx, int dtype, size_t n, size_t highest_class, const void* classes, void* out);
/* classes_grad[i] += out_grad[i*highest_class + classes[i]].  ref: src/ops2/onehot/grad/cpu.rs:3-12 */
int sl_onehot_grad(sl_ctx* ctx, int dtype, size_t n, size_t highest_class, const void* classes, void* classes_grad,
                   const void* out_grad);
/* Per-row argmax of preds[rows x cols] compared with int32 labels -> number of matches in *count_dev (int32 device scalar).
 * ref: examples/nn.rs:195-211 (host loop in the reference). */
int sl_count_correct(sl_ctx* ctx, int dtype, size_t rows, size_t cols, const void* preds, const int32_t* labels,
                     int32_t* count_dev);

/* ---------------------------------------------------------------- DP: data-parallel exchange (new) */

#define SL_COMM_ID_BYTES 128
/* Fills a 128-byte opaque id on ONE rank; the caller ships it to the other ranks (e.g. torch.distributed broadcast). */
int sl_comm_unique_id(void* id_out_128);
int sl_comm_init_rank(sl_ctx* ctx, int nranks, int rank, const void* id_128);
/* In-place sum all-reduce of a device buffer across the ranks of the communicator (NCCL over NVLink). */
int sl_allreduce_sum(sl_ctx* ctx, int dtype, void* buf, size_t n);
/* Overlapped variant: the all-reduce is enqueued on the context's communication stream, ordered after everything issued so far on
 * the compute stream, and the compute stream continues (e.g. with the remaining backward gemms).  sl_comm_wait makes the compute
 * stream wait for every exchange issued so far (call it before the SGD step reads the gradients). */
int sl_allreduce_sum_async(sl_ctx* ctx, int dtype, void* buf, size_t n);
int sl_comm_wait(sl_ctx* ctx);
/* Finer join: the compute stream waits only for the first n exchanges issued (sl_allreduce_sum_async) since the last sl_comm_wait —
 * they complete in issue order — so the update of a layer whose gradients have arrived runs while later exchanges are still in
 * flight.  sl_comm_issued returns how many have been issued since the last sl_comm_wait. */
int sl_comm_wait_n(sl_ctx* ctx, int n);
int sl_comm_issued(sl_ctx* ctx);
/* Ranks of the context's communicator (1 without one). */
int sl_comm_nranks(sl_ctx* ctx);
int sl_comm_destroy(sl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SLICED_B200_H */
