/*
 * sliced_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, single thread, -ffp-contract=off) of the reference's custos-CPU implementation of
 * sliced's forward + backward op set (elftausend/sliced, /root/reference).  It exists only as the checker for
 * the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Nothing under sliced_b200/ links, imports or calls it.
 *
 * The reference itself cannot be compiled here (no rustc/cargo; `custos`, `rawsliced`, `graplot`, `purpur` are
 * un-vendored dependencies — Cargo.toml:11,31-33), so this restatement is PINNED by the reference's own golden
 * vectors instead: tests/test_oracle_golden.py replays every known-answer test listed in SURVEY.md Appendix B.
 *
 * Third-party arithmetic restated here because it lives outside /root/reference:
 *   - custos `GenericBlas::{gemm,gemmT,Tgemm}` -> system CBLAS sgemm/dgemm (custos: path dep `../custos`, git branch
 *     `autograd`, NO pinned version — no Cargo.lock, .gitignore:2).  Restated as the textbook product with
 *     sequential-k accumulation; a CBLAS sgemm can be plugged in with orc_set_sgemm() for CPU-baseline timing
 *     (that is what the reference links under its `blas` feature).
 *   - custos `ApplyFunction::apply_fn` / `UnaryGrad::add_unary_grad` / tape (`backward` seeds ones, reverse order).
 *
 * The typed functions live in sliced_oracle_typed.inc (instantiated for f32 / f64 / i32 below).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T float
#define SUF f32
#define IS_FLOAT 1
#define T_IS_F32 1
#include "sliced_oracle_typed.inc"
#undef T
#undef SUF
#undef IS_FLOAT
#undef T_IS_F32

#define T double
#define SUF f64
#define IS_FLOAT 1
#include "sliced_oracle_typed.inc"
#undef T
#undef SUF
#undef IS_FLOAT

#define T int32_t
#define SUF i32
#define IS_FLOAT 0
#include "sliced_oracle_typed.inc"
#undef T
#undef SUF
#undef IS_FLOAT

/* ------------------------------------------------------------------------------------------------
 * Optional CBLAS sgemm hook (row-major), signature of cblas_sgemm with 32-bit ints (LP64 OpenBLAS).
 * ---------------------------------------------------------------------------------------------- */
typedef void (*cblas_sgemm_fn)(int order, int transa, int transb, int m, int n, int k, float alpha, const float* a,
                               int lda, const float* b, int ldb, float beta, float* c, int ldc);
static cblas_sgemm_fn g_sgemm = NULL;
void orc_set_sgemm(void* fn) { g_sgemm = (cblas_sgemm_fn)fn; }

/* f32 gemm that goes through the plugged CBLAS when present (what the reference does), else the restatement */
void orc_sgemm(int trans_a, int trans_b, size_t m, size_t n, size_t k, const float* a, const float* b, float* c,
               int accumulate) {
    if (g_sgemm) {
        int lda = trans_a ? (int)m : (int)k;
        int ldb = trans_b ? (int)k : (int)n;
        g_sgemm(101 /*RowMajor*/, trans_a ? 112 : 111, trans_b ? 112 : 111, (int)m, (int)n, (int)k, 1.0f, a, lda, b, ldb,
                accumulate ? 1.0f : 0.0f, c, (int)n);
    } else {
        orc_gemm_ex_f32(trans_a, trans_b, m, n, k, a, b, c, accumulate);
    }
}

/* fp64-accumulated truth for tolerance budgeting of gemm parity tests */
void orc_gemm_truth_f32(int trans_a, int trans_b, size_t m, size_t n, size_t k, const float* a, const float* b, double* c) {
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) c[i * n + j] = 0.0;
    for (size_t i = 0; i < m; ++i)
        for (size_t p = 0; p < k; ++p) {
            double av = trans_a ? a[p * m + i] : a[i * k + p];
            for (size_t j = 0; j < n; ++j) {
                double bv = trans_b ? b[j * k + p] : b[p * n + j];
                c[i * n + j] += av * bv;
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * MLP training step of examples/nn.rs:184-237 (loss_kind 0: relu MLP -> softmax -> cce) and of
 * examples/sine_net.rs:135-163 (loss_kind 1: relu MLP -> (out - y)^2), replayed op by op in the reference's
 * order, including the tape (reverse registration order) and SGD::step.
 *
 *   dims[0..n_layers]           layer widths; W[l] is dims[l] x dims[l+1], B[l] is 1 x dims[l+1]
 *   x [batch x dims[0]]         no_grad input;  y [batch x dims[n_layers]] targets (one-hot for loss_kind 0)
 *   labels                      class ids for the accuracy count (may be NULL)
 *   grad_rows                   the `rows` cce_grad divides by (nn.rs:151) — the GLOBAL batch under DP sharding
 *   apply_sgd                   0: leave W/B untouched and only return the gradients in dW/dB (DP: all-reduce first)
 *   dW/dB                       gradient outputs (each may be NULL)
 *   loss_sum_out                sum of the per-sample loss (nn.rs:213 then summed; mean = /batch), or of (out-y)^2
 *   correct_out                 number of rows whose argmax equals labels (nn.rs:195-211, strict `>` scan)
 * ---------------------------------------------------------------------------------------------- */
int orc_mlp_step_f32(int loss_kind, int n_layers, const size_t* dims, size_t batch, const float* x, const float* y,
                     const int32_t* labels, float** W, float** B, float lr, size_t grad_rows, int apply_sgd,
                     float** dW, float** dB, double* loss_sum_out, int64_t* correct_out) {
    if (n_layers < 1 || n_layers > 16) return -1;
    float* z[16];   /* pre-activation (after in-place add_row_mut) */
    float* a[16];   /* post-activation (relu) ; last layer: softmax output / identity */
    float* gz[16];
    float* ga[16];
    for (int l = 0; l < n_layers; ++l) {
        size_t sz = batch * dims[l + 1];
        z[l] = (float*)malloc(sz * sizeof(float));
        a[l] = (float*)malloc(sz * sizeof(float));
        gz[l] = (float*)calloc(sz, sizeof(float)); /* zero_grad(): nn.rs:186-188 */
        ga[l] = (float*)calloc(sz, sizeof(float));
    }
    /* forward: Linear::forward (nn.rs:38-46) = gemm + add_row_mut ; then relu / softmax */
    for (int l = 0; l < n_layers; ++l) {
        const float* in = l == 0 ? x : a[l - 1];
        orc_sgemm(0, 0, batch, dims[l + 1], dims[l], in, W[l], z[l], 0);
        orc_add_row_mut_f32(batch, dims[l + 1], z[l], B[l]);
        if (l + 1 < n_layers) orc_unary_f32(2, 0, 0, z[l], a[l], batch * dims[l + 1]);
        else if (loss_kind == 0) orc_softmax_f32(batch, dims[l + 1], z[l], a[l]);
        else memcpy(a[l], z[l], batch * dims[l + 1] * sizeof(float));
    }
    int L = n_layers - 1;
    size_t oc = dims[n_layers];
    size_t on = batch * oc;
    float* out = a[L];
    /* accuracy: nn.rs:195-211 */
    if (correct_out) {
        int64_t correct = 0;
        if (labels)
            for (size_t r = 0; r < batch; ++r) {
                float mx = out[r * oc];
                size_t mi = 0;
                for (size_t c = 1; c < oc; ++c)
                    if (out[r * oc + c] > mx) { mx = out[r * oc + c]; mi = c; }
                if ((size_t)labels[r] == mi) ++correct;
            }
        *correct_out = correct;
    }
    float* gout = (float*)calloc(on, sizeof(float));
    double loss_sum = 0.0;
    if (loss_kind == 0) {
        /* cce: nn.rs:124-138  clip -> mul(targets) -> sum_cols -> -ln */
        float* t1 = (float*)malloc(on * sizeof(float));
        float* t2 = (float*)malloc(batch * sizeof(float));
        orc_unary_f32(8, 1E-7, 1. - 1E-7, out, t1, on);
        orc_binary_ew_f32(2, t1, y, t1, on);
        orc_sum_cols_f32(batch, oc, t1, t2);
        orc_unary_f32(7, 0, 0, t2, t2, batch);
        for (size_t r = 0; r < batch; ++r) loss_sum += t2[r];
        /* cce_grad: nn.rs:140-152  div(targets, preds) -> (-v)/rows */
        orc_binary_ew_f32(3, y, out, gout, on);
        orc_unary_f32(11, (double)grad_rows, 0, gout, gout, on);
        free(t1); free(t2);
        /* backward_with(grad): tape in reverse. softmax_grad (SET) — closed form == Jacobian form up to rounding;
         * the Jacobian form is O(F^2) per row, use it for small F as the reference does. */
        if (oc <= 64) orc_softmax_grad_f32(batch, oc, gz[L], out, gout);
        else orc_softmax_grad_closed_f32(batch, oc, gz[L], out, gout);
    } else {
        /* sine_net.rs:150 loss = (out - y).pow(2.) ; loss.backward() seeds ones */
        float* d = (float*)malloc(on * sizeof(float));
        float* gd = (float*)calloc(on, sizeof(float));
        float* ones = (float*)malloc(on * sizeof(float));
        float* lossv = (float*)malloc(on * sizeof(float));
        orc_binary_ew_f32(1, out, y, d, on);
        orc_unary_f32(1, 2.0, 0, d, lossv, on);
        for (size_t i = 0; i < on; ++i) { loss_sum += lossv[i]; ones[i] = 1.0f; }
        orc_unary_grad_f32(1, 2.0, 0, d, gd, ones, on);               /* pow grad */
        orc_binary_ew_grad_f32(1, out, y, gz[L], NULL, gd, on);       /* sub grad (y is no_grad) ; out == z[L] */
        free(d); free(gd); free(ones); free(lossv);
    }
    if (loss_sum_out) *loss_sum_out = loss_sum;
    for (int l = L; l >= 0; --l) {
        size_t no = dims[l + 1], ni = dims[l];
        if (l < L) orc_unary_grad_f32(2, 0, 0, z[l], gz[l], ga[l], batch * no);   /* relu grad: matrix.rs:183-188 */
        float* gb = (float*)calloc(no, sizeof(float));
        float* gw = (float*)malloc(ni * no * sizeof(float));
        orc_add_row_mut_grad_f32(batch, no, gb, gz[l]);                             /* ops.rs:367-373 */
        const float* in = l == 0 ? x : a[l - 1];
        if (l > 0) orc_sgemm(0, 1, batch, ni, no, gz[l], W[l], ga[l - 1], 0);       /* gemmT(m,k,n,og,rhs,lhs_grad) */
        orc_sgemm(1, 0, ni, no, batch, in, gz[l], gw, 0);                           /* Tgemm(k,n,m,lhs,og,rhs_grad) */
        if (dW && dW[l]) memcpy(dW[l], gw, ni * no * sizeof(float));
        if (dB && dB[l]) memcpy(dB[l], gb, no * sizeof(float));
        if (apply_sgd) {
            orc_sgd_step_f32(W[l], gw, lr, ni * no);
            orc_sgd_step_f32(B[l], gb, lr, no);
        }
        free(gb); free(gw);
    }
    for (int l = 0; l < n_layers; ++l) { free(z[l]); free(a[l]); free(gz[l]); free(ga[l]); }
    free(gout);
    return 0;
}
