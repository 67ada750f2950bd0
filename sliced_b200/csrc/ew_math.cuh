// ew_math.cuh — the per-element arithmetic of the element-wise family, shared by the fixed-function kernels (elementwise.cu) and the
// micro-op interpreter (chain.cu): binary operators and their partial derivatives, the unary functions custos `apply_fn` /
// `add_unary_grad` are called with from sliced, and their derivatives — written exactly as the reference's closures
// (src/ops.rs, src/matrix.rs; see include/sliced_b200.h) so that every path produces the same bits.
#pragma once

#include "common.cuh"

namespace {

__device__ __forceinline__ float m_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double m_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float m_exp(float a) { return expf(a); }
__device__ __forceinline__ double m_exp(double a) { return exp(a); }
__device__ __forceinline__ float m_log(float a) { return logf(a); }
__device__ __forceinline__ double m_log(double a) { return log(a); }
__device__ __forceinline__ float m_tanh(float a) { return tanhf(a); }
__device__ __forceinline__ double m_tanh(double a) { return tanh(a); }
// integer instantiations exist only so that the templates compile; the host rejects them before launch
__device__ __forceinline__ int32_t m_pow(int32_t a, int32_t) { return a; }
__device__ __forceinline__ int32_t m_exp(int32_t a) { return a; }
__device__ __forceinline__ int32_t m_log(int32_t a) { return a; }
__device__ __forceinline__ int32_t m_tanh(int32_t a) { return a; }

template <int OP, typename T>
__device__ __forceinline__ T binop(T l, T r) {
    if (OP == SL_ADD) return l + r;
    if (OP == SL_SUB) return l - r;
    if (OP == SL_MUL) return l * r;
    return l / r;
}
// d out / d lhs and d out / d rhs, written as the reference's grad closures (src/ops.rs:125-126,144-145,163-164)
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

template <int OP, typename T>
__device__ __forceinline__ T unary_f(T x, T p0, T p1) {
    switch (OP) {
    case SL_UN_SQUARE: return x * x;
    case SL_UN_POW:
        // small integer exponents (sine_net's pow(2.), tests' pow(3.)) as products: <= 1.5 ulp from the exact power, inside
        // the 1e-6 tolerance, and the op stays HBM-bound instead of powf-bound (warp-uniform branch on the scalar p0)
        if (p0 == T(2)) return x * x;
        if (p0 == T(3)) return x * x * x;
        return m_pow(x, p0);
    case SL_UN_RELU: return T(x >= T(0) ? 1 : 0) * x;
    case SL_UN_TANH: return m_tanh(x);
    case SL_UN_SIGMOID: return T(1) / (T(1) + m_exp(-x));
    case SL_UN_EXP: return m_exp(x);
    case SL_UN_LN: return m_log(x);
    case SL_UN_NEG_LN: return -m_log(x);
    case SL_UN_CLIP: { T m = x > p0 ? x : p0; return m < p1 ? m : p1; }
    case SL_UN_NEG: return -x;
    case SL_UN_MUL_SCALAR: return x * p0;
    case SL_UN_NEG_DIV_SCALAR: return (-x) / p0;
    case SL_UN_ADD_SCALAR: return x + p0;
    }
    return x;
}
template <int OP, typename T>
__device__ __forceinline__ T unary_d(T x, T p0, T p1) {
    switch (OP) {
    case SL_UN_SQUARE: return x * T(2);
    case SL_UN_POW:
        if (p0 == T(2)) return x * p0;          // x^(2-1) * 2
        if (p0 == T(3)) return (x * x) * p0;    // x^(3-1) * 3
        return m_pow(x, p0 - T(1)) * p0;
    case SL_UN_RELU: return T(x >= T(0) ? 1 : 0);
    case SL_UN_TANH: { T t = m_tanh(x); return T(1) - t * t; }
    case SL_UN_SIGMOID: { T e = m_exp(-x); T d = T(1) + e; return e / (d * d); }
    case SL_UN_EXP: return m_exp(x);
    case SL_UN_LN: return T(1) / x;
    case SL_UN_NEG_LN: return -(T(1) / x);
    case SL_UN_CLIP: return T((p0 <= x && x <= p1) ? 1 : 0);
    case SL_UN_NEG: return -T(1);
    case SL_UN_MUL_SCALAR: return p0;
    case SL_UN_NEG_DIV_SCALAR: return (-T(1)) / p0;
    case SL_UN_ADD_SCALAR: return T(1);
    }
    return T(1);
}


}  // namespace
