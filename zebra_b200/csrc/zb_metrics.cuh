// zb_metrics.cuh -- the ten scalar metrics of /root/reference/src/distance.rs:51-190 (Chebyshev, Canberra, Bray-Curtis,
// Manhattan, L3, L4, Hamming, Minkowski, p-norm), SURVEY.md 8(f) row 3.
//
// The reference evaluates them with the `distances` crate (^1.8.0, Cargo.toml:38): one f32 accumulator per pair,
// folded over the elements IN INPUT ORDER, result `.to_bits()` (u32) zero-extended to DistanceUnit (distance.rs:59,
// :71, :83, :95, :124, :136, :171, :188; quirk Q6).  A strictly sequential fold has no parallelism inside a pair, so on
// the device ONE THREAD owns a pair and walks the row once (128-bit loads); the parallelism is across pairs.  Every
// operation is an explicitly rounded IEEE intrinsic (no contraction, no reassociation), so the sums are bit-identical to
// the scalar CPU order.  cbrt / powf(., 1/p) are not IEEE operations: both sides use the deterministic Newton root
// `root_p` below (IEEE f64 operations only), see oracle/README.md "scalar metrics".
//
// The functions are __host__ __device__: tests/metric_twin.cpp compiles this very header for the CPU (g++,
// -ffp-contract=off) so the arithmetic the kernels execute is checked against the oracle without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define ZB_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define ZB_HD static inline
#endif

namespace zb {

enum MetricCode {  // == zb_metric of include/zebra_b200.h
    M_COSINE = 0, M_L2SQ = 1, M_L2 = 2,
    M_CHEBYSHEV = 3, M_CANBERRA = 4, M_BRAY_CURTIS = 5, M_MANHATTAN = 6, M_L3 = 7, M_L4 = 8, M_HAMMING = 9,
    M_MINKOWSKI = 10, M_PNORM = 11, M_COUNT = 12
};
#define ZB_MAX_POWER 64

// ---- explicitly rounded primitives (device: intrinsics the compiler never fuses; host: plain IEEE operations) ----
#if defined(__CUDA_ARCH__)
ZB_HD float m_add(float a, float b) { return __fadd_rn(a, b); }
ZB_HD float m_sub(float a, float b) { return __fsub_rn(a, b); }
ZB_HD float m_mul(float a, float b) { return __fmul_rn(a, b); }
ZB_HD float m_div(float a, float b) { return __fdiv_rn(a, b); }
ZB_HD float m_sqrt(float a) { return __fsqrt_rn(a); }
ZB_HD float m_abs(float a) { return fabsf(a); }
ZB_HD double d_add(double a, double b) { return __dadd_rn(a, b); }
ZB_HD double d_mul(double a, double b) { return __dmul_rn(a, b); }
ZB_HD double d_div(double a, double b) { return __ddiv_rn(a, b); }
ZB_HD float d_to_f(double a) { return __double2float_rn(a); }
ZB_HD uint32_t f_bits(float a) { return __float_as_uint(a); }
ZB_HD uint64_t d_bits(double a) { return (uint64_t)__double_as_longlong(a); }
ZB_HD double d_from_bits(uint64_t u) { return __longlong_as_double((long long)u); }
ZB_HD uint32_t popc8(uint32_t x) { return (uint32_t)__popc(x & 0xFFu); }
#else
ZB_HD float m_add(float a, float b) { return a + b; }
ZB_HD float m_sub(float a, float b) { return a - b; }
ZB_HD float m_mul(float a, float b) { return a * b; }
ZB_HD float m_div(float a, float b) { return a / b; }
ZB_HD float m_sqrt(float a) { return sqrtf(a); }
ZB_HD float m_abs(float a) { return fabsf(a); }
ZB_HD double d_add(double a, double b) { return a + b; }
ZB_HD double d_mul(double a, double b) { return a * b; }
ZB_HD double d_div(double a, double b) { return a / b; }
ZB_HD float d_to_f(double a) { return (float)a; }
ZB_HD uint32_t f_bits(float a) { uint32_t u; memcpy(&u, &a, 4); return u; }
ZB_HD uint64_t d_bits(double a) { uint64_t u; memcpy(&u, &a, 8); return u; }
ZB_HD double d_from_bits(uint64_t u) { double a; memcpy(&a, &u, 8); return a; }
ZB_HD uint32_t popc8(uint32_t x) { return (uint32_t)__builtin_popcount(x & 0xFFu); }
#endif

// f32 result -> DistanceUnit: the bits, zero-extended; NaN canonicalised to x86's default NaN (0.0 / 0.0 on the
// reference's host), because the GPU's canonical NaN has another bit pattern.
ZB_HD uint64_t f32_key(float r) { return r != r ? 0xFFC00000ull : (uint64_t)f_bits(r); }

// compiler-rt's __powisf2 / __powidf2 (what f32::powi lowers to), b >= 0: square and multiply.
ZB_HD float powi_f32(float a, int b) {
    float r = 1.0f;
    for (;;) {
        if (b & 1) r = m_mul(r, a);
        b /= 2;
        if (b == 0) break;
        a = m_mul(a, a);
    }
    return r;
}
ZB_HD double powi_f64(double a, int b) {
    double r = 1.0;
    for (;;) {
        if (b & 1) r = d_mul(r, a);
        b /= 2;
        if (b == 0) break;
        a = d_mul(a, a);
    }
    return r;
}

// p-th root of a non-negative f32, p >= 1: y <- ((p-1) y + x / y^(p-1)) / p in f64 from a bit-pattern guess, stopped at
// the first step that does not decrease y (after the first step Newton descends monotonically onto the root), then
// rounded once to f32.  Within 1 ulp(f32) of cbrtf / powf; identical on CPU and GPU by construction.
ZB_HD float root_p(float s, int p) {
    if (p == 1 || s != s || s == 0.0f || f_bits(s) == 0x7F800000u) return s;
    const double x = (double)s;
    const uint64_t by = d_bits(x) / (uint64_t)p + (0x3FF0000000000000ull / (uint64_t)p) * (uint64_t)(p - 1);
    double y = d_from_bits(by);
    for (int it = 0; it < 1000; ++it) {
        const double t = d_div(x, powi_f64(y, p - 1));
        const double yn = d_div(d_add(d_mul((double)(p - 1), y), t), (double)p);
        if (it > 0 && yn >= y) break;
        y = yn;
    }
    return d_to_f(y);
}

struct SeqAcc {
    float acc, den;
    uint32_t ham;
};
ZB_HD void seq_init(SeqAcc& st) {
    st.acc = 0.0f;
    st.den = 0.0f;
    st.ham = 0u;
}
// one element of the fold: x = stored row element, y = query element (argument order of lsh.rs:314, :559)
template <int CODE>
ZB_HD void seq_step(SeqAcc& st, float x, float y, int power) {
    if (CODE == M_HAMMING) {  // distance.rs:147-148: `x.to_bits() as u8` keeps the low byte of the bit pattern
        st.ham += popc8(f_bits(x) ^ f_bits(y));
        return;
    }
    const float v = m_abs(m_sub(x, y));
    if (CODE == M_CHEBYSHEV) st.acc = st.acc > v ? st.acc : v;
    else if (CODE == M_CANBERRA) st.acc = m_add(st.acc, m_div(v, m_add(m_abs(x), m_abs(y))));
    else if (CODE == M_BRAY_CURTIS) {
        st.acc = m_add(st.acc, v);
        st.den = m_add(st.den, m_abs(m_add(x, y)));
    } else if (CODE == M_MANHATTAN) st.acc = m_add(st.acc, v);
    else if (CODE == M_L3) st.acc = m_add(st.acc, m_mul(m_mul(v, v), v));
    else if (CODE == M_L4) {
        const float v2 = m_mul(v, v);
        st.acc = m_add(st.acc, m_mul(v2, v2));
    } else st.acc = m_add(st.acc, powi_f32(v, power));  // Minkowski, p-norm
}
template <int CODE>
ZB_HD uint64_t seq_finish(const SeqAcc& st, int power) {
    if (CODE == M_HAMMING) return (uint64_t)st.ham;
    if (CODE == M_BRAY_CURTIS) return f32_key(m_div(st.acc, st.den));
    if (CODE == M_L3) return f32_key(root_p(st.acc, 3));
    if (CODE == M_L4) return f32_key(m_sqrt(m_sqrt(st.acc)));
    if (CODE == M_MINKOWSKI) {
        if (power == 0) return st.acc == 1.0f ? 0x3F800000ull : 0x7F800000ull;  // powf(n, 1/0 = +inf), n >= 1
        return f32_key(root_p(st.acc, power));
    }
    return f32_key(st.acc);
}

// whole pair, a and b 16-byte aligned with at least ceil(dim / 4) * 4 readable floats (rows are padded to 16)
template <int CODE>
ZB_HD uint64_t seq_distance(const float* a, const float* b, int dim, int power) {
    SeqAcc st;
    seq_init(st);
    for (int i = 0; i < dim; ++i) seq_step<CODE>(st, a[i], b[i], power);
    return seq_finish<CODE>(st, power);
}

}  // namespace zb
