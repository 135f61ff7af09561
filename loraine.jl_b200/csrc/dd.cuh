// Double-double ("Float64x2": an unevaluated sum hi + lo of two doubles, |lo| <= ulp(hi)/2, ~106 significant bits) arithmetic
// for the high-precision LP path (ddlp.cu).  The reference gets this element type from MultiFloats.jl (Project.toml:
// MultiFloats = "2", `Optimizer{Float64x2}`, examples/k.jl:8, README.md:37-54); the algorithms here are the published
// error-free transformations (Knuth two-sum, Dekker quick-two-sum, FMA two-product) and the QD-library compositions.
// Every operation is spelled with explicitly rounded intrinsics on the device so that nvcc cannot contract or reassociate
// them; the host versions (same code, used by the CPU unit test) need -ffp-contract=off.
#pragma once
#include <cuda_runtime.h>
#include <cmath>

namespace lrn {

struct __align__(16) dd {
    double hi, lo;
};

#ifdef __CUDA_ARCH__
#define DD_ADD(a, b) __dadd_rn((a), (b))
#define DD_MUL(a, b) __dmul_rn((a), (b))
#define DD_FMA(a, b, c) __fma_rn((a), (b), (c))
#define DD_DIV(a, b) __ddiv_rn((a), (b))
#define DD_SQRT(a) __dsqrt_rn((a))
#else
#define DD_ADD(a, b) ((a) + (b))
#define DD_MUL(a, b) ((a) * (b))
#define DD_FMA(a, b, c) std::fma((a), (b), (c))
#define DD_DIV(a, b) ((a) / (b))
#define DD_SQRT(a) std::sqrt((a))
#endif
#define DD_FN __host__ __device__ __forceinline__

DD_FN dd dd_make(double hi, double lo = 0.0) {
    dd r;
    r.hi = hi;
    r.lo = lo;
    return r;
}

// s + e = a + b exactly (no assumption on the magnitudes)
DD_FN dd two_sum(double a, double b) {
    const double s = DD_ADD(a, b);
    const double bb = DD_ADD(s, -a);
    const double e = DD_ADD(DD_ADD(a, -DD_ADD(s, -bb)), DD_ADD(b, -bb));
    return dd_make(s, e);
}
// s + e = a + b exactly when |a| >= |b|
DD_FN dd quick_two_sum(double a, double b) {
    const double s = DD_ADD(a, b);
    const double e = DD_ADD(b, -DD_ADD(s, -a));
    return dd_make(s, e);
}
// p + e = a * b exactly
DD_FN dd two_prod(double a, double b) {
    const double p = DD_MUL(a, b);
    const double e = DD_FMA(a, b, -p);
    return dd_make(p, e);
}

DD_FN dd dd_neg(dd a) { return dd_make(-a.hi, -a.lo); }

// accurate ("IEEE") sum: relative error <= 2 * 2^-106
DD_FN dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo = DD_ADD(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = DD_ADD(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
DD_FN dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
DD_FN dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo = DD_ADD(s.lo, a.lo);
    return quick_two_sum(s.hi, s.lo);
}

DD_FN dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = DD_ADD(p.lo, DD_FMA(a.hi, b.lo, DD_MUL(a.lo, b.hi)));
    return quick_two_sum(p.hi, p.lo);
}
DD_FN dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo = DD_FMA(a.lo, b, p.lo);
    return quick_two_sum(p.hi, p.lo);
}
// acc + a * b
DD_FN dd dd_fma(dd a, dd b, dd acc) { return dd_add(acc, dd_mul(a, b)); }
// acc - a * b
DD_FN dd dd_fms(dd a, dd b, dd acc) { return dd_add(acc, dd_neg(dd_mul(a, b))); }

// accurate quotient (three correction steps of long division)
DD_FN dd dd_div(dd a, dd b) {
    const double q1 = DD_DIV(a.hi, b.hi);
    dd r = dd_sub(a, dd_mul_d(b, q1));
    const double q2 = DD_DIV(r.hi, b.hi);
    r = dd_sub(r, dd_mul_d(b, q2));
    const double q3 = DD_DIV(r.hi, b.hi);
    const dd q = quick_two_sum(q1, q2);
    return dd_add_d(q, q3);
}
DD_FN dd dd_recip(dd b) { return dd_div(dd_make(1.0), b); }

// sqrt(a) ~ a x + (a - (a x)^2) x / 2 with x = 1/sqrt(a.hi) (Karp & Markstein); a > 0
DD_FN dd dd_sqrt(dd a) {
    if (a.hi == 0.0 && a.lo == 0.0) return dd_make(0.0);
    const double x = DD_DIV(1.0, DD_SQRT(a.hi));
    const double ax = DD_MUL(a.hi, x);
    const dd sq = two_prod(ax, ax);
    const dd rem = dd_sub(a, sq);
    return dd_add_d(dd_make(ax), DD_MUL(rem.hi, DD_MUL(x, 0.5)));
}

// 1/sqrt(a): one Newton step in double-double from the double estimate x: x + x (1 - a x^2) / 2; a > 0.  (The Cholesky tile
// kernel needs both sqrt(a) = a * rsqrt(a) and its reciprocal: one rsqrt is ~1/3 of a sqrt followed by a division.)
DD_FN dd dd_rsqrt(dd a) {
    const double x = DD_DIV(1.0, DD_SQRT(a.hi));
    const dd ax2 = dd_mul(a, two_prod(x, x));
    const dd e = dd_sub(dd_make(1.0), ax2);
    return dd_add(dd_make(x), dd_mul_d(e, DD_MUL(x, 0.5)));
}

DD_FN bool dd_lt(dd a, dd b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
DD_FN bool dd_gt(dd a, dd b) { return dd_lt(b, a); }
DD_FN bool dd_le_zero(dd a) { return !(a.hi > 0.0 || (a.hi == 0.0 && a.lo > 0.0)); }
DD_FN dd dd_min(dd a, dd b) { return dd_lt(b, a) ? b : a; }
DD_FN dd dd_max(dd a, dd b) { return dd_lt(a, b) ? b : a; }
DD_FN dd dd_abs(dd a) { return (a.hi < 0.0 || (a.hi == 0.0 && a.lo < 0.0)) ? dd_neg(a) : a; }

#ifdef __CUDACC__
__device__ __forceinline__ dd dd_shfl_xor(dd v, int o) {
    return dd_make(__shfl_xor_sync(0xffffffffu, v.hi, o), __shfl_xor_sync(0xffffffffu, v.lo, o));
}
__device__ __forceinline__ dd dd_shfl(dd v, int src) {
    return dd_make(__shfl_sync(0xffffffffu, v.hi, src), __shfl_sync(0xffffffffu, v.lo, src));
}
__device__ __forceinline__ dd dd_warp_sum(dd v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dd_add(v, dd_shfl_xor(v, o));
    return v;
}
__device__ __forceinline__ dd dd_warp_min(dd v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dd_min(v, dd_shfl_xor(v, o));
    return v;
}
#endif

}  // namespace lrn
