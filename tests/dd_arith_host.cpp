// CPU unit test of loraine.jl_b200/csrc/dd.cuh (host instantiation of the same code the kernels use): every double-double
// operation against __float128 (113-bit significand) on random operands of mixed magnitudes.
// build: g++ -O2 -ffp-contract=off -I/usr/local/cuda/include tests/dd_arith_host.cpp -lquadmath
#include "../loraine.jl_b200/csrc/dd.cuh"
#include <quadmath.h>
#include <cstdio>
#include <cstdlib>
#include <random>

using lrn::dd;
typedef __float128 q;

static q val(dd a) { return (q)a.hi + (q)a.lo; }
static double relerr(dd a, q ref) {
    if (ref == 0) return (double)fabsq(val(a));
    return (double)fabsq((val(a) - ref) / ref);
}

int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::uniform_int_distribution<int> ex(-40, 40);
    double worst[7] = {0, 0, 0, 0, 0, 0, 0};
    const char* name[7] = {"add", "sub", "mul", "div", "sqrt", "fma", "rsqrt"};
    for (int t = 0; t < 2000000; t++) {
        auto rnd = [&]() {
            const double hi = std::ldexp(u(rng), ex(rng));
            const dd r = lrn::quick_two_sum(hi, hi * u(rng) * 0x1p-53);
            return r;
        };
        const dd a = rnd(), b = rnd(), c = rnd();
        const q qa = val(a), qb = val(b), qc = val(c);
        double e;
        // a + b can cancel: bound the error relative to |a| + |b| as the algorithm guarantees
        e = (double)fabsq((val(lrn::dd_add(a, b)) - (qa + qb)) / (fabsq(qa) + fabsq(qb))); if (e > worst[0]) worst[0] = e;
        e = (double)fabsq((val(lrn::dd_sub(a, b)) - (qa - qb)) / (fabsq(qa) + fabsq(qb))); if (e > worst[1]) worst[1] = e;
        e = relerr(lrn::dd_mul(a, b), qa * qb); if (e > worst[2]) worst[2] = e;
        if (b.hi != 0.0) { e = relerr(lrn::dd_div(a, b), qa / qb); if (e > worst[3]) worst[3] = e; }
        const dd p = lrn::dd_abs(a);
        e = relerr(lrn::dd_sqrt(p), sqrtq(val(p))); if (e > worst[4]) worst[4] = e;
        if (val(p) > 0) { e = relerr(lrn::dd_rsqrt(p), 1 / sqrtq(val(p))); if (e > worst[6]) worst[6] = e; }
        e = (double)fabsq((val(lrn::dd_fma(a, b, c)) - (qa * qb + qc)) / (fabsq(qa * qb) + fabsq(qc))); if (e > worst[5]) worst[5] = e;
    }
    int bad = 0;
    const double tol = 0x1p-100;      // 2^-100 ~ 7.9e-31 (double-double: ~2^-104 per operation, a few ulps for div / sqrt)
    for (int k = 0; k < 7; k++) {
        printf("%s %.3e\n", name[k], worst[k]);
        if (!(worst[k] <= tol)) bad++;
    }
    // comparisons and the error-free transformations on a case that plain doubles get wrong
    const dd one = lrn::dd_make(1.0), tiny = lrn::dd_make(0x1p-80);
    const dd s = lrn::dd_add(one, tiny);
    if (!(lrn::dd_gt(s, one) && lrn::dd_lt(one, s) && val(lrn::dd_sub(s, one)) == (q)0x1p-80)) { printf("compare FAILED\n"); bad++; }
    if (!lrn::dd_le_zero(lrn::dd_make(0.0)) || lrn::dd_le_zero(tiny) || !lrn::dd_le_zero(lrn::dd_neg(tiny))) { printf("sign FAILED\n"); bad++; }
    printf(bad ? "FAILED\n" : "OK\n");
    return bad;
}
