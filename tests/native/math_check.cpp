// math_check.cpp -- CPU accuracy check of the product's branch-free FP64 math
// (kf_bayesopt_b200/csrc/kfbo_math.cuh, compiled here as plain host C++) against long-double libm.
// Prints "name max_ulp_or_rel_err" lines; tests/test_math_host.py parses and bounds them.
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <random>

#include "../../kf_bayesopt_b200/csrc/kfbo_math.cuh"

using namespace kfbo::math;

static double ulp_err(double got, long double want)
{
    if (want == 0.0L) return got == 0.0 ? 0.0 : 1e30;
    int e;
    frexpl(want, &e);
    const long double ulp = ldexpl(1.0L, e - 53);
    return (double)(fabsl((long double)got - want) / ulp);
}

int main()
{
    static LogTables t;
    build_log_tables(&t);
    std::mt19937_64 g(12345);
    double e_rcp = 0, e_sqrt = 0, e_log = 0, e_sin = 0, e_cos = 0, e_na = 0, e_nb = 0;
    // rcp / sqrt over 40 decades
    for (int i = 0; i < 2000000; ++i) {
        const double m = 1.0 + (double)(g() >> 11) * 0x1p-53;
        const double x = std::ldexp(m, (int)(g() % 140) - 70);
        e_rcp = std::fmax(e_rcp, ulp_err(rcp(x), 1.0L / x));
        e_rcp = std::fmax(e_rcp, ulp_err(rcp(-x), -1.0L / x));
        e_sqrt = std::fmax(e_sqrt, ulp_err(sqrt_pos(x), sqrtl((long double)x)));
    }
    // -2 ln u over the whole domain [2^-52, 1), with emphasis on both ends
    for (int i = 0; i < 4000000; ++i) {
        uint64_t m1 = (g() >> 12) | 1u;
        if (i % 4 == 1) m1 >>= (g() % 52);          // u near 1
        if (i % 4 == 2) m1 |= ~0ull << (g() % 52) >> 12 << 0, m1 &= (1ull << 52) - 1, m1 |= 1;  // u near 2^-52
        const double u = 1.0 - (double)m1 * 0x1p-52;
        if (!(u >= 0x1p-52 && u < 1.0)) continue;
        e_log = std::fmax(e_log, ulp_err(neg2_log(u, t.invc_logc), -2.0L * logl((long double)u)));
    }
    // edge values
    for (double u : {0x1p-52, 1.0 - 0x1p-52, 0.6875, 0.6875 - 0x1p-54, 1.0 - 0x1p-9, 1.0 - 0x1p-9 - 0x1p-53, 0.5, 0.25 + 0x1p-54})
        e_log = std::fmax(e_log, ulp_err(neg2_log(u, t.invc_logc), -2.0L * logl((long double)u)));
    // rad * (cos, sin) of one of 2^20 directions (coarse x fine table product), rad = 1: absolute error in ulps of 1
    const long double TWO_PI = 6.283185307179586476925286766559005768L;
    for (int i = 0; i < 4000000; ++i) {
        uint32_t dir = (uint32_t)g() & 0xFFFFFu;
        if (i % 8 == 0) dir = (dir & ~0x3FFu) | ((i & 8) ? 0x3FFu : 0u);            // both ends of a coarse sub-interval
        if (i % 8 == 1) dir &= 0xC0000u;                                              // next to the axes
        double s, c;
        rad_direction(1.0, dir | ((uint32_t)g() << 20), t, c, s);                     // the bits above 20 must be ignored
        const long double th = TWO_PI * ((long double)dir + 0.5L) * 0x1p-20L;
        e_sin = std::fmax(e_sin, (double)(fabsl((long double)s - sinl(th)) * 0x1p53L));
        e_cos = std::fmax(e_cos, (double)(fabsl((long double)c - cosl(th)) * 0x1p53L));
    }
    // the full pair against the spec evaluated in long double (absolute error: |n| <= 7.9)
    for (int i = 0; i < 2000000; ++i) {
        uint32_t w0 = (uint32_t)g(), w1 = (uint32_t)g();
        if (i % 16 == 0) { w0 = 0xFFFFFFFFu; w1 |= 0xFFF00000u; }                     // the largest radius
        if (i % 16 == 1) { w0 = 0u; w1 &= 0x000FFFFFu; }                              // the smallest
        double na, nb;
        normal_pair(w0, w1, t, na, nb);
        const uint64_t m = ((uint64_t)w0 << 12) | (w1 >> 20);
        const long double u1 = 1.0L - ((long double)m + 0.5L) * 0x1p-44L;
        const long double theta = TWO_PI * ((long double)(w1 & 0xFFFFFu) + 0.5L) * 0x1p-20L;
        const long double rad = sqrtl(-2.0L * logl(u1));
        e_na = std::fmax(e_na, (double)fabsl((long double)na - rad * cosl(theta)));
        e_nb = std::fmax(e_nb, (double)fabsl((long double)nb - rad * sinl(theta)));
    }
    std::printf("rcp_ulp %.3f\nsqrt_ulp %.3f\nneg2log_ulp %.3f\nsin_ulp %.3f\ncos_ulp %.3f\nna_abs %.3e\nnb_abs %.3e\n",
                e_rcp, e_sqrt, e_log, e_sin, e_cos, e_na, e_nb);
    return 0;
}
