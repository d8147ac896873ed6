// kfbo_math.cuh -- branch-free FP64 math for the on-device noise generator and the filter step.
//
// The stock CUDA log / sqrt / sincospi / division each carry a slow-path branch (denormals,
// huge arguments, correct rounding) and materialise their polynomial coefficients through
// UMOV pairs; in the rollout loop that made ~420 issued instructions and ~190 FP64 instructions
// per filter-step (profiles/ncu_philox_r1_first.txt).  The arguments here are known to be well
// scaled (u1 in [2^-52, 1), S and det(P) of a covariance recursion, angle in one octant), so the
// special cases are dropped; sin/cos of the Box-Muller angle come from a 1024-entry direction table
// plus a two-term rotation instead of degree-16 polynomials.
//
// Measured on B200 (tools/microbench/, DESIGN.md "What the FP64 pipe really costs"): a DFMA whose
// three sources are three DISTINCT vector registers occupies the FP64 pipe for 3 cycles per warp
// instead of 2.  A Horner step fma(p, x2, C) is therefore only full-rate when C is NOT a vector
// register: the low-order coefficients live in __constant__ memory (loaded once into uniform
// registers), the high-order ones are rounded to doubles whose low 32 bits are zero, which the
// assembler encodes as immediates (the rounding moves the results by < 1e-16 relative).
//
// Under nvcc the functions are __device__; compiled as plain C++ (g++) the same source gives host
// functions, which is how tests/native/math_check.cpp checks the accuracy against long-double
// libm on the CPU (the MUFU seeds are emulated there with float precision).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define KFBO_HD __device__ __forceinline__
#else
#define KFBO_HD inline
#endif

namespace kfbo {
namespace math {

constexpr int kLogTableBits = 10;
constexpr int kLogTableSize = 1 << kLogTableBits;   // 1024 entries x {invc, 2 ln(invc)}: |r| <= 2^-11, one polynomial term fewer than 256 entries need
constexpr int kLogKTableSize = 56;                  // j = 0..52 used: 2 j ln2
constexpr uint32_t kLogOffHi = 0x3FE60000u;         // z in [0.6875, 1.375)
constexpr int kLogBelowOne = 160 << (kLogTableBits - 8);   // sub-intervals of [0.6875, 1): (1 - 0.6875) 2^(bits+1)
constexpr int kLogUnitEntry = kLogBelowOne - 1;     // the sub-interval [1-2^-(bits+1), 1): c = 1 exactly

constexpr int kAngleTableBits = 10;
constexpr int kAngleTableSize = 1 << kAngleTableBits;   // 1024 directions x {cos, sin}

struct LogTables {
    double invc_logc[kLogTableSize][2];  // {1/c_i rounded, 2 ln(that rounded value)}
    double k2ln2[kLogKTableSize];        // entry 52 + k: -2 k ln 2
    double cossin[kAngleTableSize][2];   // {cos, sin} of 2 pi (i + 1/2) / 1024
};

// Host: build the tables in long double (x87 extended, 64-bit mantissa).
inline void build_log_tables(LogTables *t)
{
    for (int i = 0; i < kLogTableSize; ++i) {
        // sub-interval i of tmp = bits(z) - OFF: [i, i+1) * 2^(52-bits); below z = 1 one unit of tmp is
        // 2^-53 in z, above it 2^-52 (z = 1 sits exactly at i = kLogBelowOne)
        const long double lo = 0.6875L, w_lo = ldexpl(1.0L, -(kLogTableBits + 1)), w_hi = ldexpl(1.0L, -kLogTableBits);
        const long double centre = (i < kLogBelowOne) ? lo + (i + 0.5L) * w_lo : 1.0L + (i - kLogBelowOne + 0.5L) * w_hi;
        double invc = (double)(1.0L / centre);
        if (i == kLogUnitEntry) invc = 1.0;
        t->invc_logc[i][0] = invc;
        t->invc_logc[i][1] = (i == kLogUnitEntry) ? 0.0 : (double)(2.0L * logl((long double)invc));
    }
    for (int j = 0; j < kLogKTableSize; ++j)   // indexed by 52 + k, k = exponent of u in [-52, 0]: a non-negative index
        t->k2ln2[j] = (double)(2.0L * (kLogKTableSize - 4 - j) * 0.693147180559945309417232121458176568L);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int i = 0; i < kAngleTableSize; ++i) {
        const long double a = two_pi * (i + 0.5L) / kAngleTableSize;
        t->cossin[i][0] = (double)cosl(a);
        t->cossin[i][1] = (double)sinl(a);
    }
}

KFBO_HD double make_double(uint32_t hi, uint32_t lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    std::memcpy(&d, &b, sizeof d);
    return d;
#endif
}

KFBO_HD uint32_t hi_word(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(d);
#else
    uint64_t b;
    std::memcpy(&b, &d, sizeof b);
    return (uint32_t)(b >> 32);
#endif
}

// (hi : lo) << n, upper word: SHF.L.W on the device (the ALU pipe -- a plain C shift may become an IMAD
// on the FMA-heavy pipe, which this kernel cannot spare).  0 < n < 32.
KFBO_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t n)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, n);
#else
    return (hi << n) | (lo >> (32u - n));
#endif
}

KFBO_HD uint32_t lo_word(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(d);
#else
    uint64_t b;
    std::memcpy(&b, &d, sizeof b);
    return (uint32_t)b;
#endif
}

// ~22-bit seeds: MUFU.RCP64H / MUFU.RSQ64H on the device.
KFBO_HD double rcp_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    // MUFU.RCP64H reads and writes only the high 32-bit word (20 mantissa bits): emulate that
    const double xt = make_double(hi_word(x), 0u);
    const double y = 1.0 / xt;
    return make_double(hi_word(y), 0u);
#endif
}

KFBO_HD double rsqrt_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    const double xt = make_double(hi_word(x), 0u);     // MUFU.RSQ64H: high word in, high word out
    const double y = 1.0 / std::sqrt(xt);
    return make_double(hi_word(y), 0u);
#endif
}

// 1/x to <= 1 ulp for normal, well-scaled x.  The hardware seed y0 = (1/x)(1 - e) is good to
// |e| < 2^-19 (it sees 20 mantissa bits), so ONE third-order step y0 (1 + e + e^2) leaves e^3 < 2^-57:
// 3 FP64 instructions instead of the 4 of two Newton steps.
KFBO_HD double rcp(double x)
{
    const double y = rcp_seed(x);
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}

// sqrt(a) to <= 1 ulp for normal a > 0: g = a y0 = sqrt(a)(1 - e), r = 1/2 - g y0/2 = e - e^2/2, and
// sqrt(a) = g (1 - 2r)^(-1/2) = g (1 + r + 3/2 r^2 + O(r^3)): one third-order step on the 20-bit seed
// (|r|^3 < 2^-57), 5 FP64 instructions instead of the 7 of two coupled Goldschmidt iterations.
KFBO_HD double sqrt_pos(double a)
{
    const double y = rsqrt_seed(a);
    // h = y / 2 exactly, by decrementing the exponent field (y = 1/sqrt(a) is a normal number far from the
    // subnormal range): one integer add instead of a multiply on the FP64 pipe
    const double g = a * y, h = make_double(hi_word(y) - 0x00100000u, lo_word(y));
    const double r = fma(-h, g, 0.5);
    const double t = fma(r, 1.5, 1.0);
    return fma(g * r, t, g);
}

#if defined(__CUDACC__)
#define KFBO_CONST static __constant__
#else
#define KFBO_CONST static const
#endif

// -2 ln(1+r)/r = -2 + r - (2/3) r^2 + (2/4) r^3 - (2/5) r^4 + ... ; |r| <= 2^-11 (1024 sub-intervals): the first
// dropped term (1/3) r^5 is < 5e-18 of the leading -2.
// -2, 1, 1/2 are immediates, the hi32-rounded -2/5 too (its 4e-7 relative rounding is scaled by r^4 < 6e-14);
// -2/3 needs full precision.
KFBO_CONST double kLogPolyU[1] = {-2.0 / 3.0};
#define KFBO_LOG_C4 -0x1.9999900000000p-2
// L = -2 ln(u) for u in [2^-52, 1): table-driven, no division, no branch.
//   u = 2^k z, z in [0.6875, 1.375);  i = top 10 bits of (bits(z) - OFF);  r = z/c_i - 1, |r| <= 2^-11
//   -2 ln u = 2 (-k) ln2 + 2 ln(invc_i) + r * (-2 ln(1+r)/r)
// tab / ktab may live in shared memory (device) or anywhere (host).
KFBO_HD double neg2_log(double u, const double (*tab)[2], const double *ktab)
{
    const uint32_t hi = hi_word(u), lo = lo_word(u);
    const int32_t tmp = (int32_t)(hi - kLogOffHi);
    const int i = (tmp >> (20 - kLogTableBits)) & (kLogTableSize - 1);
    const int k = tmp >> 20;                                   // arithmetic: k in [-52, 0]
    const double z = make_double(hi - ((uint32_t)tmp & 0xFFF00000u), lo);
    const double invc = tab[i][0], logc2 = tab[i][1];
    const double r = fma(z, invc, -1.0);
    double s = fma(KFBO_LOG_C4, r, 0.5);
    s = fma(s, r, kLogPolyU[0]);
    s = fma(s, r, 1.0);
    s = fma(s, r, -2.0);
#if defined(KFBO_LOG_K_I2F) && defined(__CUDA_ARCH__)
    (void)ktab;
    return fma(r, s, fma(__int2double_rn(k), -0x1.62e42fefa39efp+0 /* -2 ln 2 */, logc2));
#else
    return fma(r, s, ktab[kLogKTableSize - 4 + k] + logc2);    // entry j holds 2 (52 - j) ln 2, j = 52 + k >= 0
#endif
}

// cos/sin of theta = 2 pi ang / 2^32 for a 32-bit angle: the top 10 bits pick one of 1024 tabulated
// directions (cos, sin of the sub-interval's centre), the low 22 bits give the offset
// |d| <= pi/1024 from it; cos d and sin d need two polynomial terms each (d^6/720 < 2e-18,
// d^7/5040 < 6e-22) and the result is the rotation of the tabulated direction by d, scaled by `rad`:
//   c = rad cos(theta), s = rad sin(theta).   13 FP64 instructions, no octant logic.
KFBO_CONST double kAngU[4] = {0x1.921fb54442d18p-8 /* 2 pi / 1024 */, -0x1.2d97c7f3321d2p-7 /* -1.5 * that */,
                              1.0 / 24.0, 1.0 / 120.0};
KFBO_HD void rad_sincos_2pi(double rad, uint32_t ang, const double (*cossin)[2], double &c, double &s)
{
    const uint32_t i = ang >> (32 - kAngleTableBits), r = ang & ((1u << (32 - kAngleTableBits)) - 1u);
    const double y = make_double(0x3FF00000u | (r >> 2), funnel_l(0u, r, 30));   // 1 + r 2^-22 in [1, 2)
    const double d = fma(y, kAngU[0], kAngU[1]);                             // (2 pi / 1024)(r 2^-22 - 1/2)
    const double d2 = d * d;
    const double cd = fma(fma(d2, kAngU[2], -0.5), d2, 1.0);                 // cos d
    const double sp = fma(d2, kAngU[3], -1.0 / 6.0) * d2;
    const double sd = fma(sp, d, d);                                         // sin d
    const double a = rad * cossin[i][0], b = rad * cossin[i][1];
    c = fma(a, cd, -(b * sd));
    s = fma(b, cd, a * sd);
}

// 84 random bits -> two independent N(0,1) draws (Box-Muller).  THE SPEC (restated draw for draw
// by the CPU checker):
//   radius: m1 = (rad_hi << 20 | rad_lo20) with the lowest bit forced to 1 (52 bits), rad_lo20 = rad_lo_top20 >> 12;
//           u1 = 1 - m1 2^-52 in [2^-52, 1 - 2^-52]; rad = sqrt(-2 ln u1)
//   angle : o = top 3 bits of ang (octant), f = low 29 bits of ang times 2^-29 in [0,1);
//           theta = (pi/4)(o + f)
//   na = rad cos(theta), nb = rad sin(theta)
// (32 angle bits: 2^32 directions, 1.5e-9 rad apart; the radius keeps 52 bits, so each normal is
// continuous to FP64 resolution and the tails reach 8.5 sigma.  Random bits are the expensive part
// on this chip -- see the Philox note in kfbo_device.cuh.)
// theta = (pi/4)(o + f) = 2 pi ang / 2^32: evaluated by rad_sincos_2pi above.
// rad_lo_top20: a word whose TOP 20 bits are the low 20 bits of the radius mantissa (one funnel shift joins them).
KFBO_HD void normal_pair_bits(uint32_t rad_hi, uint32_t rad_lo_top20, uint32_t ang, const LogTables &tab, double &na, double &nb)
{
    const double y1 = make_double(0x3FF00000u | (rad_hi >> 12), funnel_l(rad_lo_top20, rad_hi, 20) | 1u);   // [1,2)
    const double u1 = 2.0 - y1;
    const double rad = sqrt_pos(neg2_log(u1, tab.invc_logc, tab.k2ln2));
    rad_sincos_2pi(rad, ang, tab.cossin, na, nb);
}

}  // namespace math
}  // namespace kfbo
