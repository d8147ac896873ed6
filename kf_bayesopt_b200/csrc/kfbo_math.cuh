// kfbo_math.cuh -- branch-free FP64 math for the on-device noise generator and the filter step.
//
// The stock CUDA log / sqrt / sincospi / division each carry a slow-path branch (denormals,
// huge arguments, correct rounding) and materialise their polynomial coefficients through
// UMOV pairs; in the rollout loop that made ~420 issued instructions and ~190 FP64 instructions
// per filter-step (profiles/ncu_philox_r1_first.txt).  The arguments here are known to be well
// scaled (u1 in [2^-52, 1), S and det(P) of a covariance recursion, angle in one octant), so the
// special cases are dropped; the Box-Muller direction is the product of two tabulated unit vectors (a coarse and a
// fine 1024-entry table: 2^20 equally spaced directions) instead of sin/cos polynomials.
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

// KFBO_SQRT_PRODUCT_FORM: the last step of sqrt_pos as g * (1 + r t) (1) or as g + (g r) t (0).  (1) has no DFMA with three
//   distinct vector-register sources (2.9 pipe cycles instead of 2.0, tools/microbench/rf_ports.cu) and costs 1 ulp.
// KFBO_LOG_DEGREE: 4 = the Taylor polynomial of -2 ln(1+r)/r to r^4 (<= 1.7 ulp); 30 = degree 3 with the r^4 term folded
//   into the lower coefficients (one DFMA fewer per normal pair; <= 15 ulp of L, i.e. <= 3e-18 absolute where L is small).
// Measured on the cfg3 evaluation together with KFBO_SHIFT_FIRST_SAMPLE = 0 (kfbo_device.cuh): 13.37 -> 12.99 ms.
#ifndef KFBO_SQRT_PRODUCT_FORM
#define KFBO_SQRT_PRODUCT_FORM 1
#endif
#ifndef KFBO_LOG_DEGREE
#define KFBO_LOG_DEGREE 30
#endif

// KFBO_EXPERIMENT_NO_BANK_CONFLICTS (timing experiment only -- the results are WRONG): the low three bits of every table
// index are replaced by the lane number, so the 8 lanes of a quarter-warp read 8 different 16-byte bank groups and a
// lookup is 4 shared-memory wavefronts instead of ~10.  Measures what conflict-free tables could buy before building them.
#ifndef KFBO_EXPERIMENT_NO_BANK_CONFLICTS
#define KFBO_EXPERIMENT_NO_BANK_CONFLICTS 0
#endif
#if defined(__CUDA_ARCH__) && KFBO_EXPERIMENT_NO_BANK_CONFLICTS
#define KFBO_TABLE_INDEX(i) (((i) & ~7) | (int)(threadIdx.x & 7u))
#else
#define KFBO_TABLE_INDEX(i) (i)
#endif

namespace kfbo {
namespace math {

constexpr int kLogTableBits = 10;
constexpr int kLogTableSize = 1 << kLogTableBits;   // 1024 entries x {invc, 2 ln(invc)}: |r| <= 2^-11, one polynomial term fewer than 256 entries need
constexpr uint32_t kLogOffHi = 0x3FE60000u;         // z in [0.6875, 1.375)
constexpr int kLogBelowOne = 160 << (kLogTableBits - 8);   // sub-intervals of [0.6875, 1): (1 - 0.6875) 2^(bits+1)
constexpr int kLogUnitEntry = kLogBelowOne - 1;     // the sub-interval [1-2^-(bits+1), 1): c = 1 exactly

constexpr int kDirTableBits = 10;
constexpr int kDirTableSize = 1 << kDirTableBits;   // per level: 1024 unit vectors {cos, sin}; two levels -> 2^20 directions
constexpr int kDirBits = 2 * kDirTableBits;         // direction index bits of a normal pair
constexpr int kRadiusBits = 64 - kDirBits;          // radius bits of a normal pair (44)

struct LogTables {
    double invc_logc[kLogTableSize][2];  // {1/c_i rounded, 2 ln(that rounded value)}
    double dir_coarse[kDirTableSize][2]; // {cos, sin} of 2 pi i / 2^10
    double dir_fine[kDirTableSize][2];   // {cos, sin} of 2 pi (j + 1/2) / 2^20
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
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int i = 0; i < kDirTableSize; ++i) {
        const long double a = two_pi * i / kDirTableSize, b = two_pi * (i + 0.5L) / ((long double)kDirTableSize * kDirTableSize);
        t->dir_coarse[i][0] = (double)cosl(a);
        t->dir_coarse[i][1] = (double)sinl(a);
        t->dir_fine[i][0] = (double)cosl(b);
        t->dir_fine[i][1] = (double)sinl(b);
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

// sqrt(a) to <= 2 ulp for normal a > 0: g = a y0 = sqrt(a)(1 - e), r = 1/2 - g y0/2 = e - e^2/2, and
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
#if KFBO_SQRT_PRODUCT_FORM
    return g * fma(r, t, 1.0);            // no DFMA with three distinct vector-register sources (3 pipe cycles); <= 2 ulp
#else
    return fma(g * r, t, g);
#endif
}

#if defined(__CUDACC__)
#define KFBO_CONST static __constant__
#else
#define KFBO_CONST static const
#endif

// -2 ln(1+r)/r = -2 + r - (2/3) r^2 + (2/4) r^3 - (2/5) r^4 + ... ; |r| <= h = 2^-11 (1024 sub-intervals): the first
// dropped term (1/3) r^5 is < 5e-18 of the leading -2.
// Degree 4 (KFBO_LOG_DEGREE 4): -2, 1, 1/2 are immediates, the hi32-rounded -2/5 too (its 4e-7 relative rounding is
// scaled by r^4 < 6e-14); -2/3 needs full precision.
KFBO_CONST double kLogPolyU[1] = {-2.0 / 3.0};
#define KFBO_LOG_C4 -0x1.9999900000000p-2
// Degree 3 (KFBO_LOG_DEGREE 30, the default): Chebyshev economisation of the r^4 term.  On |r| <= h,
// r^4 = h^2 r^2 - h^4/8 + (h^4/8) T4(r/h), so -(2/5) r^4 ~ -(2/5) h^2 r^2 + (1/20) h^4 with |error| <= h^4/20 = 2.8e-15
// of the leading -2 -- 1.4e-15 relative on r * poly, which is the whole of L only next to u = 1 (L < 2^-10).
// What matters downstream is the ABSOLUTE error of a normal draw, rad * (cos, sin) with rad = sqrt(L): 3e-15 for |n| up
// to 7.9 in either form (tests/native/math_check.cpp), unchanged.
KFBO_CONST double kLogPolyE[2] = {-2.0 / 3.0 - 0.4 * 0x1p-22, -2.0 + 0.05 * 0x1p-44};
// L = -2 ln(u) for u in [2^-52, 1): table-driven, no division, no branch.
//   u = 2^k z, z in [0.6875, 1.375);  i = top 10 bits of (bits(z) - OFF);  r = z/c_i - 1, |r| <= 2^-11
//   -2 ln u = 2 (-k) ln2 + 2 ln(invc_i) + r * (-2 ln(1+r)/r)
// tab may live in shared memory (device) or anywhere (host).
KFBO_HD double neg2_log(double u, const double (*tab)[2])
{
    const uint32_t hi = hi_word(u), lo = lo_word(u);
    const int32_t tmp = (int32_t)(hi - kLogOffHi);
    const int i = (tmp >> (20 - kLogTableBits)) & (kLogTableSize - 1);
    const int k = tmp >> 20;                                   // arithmetic: k in [-52, 0]
    const double z = make_double(hi - ((uint32_t)tmp & 0xFFF00000u), lo);
    const double invc = tab[KFBO_TABLE_INDEX(i)][0], logc2 = tab[KFBO_TABLE_INDEX(i)][1];
    const double r = fma(z, invc, -1.0);
#if KFBO_LOG_DEGREE == 4
    double s = fma(KFBO_LOG_C4, r, 0.5);
    s = fma(s, r, kLogPolyU[0]);
#else   // 30: degree 3 with the dropped r^4 term folded into the r^2 and constant coefficients (Chebyshev economisation)
    double s = fma(0.5, r, kLogPolyE[0]);
#endif
    s = fma(s, r, 1.0);
#if KFBO_LOG_DEGREE == 30
    s = fma(s, r, kLogPolyE[1]);
#else
    s = fma(s, r, -2.0);
#endif
    // the exponent's share -2 k ln 2 by one conversion + multiply-add (k in [-52, 0] is exact in FP64); a 53-entry table
    // would cost a shared-memory lookup, and shared-memory wavefronts are the scarcer resource (measured: 13.42 vs 13.59 ms)
#if defined(__CUDA_ARCH__)
    const double kd = __int2double_rn(k);
#else
    const double kd = (double)k;
#endif
    return fma(r, s, fma(kd, -0x1.62e42fefa39efp+0 /* -2 ln 2 */, logc2));
}

// rad * (cos, sin) of direction `dir` out of 2^20: theta = 2 pi (dir + 1/2) / 2^20 = 2 pi i / 2^10 + 2 pi (j + 1/2) / 2^20
// with dir = i 2^10 + j -- the product of a coarse and a fine tabulated unit vector: 6 FP64 instructions, no polynomial
// (13 for a table + rotation by a polynomial sin/cos; ~40 for a libdevice sincospi).  Absolute error <= 2.5e-16 rad.
KFBO_HD void rad_direction(double rad, uint32_t dir, const LogTables &tab, double &c, double &s)
{
    const uint32_t i = (dir >> kDirTableBits) & (kDirTableSize - 1u), j = dir & (kDirTableSize - 1u);
    const double a = rad * tab.dir_coarse[KFBO_TABLE_INDEX(i)][0], b = rad * tab.dir_coarse[KFBO_TABLE_INDEX(i)][1];
    const double c2 = tab.dir_fine[KFBO_TABLE_INDEX(j)][0], s2 = tab.dir_fine[KFBO_TABLE_INDEX(j)][1];
    c = fma(a, c2, -(b * s2));
    s = fma(b, c2, a * s2);
}

// 64 random bits (w0, w1) -> two independent N(0,1) draws (Box-Muller).  THE SPEC (restated draw for draw by oracle/kfbo_oracle.c):
//   radius   : 44 bits  m = (w0 << 12) | (w1 >> 20);  u1 = 1 - (m + 1/2) 2^-44  in [2^-45, 1 - 2^-45];  rad = sqrt(-2 ln u1)
//   direction: 20 bits  dir = w1 & 0xFFFFF;           theta = 2 pi (dir + 1/2) / 2^20
//   na = rad cos(theta), nb = rad sin(theta)
// Why these widths.  Random bits are the expensive resource on this chip (the 32x32->64 multiplies of Philox share
// the FP64 issue resource -- kfbo_device.cuh), so a pair takes 64 of them: one Philox4x32-10 call makes two pairs.
//   * 2^20 equally spaced directions: for a uniform grid of N angles every trigonometric moment E[cos^p sin^q] with
//     p + q < N equals the continuous one exactly, so all joint moments of (na, nb) up to order 2^20 - 1 are those of
//     two independent normals, and the marginal densities differ from the normal density by the (super-algebraically
//     small) trapezoid-rule error of a smooth periodic integrand;
//   * 44-bit radius variate: each normal is continuous (the radius is) with tails to 7.9 sigma; the mass beyond,
//     3e-15 per draw, is less than one draw in 10^14.
KFBO_HD void normal_pair(uint32_t w0, uint32_t w1, const LogTables &tab, double &na, double &nb)
{
#if defined(KFBO_ABLATE_NORMAL)   // timing experiment only (the results are WRONG): two uniforms on [-1/2, 1/2), no table, no log / sqrt
    na = make_double(0x3FF00000u | (w0 >> 12), w1) - 1.5;
    nb = make_double(0x3FF00000u | (w1 >> 12), w0) - 1.5;
    (void)tab;
    return;
#endif
    // y1 = 1 + (m + 1/2) 2^-44 in [1, 2): mantissa = w0 : w1[31:20] : 1000 0000b
    const double y1 = make_double(0x3FF00000u | (w0 >> 12), (funnel_l(w1, w0, 20) & 0xFFFFFF00u) | 0x80u);
    const double u1 = 2.0 - y1;
    const double rad = sqrt_pos(neg2_log(u1, tab.invc_logc));
    rad_direction(rad, w1, tab, na, nb);
}

}  // namespace math
}  // namespace kfbo
