// kfbo_math.cuh -- branch-free FP64 math for the on-device noise generator and the filter step.
//
// The stock CUDA log / sqrt / sincospi / division each carry a slow-path branch (denormals,
// huge arguments, correct rounding) and materialise their polynomial coefficients through
// UMOV pairs; in the rollout loop that made ~420 issued instructions and ~190 FP64 instructions
// per filter-step (profiles/ncu_philox_r1_first.txt).  The arguments here are known to be well
// scaled (u1 in [2^-52, 1), S and det(P) of a covariance recursion, angle in one octant), so the
// special cases are dropped.
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

constexpr int kLogTableBits = 8;
constexpr int kLogTableSize = 1 << kLogTableBits;   // 256 entries x {invc, 2 ln(invc)}
constexpr int kLogKTableSize = 56;                  // j = 0..52 used: 2 j ln2
constexpr uint32_t kLogOffHi = 0x3FE60000u;         // z in [0.6875, 1.375)
constexpr int kLogUnitEntry = 159;                  // the sub-interval [1-2^-9, 1): c = 1 exactly

struct LogTables {
    double invc_logc[kLogTableSize][2];  // {1/c_i rounded, 2 ln(that rounded value)}
    double k2ln2[kLogKTableSize];        // 2 j ln 2
};

// Host: build the tables in long double (x87 extended, 64-bit mantissa).
inline void build_log_tables(LogTables *t)
{
    for (int i = 0; i < kLogTableSize; ++i) {
        // sub-interval i of tmp = bits(z) - OFF: [i, i+1) * 2^44; below z = 1 one unit of tmp is
        // 2^-53 in z, above it 2^-52 (z = 1 sits exactly at i = 160)
        const long double lo = 0.6875L, w_lo = ldexpl(1.0L, -9), w_hi = ldexpl(1.0L, -8);
        const long double centre = (i < 160) ? lo + (i + 0.5L) * w_lo : 1.0L + (i - 160 + 0.5L) * w_hi;
        double invc = (double)(1.0L / centre);
        if (i == kLogUnitEntry) invc = 1.0;
        t->invc_logc[i][0] = invc;
        t->invc_logc[i][1] = (i == kLogUnitEntry) ? 0.0 : (double)(2.0L * logl((long double)invc));
    }
    for (int j = 0; j < kLogKTableSize; ++j) t->k2ln2[j] = (double)(2.0L * j * 0.693147180559945309417232121458176568L);
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
    int e;
    const double m = std::frexp(x, &e);
    return std::ldexp((double)(1.0f / (float)m), -e);
#endif
}

KFBO_HD double rsqrt_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    int e;
    double m = std::frexp(x, &e);
    if (e & 1) { m *= 2.0; --e; }
    return std::ldexp((double)(1.0f / std::sqrt((float)m)), -e / 2);
#endif
}

// 1/x to <= 1 ulp for normal, well-scaled x (two Newton steps on the hardware seed).
KFBO_HD double rcp(double x)
{
    double y = rcp_seed(x);
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// sqrt(a) to <= 1 ulp for normal a > 0 (coupled Goldschmidt iterations on the hardware seed).
KFBO_HD double sqrt_pos(double a)
{
    const double y = rsqrt_seed(a);
    double g = a * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    return fma(g, r, g);
}

#if defined(__CUDACC__)
#define KFBO_CONST static __constant__
#else
#define KFBO_CONST static const
#endif

// -2 ln(1+r)/r = -2 + r - (2/3) r^2 + (2/4) r^3 - (2/5) r^4 + (2/6) r^5 ; |r| <= 2^-9: rel. err 8e-18.
// -2, 1, 1/2 and the hi32-rounded 1/3 are immediates; -2/3 and -2/5 need full precision.
KFBO_CONST double kLogPolyU[2] = {-2.0 / 3.0, -0.4};
#define KFBO_LOG_C5 0x1.5555500000000p-2
// sin(pi/4 x) = x (S0 + S1 x^2 + ... + S7 x^14), cos(pi/4 x) = C0 + C1 x^2 + ... + C8 x^16, x in [-1,1]:
// Taylor coefficients with the powers of pi/4 folded in; S6, S7, C6, C7, C8 rounded to hi32
// immediates (max relative error of the polynomials in exact arithmetic: 1.2e-16 / 6.0e-17).
KFBO_CONST double kSinPolyU[6] = {0x1.921fb54442d18p-1,  -0x1.4abbce625be53p-4, 0x1.466bc6775aae2p-9,
                                  -0x1.32d2cce62bd86p-15, 0x1.50783487ee782p-22, -0x1.e3074fde8871fp-30};
#define KFBO_SIN_C6 0x1.e8f4300000000p-38
#define KFBO_SIN_C7 -0x1.6fadc00000000p-46
KFBO_CONST double kCosPolyU[5] = {-0x1.3bd3cc9be45dep-2, 0x1.03c1f081b5ac4p-6, -0x1.55d3c7e3cbffap-12,
                                  0x1.e1f506891babbp-19, -0x1.a6d1f2a204a8cp-26};
#define KFBO_COS_C6 0x1.f9d3900000000p-34
#define KFBO_COS_C7 -0x1.b6e2500000000p-42
#define KFBO_COS_C8 0x1.20c6300000000p-50

// L = -2 ln(u) for u in [2^-52, 1): table-driven, no division, no branch.
//   u = 2^k z, z in [0.6875, 1.375);  i = top 8 bits of (bits(z) - OFF);  r = z/c_i - 1, |r| <= 2^-9
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
    double s = fma(KFBO_LOG_C5, r, kLogPolyU[1]);
    s = fma(s, r, 0.5);
    s = fma(s, r, kLogPolyU[0]);
    s = fma(s, r, 1.0);
    s = fma(s, r, -2.0);
    return fma(r, s, ktab[-k] + logc2);
}

// sin(pi/4 x), cos(pi/4 x) for x in [-1, 1].
KFBO_HD void sincos_octant(double x, double &s, double &c)
{
    const double x2 = x * x;
    double ps = fma(KFBO_SIN_C7, x2, KFBO_SIN_C6);
    double pc = fma(KFBO_COS_C8, x2, KFBO_COS_C7);
    pc = fma(pc, x2, KFBO_COS_C6);
    ps = fma(ps, x2, kSinPolyU[5]); pc = fma(pc, x2, kCosPolyU[4]);
    ps = fma(ps, x2, kSinPolyU[4]); pc = fma(pc, x2, kCosPolyU[3]);
    ps = fma(ps, x2, kSinPolyU[3]); pc = fma(pc, x2, kCosPolyU[2]);
    ps = fma(ps, x2, kSinPolyU[2]); pc = fma(pc, x2, kCosPolyU[1]);
    ps = fma(ps, x2, kSinPolyU[1]); pc = fma(pc, x2, kCosPolyU[0]);
    ps = fma(ps, x2, kSinPolyU[0]); pc = fma(pc, x2, 1.0);
    s = ps * x;
    c = pc;
}

// 84 random bits -> two independent N(0,1) draws (Box-Muller).  THE SPEC (restated draw for draw
// by the CPU checker):
//   radius: m1 = (rad_hi << 20 | rad_lo20) with the lowest bit forced to 1 (52 bits);
//           u1 = 1 - m1 2^-52 in [2^-52, 1 - 2^-52]; rad = sqrt(-2 ln u1)
//   angle : o = top 3 bits of ang (octant), f = low 29 bits of ang times 2^-29 in [0,1);
//           theta = (pi/4)(o + f)
//   na = rad cos(theta), nb = rad sin(theta)
// (32 angle bits: 2^32 directions, 1.5e-9 rad apart; the radius keeps 52 bits, so each normal is
// continuous to FP64 resolution and the tails reach 8.5 sigma.  Random bits are the expensive part
// on this chip -- see the Philox note in kfbo_device.cuh.)
// Evaluated per octant: x = f (even o) or 1-f (odd o); sin/cos of (pi/4)x; swap and signs from o.
KFBO_HD void normal_pair_bits(uint32_t rad_hi, uint32_t rad_lo20, uint32_t ang, const double (*tab)[2],
                              const double *ktab, double &na, double &nb)
{
    // ---- radius
    const double y1 = make_double(0x3FF00000u | (rad_hi >> 12), (rad_hi << 20) | rad_lo20 | 1u);   // [1,2)
    const double u1 = 2.0 - y1;
    double rad = sqrt_pos(neg2_log(u1, tab, ktab));
    // ---- angle
    const uint32_t o = ang >> 29;
    const double y2 = make_double(0x3FF00000u | ((ang >> 9) & 0xFFFFFu), ang << 23);   // 1 + f
    const uint32_t odd = o & 1u;
    const uint32_t swap = (o ^ (o >> 1)) & 1u;             // sin(theta) takes the cos polynomial
    const uint32_t neg_sin = (o >> 2) & 1u;                // quadrants 2,3
    const uint32_t neg_cos = ((o >> 1) ^ (o >> 2)) & 1u;   // quadrants 1,2
    // sign bookkeeping: the cos-polynomial value gets its sign through rad, the sin-polynomial
    // value (odd in x) through x.  Which of sin/cos(theta) the polynomials feed depends on swap.
    const uint32_t neg_pc = swap ? neg_sin : neg_cos;      // sign wanted on the cos-polynomial value
    const uint32_t neg_ps = swap ? neg_cos : neg_sin;      // sign wanted on the sin-polynomial value
    rad = make_double(hi_word(rad) ^ (neg_pc << 31), lo_word(rad));
    const uint32_t flip_x = neg_ps ^ neg_pc;               // rad already carries neg_pc
    // t = f (even octant) or f - 1 (odd); |x| = f or 1 - f = -t; then the sign flip
    const double t = y2 + make_double(odd ? 0xC0000000u : 0xBFF00000u, 0u);     // y2 - 2 or y2 - 1
    const double x = make_double(hi_word(t) ^ ((odd ^ flip_x) << 31), lo_word(t));
    double ps, pc;
    sincos_octant(x, ps, pc);
    const double a = rad * pc, b = rad * ps;
    na = swap ? b : a;     // rad cos(theta)
    nb = swap ? a : b;     // rad sin(theta)
}

}  // namespace math
}  // namespace kfbo
