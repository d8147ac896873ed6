// kfbo_device.cuh -- device-side building blocks of the Monte Carlo cost evaluation (sm_100a).
//
// One trajectory (trial x candidate x sample time) per thread; truth state, estimate, covariance
// and the NEES/NIS accumulators live in registers; all arithmetic is FP64.
//
// Reference semantics restated (file:line relative to the reference tree):
//   truth step      simulator.cpp:15-19, include/system_models.hpp:80-100,152-162
//   filter step     kalman_filter.cpp:18-38
//   NEES / NIS      trial.cpp:53-66
//   time variance   trial.cpp:84-98 (two-pass there; shifted one-pass sums here)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kfbo_kernels.h"
#include "kfbo_math.cuh"

namespace kfbo {

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter-based: no state, no races -- replaces the
// process-wide static std::mt19937 of include/eigenmvn.h:51,62.
//
// Cost on B200 (tools/microbench/imad_wide.cu, pipe_mix.cu): the 32x32->64 multiply IMAD.WIDE.U32
// takes ~5.1 cycles per warp, ~75% of which do NOT overlap with DFMA (8 DFMA 17.6 cycles, 8 WIDE 40.7,
// both 53.9); IMAD.HI.U32 ~4.1 (half exclusive); the 32-bit IMAD takes 2 cycles and overlaps fully.
// One Philox call is up to 20 such multiplies ~ 80-100 pipe cycles, as much as 45 DFMAs -- random bits
// are the expensive resource of this kernel, which is why a normal pair consumes 64 of them, not 128.
// ---------------------------------------------------------------------------------------
// KFBO_PHILOX_SPLIT16_ROUNDS: how many of the 10 rounds form their two 32x32->64 products from four
// 16x16->32 products (IMAD on the FMA pipe + shifts/adds on the ALU pipe, ~11 instructions that all
// overlap with FP64) instead of one IMAD.WIDE (1 instruction, ~5 cycles, ~75% of them exclusive with
// FP64 -- tools/microbench/imad_wide.cu).  Trades issue slots for shared-pipe cycles.
#ifndef KFBO_PHILOX_SPLIT16_ROUNDS
#define KFBO_PHILOX_SPLIT16_ROUNDS 0
#endif
// KFBO_PHILOX_SPLIT_MUL (experiment, off): keep the two halves of the product apart -- IMAD.HI + IMAD instead of one
// IMAD.WIDE.  ptxas fuses mul.hi + mul.lo of the same operands back into IMAD.WIDE, so the low half is formed as a
// multiply-add with an addend it cannot prove to be zero (`zero`: a value loaded from memory that is always 0).
#ifndef KFBO_PHILOX_SPLIT_MUL
#define KFBO_PHILOX_SPLIT_MUL 0
#endif
template <uint32_t M>
__device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t &hi, uint32_t &lo, uint32_t zero = 0u)
{
#if KFBO_PHILOX_SPLIT_MUL
    hi = __umulhi(M, a);
    lo = M * a + zero;
#else
    (void)zero;
    lo = M * a;
    hi = __umulhi(M, a);
#endif
}
template <uint32_t M>
__device__ __forceinline__ void mulhilo32_split16(uint32_t a, uint32_t &hi, uint32_t &lo)
{
    constexpr uint32_t Mh = M >> 16, Ml = M & 0xFFFFu;
    const uint32_t al = a & 0xFFFFu, ah = a >> 16;
    const uint32_t p0 = al * Ml;                       // < 2^32
    const uint32_t m1 = al * Mh + (p0 >> 16);          // <= (2^16-1)^2 + 2^16-1 < 2^32
    const uint32_t m2 = ah * Ml + (m1 & 0xFFFFu);      // < 2^32
    hi = ah * Mh + (m1 >> 16) + (m2 >> 16);
    lo = M * a;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &key,
                                               uint32_t zero = 0u)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t lo0, hi0, lo1, hi1;
        if (r >= 10 - KFBO_PHILOX_SPLIT16_ROUNDS) {
            mulhilo32_split16<0xD2511F53u>(c0, hi0, lo0);
            mulhilo32_split16<0xCD9E8D57u>(c2, hi1, lo1);
        } else {
            mulhilo32<0xD2511F53u>(c0, hi0, lo0, zero);
            mulhilo32<0xCD9E8D57u>(c2, hi1, lo1, zero);
        }
        c0 = hi1 ^ c1 ^ key.k0[r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ key.k1[r];
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Noise stream layout (restated draw for draw by oracle/kfbo_oracle.c):
//   key = (seed_lo, seed_hi); counter = (trial, block, stream[31:0], sub ^ (stream[61:32] << 2))   [KFBO_PHILOX_LAYOUT below]
//   (the stream id is 62 bits: sub < 4 leaves the upper 30 bits of its word free for the id's upper half, which enters the
//   first round through an XOR with a loop-invariant word -- no instruction in the step loop)
//   block 0, sub 0         -> words (x, y): the N(0,P0) draw added to x0
//   block g+1, sub 0, 1, 2 -> calls A, B, C of the STEP GROUP g = steps 4g .. 4g+3; a call (x, y, z, w) holds the two
//                             normal pairs (x, y) and (z, w) of math::normal_pair:
//        step 4g   : process (A.x, A.y)   observation: cos half of (C.x, C.y)
//        step 4g+1 : process (A.z, A.w)   observation: sin half of (C.x, C.y)
//        step 4g+2 : process (B.x, B.y)   observation: cos half of (C.z, C.w)
//        step 4g+3 : process (B.z, B.w)   observation: sin half of (C.z, C.w)
//   3 calls per 4 steps: all 12 x 32 = 384 bits are used, none is shared between two normals.
struct GroupWords { uint4 a, b, c; };

// KFBO_PHILOX_LAYOUT: where (trial, stream, block, sub) sit in the counter (c0, c1, c2, c3).  A Philox round multiplies c0
// and c2 and only XORs c1 and c3 in, so what the counter holds THERE decides how many of the 20 multiplies of a call are
// loop-invariant or shared by the three calls of a step group (the compiler hoists / merges them):
//   1 (the spec, which the CPU checker restates): (trial, block, stream, sub) -- both multiplied words are loop-invariant, block
//      and sub enter through the XORs: rounds 1-3 cost 2 multiplies per group, 44 per group in all;
//   0 (the layout until r1o, kept as a timing knob only -- the parity tests fail with it): (trial, stream, block, sub) --
//      block is multiplied in round 1 and the three calls diverge one round earlier: 48 per group.
#ifndef KFBO_PHILOX_LAYOUT
#define KFBO_PHILOX_LAYOUT 1
#endif
// stream: bits 0..31 of the stream id; stream_x: its bits 32..61 shifted left by 2 (Case::stream_x)
__device__ __forceinline__ uint4 philox_at(uint32_t trial, uint32_t stream, uint32_t stream_x, uint32_t block, uint32_t sub, const PhiloxKeys &key, uint32_t zero = 0u)
{
#if KFBO_PHILOX_LAYOUT == 1
    return philox4x32_10(trial, block, stream, sub ^ stream_x, key, zero);
#else
    return philox4x32_10(trial, stream, block, sub ^ stream_x, key, zero);
#endif
}

__device__ __forceinline__ GroupWords philox_group(uint32_t trial, uint32_t stream, uint32_t stream_x, uint32_t group, const PhiloxKeys &key, uint32_t zero = 0u)
{
    GroupWords w;
#if defined(KFBO_ABLATE_PHILOX)   // timing experiment only (the results are WRONG): twelve cheap words instead of three Philox calls
    {
        const uint32_t s = (trial ^ stream) + (group + 1u) * 0x9E3779B9u;
        w.a = make_uint4(s * 0xD2511F53u, (s + 1u) * 0xCD9E8D57u, (s + 2u) * 0xD2511F53u, (s + 3u) * 0xCD9E8D57u);
        w.b = make_uint4((s + 4u) * 0xD2511F53u, (s + 5u) * 0xCD9E8D57u, (s + 6u) * 0xD2511F53u, (s + 7u) * 0xCD9E8D57u);
        w.c = make_uint4((s + 8u) * 0xD2511F53u, (s + 9u) * 0xCD9E8D57u, (s + 10u) * 0xD2511F53u, (s + 11u) * 0xCD9E8D57u);
        return w;
    }
#endif
    w.a = philox_at(trial, stream, stream_x, group + 1u, 0u, key, zero);
    w.b = philox_at(trial, stream, stream_x, group + 1u, 1u, key, zero);
    w.c = philox_at(trial, stream, stream_x, group + 1u, 2u, key, zero);
    return w;
}

// 1/S and 1/det(P): S and det(P) come out of a covariance recursion and are normal, well-scaled
// numbers, so the division's slow path (denormal / huge operands) is never needed; one third-order
// step on the MUFU seed gives <= 1 ulp.  -DKFBO_EXACT_DIV restores the correctly rounded IEEE division.
#if defined(KFBO_EXACT_DIV)
#define KFBO_RCP(x) (1.0 / (x))
#else
#define KFBO_RCP(x) ::kfbo::math::rcp(x)
#endif

// The one-pass time variance of trial.cpp:84-98 is formed from SHIFTED sums, sum (v - k) and sum (v - k)^2; its cancellation
// error is eps (1 + (mean - k)^2 / variance).  Two shift policies (template parameter kShiftFirst of filter_step):
//   kShiftFirst = true : k = the trial's first sample -- within a few standard deviations of the mean whatever the
//       samples are, exact for T = 2.  Costs two registers, two predicated multiplies per step group, and makes the shifted
//       value a DFMA with three distinct vector-register sources (2.9 FP64-pipe cycles instead of 2.0, microbench/rf_ports.cu).
//   kShiftFirst = false: k = the value a consistent filter averages to (trial.cpp:77,93 compare against exactly these:
//       stateDOF = 2, observationDOF = 1) -- an immediate operand.  For chi-square-like samples (mean - k)^2 / variance is
//       O(1) as soon as a trial has more than a handful of steps; only for very short trials can the samples of one
//       trial coincide to many digits (T = 2: |a - b| << |a - k| with probability ~ that ratio), so the launchers take
//       this policy only when every case of the launch has T >= kMinStepsConstShift.  Measured: 13.37 -> 13.10 ms (cfg3).
constexpr int kMinStepsConstShift = 32;
constexpr double kShiftNees = 2.0, kShiftNis = 1.0;

// Per-trajectory register state.
struct Traj {
    double x0, x1;               // truth state            (Simulator::x_)
    double xh0, xh1;             // estimate               (KalmanFilter::xest_)
    double p00, p01, p11;        // covariance.  The reference carries all four entries and never
                                 // symmetrises; P(0,1) and P(1,0) are equal in exact arithmetic (Ppred is
                                 // symmetric and P = (I-KH)Ppred with K = Ppred H^T/S), they differ by
                                 // rounding only (~1e-16, contracting), far inside the 1e-9 parity bound.
    double kn, ki;               // the shift k of the sums below (constants unless KFBO_SHIFT_FIRST_SAMPLE)
    double sdn, sqn, sdi, sqi;   // sum (v-k), sum (v-k)^2 for NEES and NIS
};

template <bool kShiftFirst = true>
__device__ __forceinline__ void traj_init(Traj &t, double x0n0, double x0n1)
{
    t.x0 = 0.0 + x0n0;           // x0 = 0, P0 = I hard-coded (trial.cpp:31-32); simulator.cpp:12
    t.x1 = 0.0 + x0n1;
    t.xh0 = 0.0; t.xh1 = 0.0;    // kalman_filter.cpp:14-15
    t.p00 = 1.0; t.p01 = 0.0; t.p11 = 1.0;
    t.kn = kShiftFirst ? 0.0 : kShiftNees;
    t.ki = kShiftFirst ? 0.0 : kShiftNis;
    t.sdn = 0.0; t.sqn = 0.0; t.sdi = 0.0; t.sqi = 0.0;
}

// The noise of one filter-step, either as post-transform draws (supplied mode: w0, w1, v are added
// as they come, like system_models.hpp:87,96,158) or as standard normals with the truth transform
// (w = L n, v = sqrt(R) n2) folded into the state update's multiply-adds.
struct StepNoise { double a, b, c; };

// What the single-trial trace (trial.cpp:69-78) records of a step, beyond the trajectory state.
struct StepProbe { double nees, nis; };

// One filter-step: truth step, Kalman predict/update, NEES/NIS accumulation.
// gu = g*u[step] = (0.5 dt^2 u, dt u), rounded products exactly as the reference forms them.
// f = Fd(0,1) comes in as a uniform value (SampleTimeConsts), the candidate's Qd / R from the Case.
// kFused: noise = standard normals (n0, n1, n2), l00/l10/l11/sr = the truth transform.
template <bool kFirst, bool kFused, bool kShiftFirst = true>
__device__ __forceinline__ void filter_step(Traj &t, const Case &m, const double f, double2 gu, const StepNoise nz,
                                            const double l00 = 0.0, const double l10 = 0.0, const double l11 = 0.0,
                                            const double sr = 0.0, StepProbe *probe = nullptr, const bool first_rt = false)
{
    // ---- truth: x = Fd x + g u + w ; z = H x + v
    const double xa = fma(f, t.x1, t.x0) + gu.x;
    const double xb = t.x1 + gu.y;
    double z;
    if (kFused) {
        t.x0 = fma(l00, nz.a, xa);
        t.x1 = fma(l11, nz.b, fma(l10, nz.a, xb));
        z = fma(sr, nz.c, t.x0);
    } else {
        t.x0 = xa + nz.a;
        t.x1 = xb + nz.b;
        z = t.x0 + nz.c;
    }
    // ---- predict: xpred = Fd xest + g u ; Ppred = (Fd P) Fd^T + Qd  (symmetric: 3 entries)
    const double xp0 = fma(f, t.xh1, t.xh0) + gu.x;
    const double xp1 = t.xh1 + gu.y;
    const double fp01 = fma(f, t.p11, t.p01);
    const double pp00 = fma(fp01, f, fma(f, t.p01, t.p00)) + m.q00;
    const double pp01 = fp01 + m.q01;
    const double pp11 = t.p11 + m.q11;
    // ---- gain: c = Ppred H^T ; S = H c + R ; K = c * (1/S)
    const double s = pp00 + m.r_est;
    const double sinv = KFBO_RCP(s);      // sob_.inverse() of a 1x1 (kalman_filter.cpp:28)
    const double k0 = pp00 * sinv;
    const double k1 = pp01 * sinv;
    // ---- update: nu = z - H xpred ; xest = xpred + K nu ; P = (I - K H) Ppred
    const double nu = z - xp0;
    t.xh0 = fma(k0, nu, xp0);
    t.xh1 = fma(k1, nu, xp1);
    // (I - K H) Ppred with K = Ppred H^T / S: (1 - k0) pp00 = (R/S) pp00 = R k0 and (1 - k0) pp01 = R k1 -- the same
    // numbers without forming 1 - k0 (one FP64 instruction fewer, and no cancellation when k0 -> 1)
    t.p00 = m.r_est * k0;
    t.p01 = m.r_est * k1;
    t.p11 = fma(-k1, pp01, pp11);
    // ---- NEES = e^T P^-1 e = (e^T adj(P) e) / det(P) (the closed-form 2x2 inverse of trial.cpp:58 with
    //      the division applied once, to the quadratic form); NIS = (nu / S) nu
    const double e0 = t.x0 - t.xh0, e1 = t.x1 - t.xh1;
    const double det = fma(t.p00, t.p11, -(t.p01 * t.p01));
    const double invdet = KFBO_RCP(det);
    const double u0 = fma(t.p11, e0, -(t.p01 * e1));
    const double u1 = fma(t.p00, e1, -(t.p01 * e0));
    const double num = fma(e1, u1, e0 * u0);
    const double nu_s = nu * sinv;
    // ---- accumulate, shifted by the first sample of the trial (kn, ki): the shifted values
    // nees - kn and nis - ki come straight out of the last fused multiply-add of each form
    if (probe != nullptr) { probe->nees = num * invdet; probe->nis = nu_s * nu; }
    // first_rt: the same for callers that only know at run time whether this is the trial's first step (a uniform
    // predicate on two multiplies instead of a branch, so that a whole step group stays one basic block)
    double dn, di;
    if (kShiftFirst) {
        if (kFirst || first_rt) { t.kn = num * invdet; t.ki = nu_s * nu; }
        dn = fma(num, invdet, -t.kn);
        di = fma(nu_s, nu, -t.ki);
    } else {
        dn = fma(num, invdet, -kShiftNees);
        di = fma(nu_s, nu, -kShiftNis);
    }
    t.sdn += dn; t.sqn = fma(dn, dn, t.sqn);
    t.sdi += di; t.sqi = fma(di, di, t.sqi);
}

// {row_nees, row_nis, var_nees, var_nis} of trial.cpp:62-63,84-96 from the shifted sums.
__device__ __forceinline__ void traj_finish(const Traj &t, int T, double out[4])
{
    const double n = (double)T, inv_n = 1.0 / n, inv_nm1 = 1.0 / (double)(T - 1);
    out[0] = fma(n, t.kn, t.sdn);
    out[1] = fma(n, t.ki, t.sdi);
    out[2] = fma(-t.sdn * inv_n, t.sdn, t.sqn) * inv_nm1;
    out[3] = fma(-t.sdi * inv_n, t.sdi, t.sqi) * inv_nm1;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace kfbo
