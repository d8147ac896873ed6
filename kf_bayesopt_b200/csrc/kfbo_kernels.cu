// kfbo_kernels.cu -- the rollout kernels (sm_100a) and their launchers.
//
//   rollout_philox    on-device Philox4x32-10 + Box-Muller noise (FP64-pipe bound)
//   rollout_supplied  noise streamed from HBM, SoA [step][component][trial] (HBM bound)
//
// Grid: (tiles, candidates, sample times): blockIdx.y/z give the case without any division, so the
// case and its sample-time constants are provably uniform (uniform registers / parameter bank).
// Reduction: warp shuffle -> shared -> one [4] partial per CTA -> the LAST CTA of each case
// (ticket counter) sums the partials in a fixed order, so sums are deterministic run to run.
#include "kfbo_device.cuh"
#include "kfbo_kernels.h"

namespace kfbo {

// Select entry k of a small per-sample-time table held in the parameter bank WITHOUT dynamic
// indexing: a register-indexed LDC lands in a vector register, a chain of selects on the uniform
// blockIdx.z stays in the uniform datapath.
__device__ __forceinline__ double pick4(const double (&a)[4], unsigned k)
{
    return k == 0 ? a[0] : k == 1 ? a[1] : k == 2 ? a[2] : a[3];
}

// CTA partial -> global; last CTA of the case reduces all partials of the case in fixed order.
__device__ __forceinline__ void block_reduce_and_publish(const double out[4], int case_idx, int tile, int tiles,
                                                         int out_slot, double *__restrict__ partials,
                                                         unsigned *__restrict__ tickets, double *__restrict__ sums)
{
    __shared__ double s_part[kThreads / 32][4];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double r[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) r[m] = warp_sum(out[m]);
    if (lane == 0) {
#pragma unroll
        for (int m = 0; m < 4; ++m) s_part[warp][m] = r[m];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) a += s_part[w][threadIdx.x];
        partials[((size_t)case_idx * tiles + tile) * 4 + threadIdx.x] = a;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&tickets[case_idx], 1u);
        s_last = (t == (unsigned)tiles - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // fixed-order: thread j takes tiles j, j+kThreads, ...; then a fixed shared-memory tree
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int tt = threadIdx.x; tt < tiles; tt += kThreads) {
        const double *p = partials + ((size_t)case_idx * tiles + tt) * 4;
#pragma unroll
        for (int m = 0; m < 4; ++m) a[m] += __ldcg(p + m);
    }
    __shared__ double s_red[kThreads][4];
#pragma unroll
    for (int m = 0; m < 4; ++m) s_red[threadIdx.x][m] = a[m];
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
#pragma unroll
            for (int m = 0; m < 4; ++m) s_red[threadIdx.x][m] += s_red[threadIdx.x + o][m];
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) sums[(size_t)out_slot * 4 + threadIdx.x] = s_red[0][threadIdx.x];
    if (threadIdx.x == 0) tickets[case_idx] = 0u;  // ready for the next launch
}

// -----------------------------------------------------------------------------------------
// Philox rollout.  Noise stream layout (restated draw for draw by oracle/kfbo_oracle.c):
//   key = (seed_lo, seed_hi); counter = (trial, case.stream, block, sub)
//   block 0, sub 0      -> the N(0,P0) draw added to x0 (radius x : w[31:12], angle y)
//   block p+1, sub 0,1  -> calls A, B of the step pair (2p, 2p+1) -> normals n0..n5 (normals_of_pair):
//                          step 2p uses n0,n1 (process) n2 (observation); step 2p+1 uses n3,n4,n5
// -----------------------------------------------------------------------------------------
// kPhiloxPerThread independent trajectories (adjacent tiles of the same case) can advance in
// lock-step in every thread (more ILP per warp).  Measured on B200 (profiles/NOTES_r1.md): the kernel
// is bound by the pipe that FP64 and IMAD.WIDE share, not by latency, so 1 and 2 trajectories per
// thread run at the same speed (19.96 vs 20.08 ms for cfg3) and 1 needs half the registers.
#ifndef KFBO_PHILOX_PER_THREAD
#define KFBO_PHILOX_PER_THREAD 1
#endif
#ifndef KFBO_PHILOX_MIN_BLOCKS
#define KFBO_PHILOX_MIN_BLOCKS 1
#endif
constexpr int kPhiloxPerThread = KFBO_PHILOX_PER_THREAD;

__global__ void __launch_bounds__(kThreads, KFBO_PHILOX_MIN_BLOCKS)
rollout_philox(const Case *__restrict__ cases, const __grid_constant__ SampleTimeConsts st,
               const double2 *__restrict__ gu_tables, const math::LogTables *__restrict__ logtab,
               const __grid_constant__ PhiloxKeys key, uint32_t trial0, uint32_t ntrials,
               int tiles, double *__restrict__ per_trial, size_t per_trial_ld, double *__restrict__ partials,
               unsigned *__restrict__ tickets, double *__restrict__ sums)
{
    constexpr int K = kPhiloxPerThread;
    extern __shared__ double2 s_gu[];
    __shared__ math::LogTables s_log;
    const int case_idx = blockIdx.y * gridDim.z + blockIdx.z, tile = blockIdx.x;
    const Case m = cases[case_idx];
    const int T = m.T;
    const unsigned kk = st.k0 + blockIdx.z;   // == m.k, but provably uniform
    const double f = pick4(st.f, kk), l00 = pick4(st.l00, kk), l10 = pick4(st.l10, kk), l11 = pick4(st.l11, kk),
                 sr = pick4(st.sr, kk);
    {
        const double2 *src = gu_tables + (size_t)kk * kMaxSteps;
        for (int i = threadIdx.x; i < T; i += kThreads) s_gu[i] = src[i];
        const double *ls = reinterpret_cast<const double *>(logtab);
        double *ld = reinterpret_cast<double *>(&s_log);
        for (int i = threadIdx.x; i < (int)(sizeof(math::LogTables) / sizeof(double)); i += kThreads) ld[i] = ls[i];
    }
    __syncthreads();
    const double (*lt)[2] = s_log.invc_logc;
    const double *kt = s_log.k2ln2;

    // trajectory j of this thread: local trial (tile*K + j)*kThreads + threadIdx.x (coalesced outputs)
    uint32_t local[K], trial[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        local[j] = ((uint32_t)tile * K + j) * kThreads + threadIdx.x;
        trial[j] = trial0 + local[j];   // past-the-end trajectories of the last tile are computed and discarded
    }
    Traj t[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double a, b;
        const uint4 w = philox4x32_10(trial[j], m.stream, 0u, 0u, key);
        math::normal_pair_bits(w.x, w.w >> 12, w.y, lt, kt, a, b);
        traj_init(t[j], a, b);
    }
    // The counter-based generator lets the random bits of step pair p+1 be produced while the FP64
    // work of pair p is in flight: integer (Philox) and FP64 instructions interleave inside a warp.
    const int npairs = (T + 1) >> 1;     // an odd T leaves the second half of the last pair unused
    uint4 r[K][2];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) r[j][c] = philox4x32_10(trial[j], m.stream, 1u, (uint32_t)c, key);
    for (int p = 0; p < npairs; ++p) {
        uint4 q[K][2];
        double n[K][6];
#pragma unroll
        for (int j = 0; j < K; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) q[j][c] = philox4x32_10(trial[j], m.stream, (uint32_t)p + 2u, (uint32_t)c, key);
#pragma unroll
        for (int j = 0; j < K; ++j) normals_of_pair(r[j][0], r[j][1], lt, kt, n[j]);
        const double2 g0 = s_gu[2 * p];
        if (p == 0) {
#pragma unroll
            for (int j = 0; j < K; ++j)
                filter_step<true>(t[j], m, f, g0, l00 * n[j][0], fma(l11, n[j][1], l10 * n[j][0]), sr * n[j][2]);
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j)
                filter_step<false>(t[j], m, f, g0, l00 * n[j][0], fma(l11, n[j][1], l10 * n[j][0]), sr * n[j][2]);
        }
        if (2 * p + 1 < T) {
            const double2 g1 = s_gu[2 * p + 1];
#pragma unroll
            for (int j = 0; j < K; ++j)
                filter_step<false>(t[j], m, f, g1, l00 * n[j][3], fma(l11, n[j][4], l10 * n[j][3]), sr * n[j][5]);
        }
#pragma unroll
        for (int j = 0; j < K; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) r[j][c] = q[j][c];
    }
    double out[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double o[4];
        traj_finish(t[j], T, o);
        if (local[j] < ntrials) {
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] += o[k];
            if (per_trial != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) per_trial[((size_t)m.out_slot * 4 + k) * per_trial_ld + local[j]] = o[k];
            }
        }
    }
    block_reduce_and_publish(out, case_idx, tile, tiles, m.out_slot, partials, tickets, sums);
}

// -----------------------------------------------------------------------------------------
// Supplied-noise rollout: x0noise[2][B], w[T][2][B], v[T][B]; trial index fastest, so a warp
// reads 32 consecutive doubles (256 B) per load.  Loads are software-pipelined kPrefetch
// steps ahead through registers (3*kPrefetch independent 8-byte loads in flight per thread).
// -----------------------------------------------------------------------------------------
constexpr int kPrefetch = 4;

__device__ __forceinline__ double ld_stream(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kThreads)
rollout_supplied(const Case *__restrict__ cases, const __grid_constant__ SampleTimeConsts st,
                 const double2 *__restrict__ gu_tables, const double *__restrict__ x0noise,
                 const double *__restrict__ w, const double *__restrict__ v, size_t B, int tiles,
                 double *__restrict__ per_trial, double *__restrict__ partials, unsigned *__restrict__ tickets,
                 double *__restrict__ sums)
{
    extern __shared__ double2 s_gu[];
    const int case_idx = blockIdx.y * gridDim.z + blockIdx.z, tile = blockIdx.x;
    const Case m = cases[case_idx];
    const int T = m.T;
    const unsigned kk = st.k0 + blockIdx.z;
    const double f = pick4(st.f, kk);
    {
        const double2 *src = gu_tables + (size_t)kk * kMaxSteps;
        for (int i = threadIdx.x; i < T; i += kThreads) s_gu[i] = src[i];
    }
    __syncthreads();

    const size_t i = (size_t)tile * kThreads + threadIdx.x;
    const bool active = i < B;
    double out[4] = {0.0, 0.0, 0.0, 0.0};
    if (active) {
        Traj t;
        traj_init(t, ld_stream(x0noise + i), ld_stream(x0noise + B + i));
        const double *pw = w + i, *pv = v + i;
        double cw0[kPrefetch], cw1[kPrefetch], cv[kPrefetch];
#pragma unroll
        for (int j = 0; j < kPrefetch; ++j) {
            const int s = j < T ? j : T - 1;
            cw0[j] = ld_stream(pw + (size_t)(2 * s) * B);
            cw1[j] = ld_stream(pw + (size_t)(2 * s + 1) * B);
            cv[j] = ld_stream(pv + (size_t)s * B);
        }
        for (int s0 = 0; s0 < T; s0 += kPrefetch) {
            double nw0[kPrefetch], nw1[kPrefetch], nv[kPrefetch];
#pragma unroll
            for (int j = 0; j < kPrefetch; ++j) {
                int s = s0 + kPrefetch + j;
                s = s < T ? s : T - 1;  // clamped re-read of the last step past the end (never consumed)
                nw0[j] = ld_stream(pw + (size_t)(2 * s) * B);
                nw1[j] = ld_stream(pw + (size_t)(2 * s + 1) * B);
                nv[j] = ld_stream(pv + (size_t)s * B);
            }
#pragma unroll
            for (int j = 0; j < kPrefetch; ++j) {
                const int s = s0 + j;
                if (s < T) {
                    if (s == 0) filter_step<true>(t, m, f, s_gu[s], cw0[j], cw1[j], cv[j]);
                    else        filter_step<false>(t, m, f, s_gu[s], cw0[j], cw1[j], cv[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < kPrefetch; ++j) { cw0[j] = nw0[j]; cw1[j] = nw1[j]; cv[j] = nv[j]; }
        }
        traj_finish(t, T, out);
        if (per_trial != nullptr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) per_trial[((size_t)m.out_slot * 4 + k) * B + i] = out[k];
        }
    }
    block_reduce_and_publish(out, case_idx, tile, tiles, m.out_slot, partials, tickets, sums);
}

// -----------------------------------------------------------------------------------------
// DFMA-chain microbenchmark: 8 independent FMA chains per thread.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_chain(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

// ------------------------------- launchers (host) -------------------------------------------
PhiloxKeys expand_philox_key(uint64_t seed)
{
    PhiloxKeys k;
    k.zero = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = k0; k.k1[r] = k1;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;   // Weyl sequence of Philox4x32
    }
    return k;
}

cudaError_t launch_rollout_philox(const Case *d_cases, int ncases, const SampleTimeConsts &st, const double2 *d_gu, const void *d_logtab, int max_T,
                                  uint64_t seed, uint32_t trial0, uint32_t ntrials, double *d_per_trial,
                                  size_t per_trial_ld, double *d_partials, unsigned *d_tickets, double *d_sums,
                                  cudaStream_t stream)
{
    if (ncases <= 0 || ntrials == 0) return cudaSuccess;
    const int tiles = philox_tiles(ntrials);
    const size_t smem = (size_t)max_T * sizeof(double2);
    const dim3 grid((unsigned)tiles, (unsigned)(ncases / st.nk), (unsigned)st.nk);   // ncases = candidates x nk
    rollout_philox<<<grid, kThreads, smem, stream>>>(
        d_cases, st, d_gu, static_cast<const math::LogTables *>(d_logtab), expand_philox_key(seed), trial0, ntrials, tiles,
        d_per_trial, per_trial_ld, d_partials, d_tickets, d_sums);
    return cudaGetLastError();
}

cudaError_t launch_rollout_supplied(const Case *d_cases, int ncases, const SampleTimeConsts &st, const double2 *d_gu, int max_T, const double *d_x0noise,
                                    const double *d_w, const double *d_v, size_t B, double *d_per_trial,
                                    double *d_partials, unsigned *d_tickets, double *d_sums, cudaStream_t stream)
{
    if (ncases <= 0 || B == 0) return cudaSuccess;
    const int tiles = (int)((B + kThreads - 1) / kThreads);
    const size_t smem = (size_t)max_T * sizeof(double2);
    const dim3 grid((unsigned)tiles, (unsigned)(ncases / st.nk), (unsigned)st.nk);
    rollout_supplied<<<grid, kThreads, smem, stream>>>(
        d_cases, st, d_gu, d_x0noise, d_w, d_v, B, tiles, d_per_trial, d_partials, d_tickets, d_sums);
    return cudaGetLastError();
}

cudaError_t launch_dfma_chain(double *d_out, int blocks, int iters, cudaStream_t stream)
{
    dfma_chain<<<blocks, 256, 0, stream>>>(d_out, iters, 0.999999, 1e-9);
    return cudaGetLastError();
}

int rollout_tiles(size_t ntrials) { return (int)((ntrials + kThreads - 1) / kThreads); }
int philox_tiles(size_t ntrials) { return (int)((ntrials + (size_t)kThreads * kPhiloxPerThread - 1) / ((size_t)kThreads * kPhiloxPerThread)); }

}  // namespace kfbo
