// kfbo_kernels.cu -- the rollout kernels (sm_100a) and their launchers.
//
//   rollout_philox          on-device Philox4x32-10 + Box-Muller noise, the reference form:
//                           truth state, estimate and covariance per thread (FP64-pipe bound)
//   rollout_supplied_bulk   noise streamed from HBM, SoA [step][component][trial], staged through
//                           shared memory by cp.async.bulk + mbarriers (HBM bound)
//   rollout_supplied        the same with per-thread loads (streams that are not 16-byte aligned)
//   trace_philox            one trajectory replayed step by step (the Nsimruns == 1 record)
//   peer_exchange_kernel    multi-GPU: the ranks' cost sums meet in each other's memory (stores + flags over NVLink),
//                           launched behind the rollout; the same code can run in the rollout's tail (kPeer, opt-in)
//
// Grid: (tiles, candidates, sample times): blockIdx.y/z give the case without any division, so the
// case and its sample-time constants are provably uniform (uniform registers / parameter bank).
// Reduction: warp shuffle -> shared -> one [4] partial per CTA -> the LAST CTA of each case
// (ticket counter) sums the partials in a fixed order, so sums are deterministic run to run.
#include <algorithm>
#include <atomic>

#include "kfbo_device.cuh"
#include "kfbo_host.h"
#include "kfbo_kernels.h"

namespace kfbo {

// Select entry k of a small per-sample-time table held in the parameter bank so that the result
// lands in a UNIFORM register pair: a register-indexed LDC, and also a chain of selects on the
// doubles themselves (predicated LDC.64), end up in vector registers, and a DFMA whose three sources
// are distinct vector registers costs 3 FP64-pipe cycles instead of 2.  Selecting the two 32-bit
// halves with integer selects on the uniform blockIdx.z keeps them in the uniform datapath
// (SASS: DFMA R, R, UR, R).
__device__ __forceinline__ double pick4(const double (&a)[4], unsigned k)
{
    const int lo = k == 0 ? __double2loint(a[0]) : k == 1 ? __double2loint(a[1]) : k == 2 ? __double2loint(a[2]) : __double2loint(a[3]);
    const int hi = k == 0 ? __double2hiint(a[0]) : k == 1 ? __double2hiint(a[1]) : k == 2 ? __double2hiint(a[2]) : __double2hiint(a[3]);
    return __hiloint2double(hi, lo);
}

// -----------------------------------------------------------------------------------------
// Cross-GPU exchange of the cost sums over peer memory (PeerCtx, kfbo_kernels.h).  One CTA runs it: the CTA that
// finished the launch's last case (peer_tail), or the single CTA of peer_exchange_kernel.
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// part / nparts: the CTAs of peer_exchange_kernel share the stores of step 1; the last of them to finish goes on alone.
__device__ __forceinline__ void peer_exchange(const PeerCtx *__restrict__ pc, int part, int nparts)
{
    __shared__ bool s_last_part;
    const int tid = threadIdx.x, nranks = pc->nranks, me = pc->rank;
    const unsigned epoch = pc->epoch, gen = epoch & 1u;
    const bool gather = pc->gather != 0;
    const size_t cap = pc->cap, slabs = gather ? 1u : (size_t)nranks;
    const size_t my_off = ((size_t)gen * slabs + (gather ? 0u : (size_t)me)) * cap;
    double *const sums = pc->sums;
    // 1. this rank's share into the exchange buffer of every GPU (its own included: the total is formed from the
    //    exchange buffers alone).  Slots are 4 doubles, so everything is 16-byte aligned; a value is read once and
    //    stored nranks times, kPeerUnroll independent reads in flight per thread.
    constexpr int kPeerUnroll = 4;
    const int b2 = pc->begin[me] >> 1, e2 = pc->end[me] >> 1;
    const double2 *src = reinterpret_cast<const double2 *>(sums);
    const int nthreads = (int)blockDim.x;
    for (int i0 = b2 + part * nthreads * kPeerUnroll + tid; i0 < e2; i0 += nparts * nthreads * kPeerUnroll) {
        double2 v[kPeerUnroll];
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u)
            if (i0 + u * nthreads < e2) v[u] = __ldcg(src + i0 + u * nthreads);
        for (int r = 0; r < nranks; ++r) {
            int to = me + r;                  // own buffer first, then the peers staggered by rank
            if (to >= nranks) to -= nranks;
            double2 *dst = reinterpret_cast<double2 *>(pc->vals[to] + my_off);
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u)
                if (i0 + u * nthreads < e2) dst[i0 + u * nthreads] = v[u];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (nparts > 1) {
        if (tid == 0) {
            const unsigned t = atomicAdd(pc->done, 1u);
            s_last_part = (t == (unsigned)nparts - 1u);
            if (s_last_part) *pc->done = 0u;
        }
        __syncthreads();
        if (!s_last_part) return;
        __threadfence();
    }
    // 2. one flag per GPU: "evaluation `epoch` of rank `me` has landed";  3. wait for the same from every rank
    if (tid < nranks) {
        st_release_sys(pc->flags[tid] + gen * kMaxPeers + me, epoch);
        const unsigned *f = pc->flags[me] + gen * kMaxPeers + tid;
        const unsigned long long t0 = global_ns(), limit = pc->timeout_ns;
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            if (global_ns() - t0 > limit) { *pc->status = 1 + tid; break; }   // reported by the host as KFBO_EPEER
            __nanosleep(20);
        }
    }
    __syncthreads();
    // 4. trials sharded: the slabs added in rank order (candidates sharded: the gathered block is the result as it is)
    if (!gather) {
        const double2 *slab0 = reinterpret_cast<const double2 *>(pc->vals[me] + (size_t)gen * slabs * cap);
        const size_t cap2 = cap >> 1;
        const int n2 = pc->nsums >> 1;
        for (int i = tid; i < n2; i += nthreads) {
            double2 v[kMaxPeers];
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r)
                if (r < nranks) v[r] = __ldcg(slab0 + (size_t)r * cap2 + i);
            double2 a = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r)
                if (r < nranks) { a.x += v[r].x; a.y += v[r].y; }
            reinterpret_cast<double2 *>(pc->out)[i] = a;
        }
    }
    // 5. tell the host (it may be polling mapped memory for exactly this instead of synchronising the stream)
    if (pc->host_flag != nullptr) {
        __threadfence_system();
        __syncthreads();
        if (tid == 0) st_release_sys(pc->host_flag, epoch);
    }
}

// Called by all threads of a CTA that has just published the sums of one case.
__device__ __forceinline__ void peer_tail(const PeerCtx *__restrict__ pc)
{
    __shared__ bool s_final;
    __syncthreads();                      // the four writers of this case's sums have fenced
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(pc->done, 1u);
        s_final = (t == pc->ncases - 1u);
        if (s_final) *pc->done = 0u;      // ready for the next launch
    }
    __syncthreads();
    if (!s_final) return;
    __threadfence();
    peer_exchange(pc, 0, 1);
}

// the PeerCtx of a launch sits in the kPeerCtxBytes in front of its case constants
__device__ __forceinline__ const PeerCtx *peer_ctx_of(const Case *cases)
{
    return reinterpret_cast<const PeerCtx *>(reinterpret_cast<const char *>(cases) - kPeerCtxBytes);
}

__global__ void __launch_bounds__(kThreads) peer_exchange_kernel(const PeerCtx *__restrict__ pc)
{
    peer_exchange(pc, (int)blockIdx.x, (int)gridDim.x);
}

// CTA partial -> global; last CTA of the case reduces all partials of the case in fixed order.
// s_red: kThreads x 4 doubles of shared memory the caller no longer needs (or a static array).
// peer (optional, uniform): the CTA that finishes the launch's last case goes on to the cross-GPU exchange.
template <int kT = kThreads>
__device__ __forceinline__ void block_reduce_and_publish(const double out[4], int case_idx, int tile, int tiles,
                                                         int out_slot, double *__restrict__ partials,
                                                         unsigned *__restrict__ tickets, double *__restrict__ sums,
                                                         double (*s_red)[4], const PeerCtx *__restrict__ peer = nullptr)
{
    __shared__ double s_part[kT / 32][4];
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
        for (int w = 0; w < kT / 32; ++w) a += s_part[w][threadIdx.x];
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
    // fixed-order: thread j takes tiles j, j+kT, ...; then a fixed shared-memory tree
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int tt = threadIdx.x; tt < tiles; tt += kT) {
        const double *p = partials + ((size_t)case_idx * tiles + tt) * 4;
#pragma unroll
        for (int m = 0; m < 4; ++m) a[m] += __ldcg(p + m);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) s_red[threadIdx.x][m] = a[m];
    __syncthreads();
    for (int o = kT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
#pragma unroll
            for (int m = 0; m < 4; ++m) s_red[threadIdx.x][m] += s_red[threadIdx.x + o][m];
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) sums[(size_t)out_slot * 4 + threadIdx.x] = s_red[0][threadIdx.x];
    if (threadIdx.x == 0) tickets[case_idx] = 0u;  // ready for the next launch
    if (peer != nullptr) {
        if (threadIdx.x < 4) __threadfence();
        peer_tail(peer);
    }
}

// -----------------------------------------------------------------------------------------
// Philox rollout.  Noise stream layout: kfbo_device.cuh (philox_group) -- three Philox4x32-10 calls per group of four
// steps, 64 random bits per Box-Muller pair.
// The counter-based generator lets the random bits of group g+1 be produced while the FP64 work of group g is in
// flight: integer (Philox) and FP64 instructions interleave inside a warp.
// Shared memory: the noise tables only (48.4 KB, dynamic); g*u[step] is the same address for the whole warp and comes
// through the read-only path.  4 CTAs per SM.
// -----------------------------------------------------------------------------------------
#ifndef KFBO_PHILOX_MIN_BLOCKS
#define KFBO_PHILOX_MIN_BLOCKS 4
#endif
#ifndef KFBO_PHILOX_GROUP_UNROLL
#define KFBO_PHILOX_GROUP_UNROLL 1
#endif
constexpr int KFBO_PHILOX_GROUP_UNROLL_V = KFBO_PHILOX_GROUP_UNROLL;

template <int kT = kThreads>
__device__ __forceinline__ void load_noise_tables(math::LogTables *dst, const math::LogTables *__restrict__ src)
{
    const double2 *ls = reinterpret_cast<const double2 *>(src);
    double2 *ld = reinterpret_cast<double2 *>(dst);
    static_assert(sizeof(math::LogTables) % sizeof(double2) == 0, "tables are copied 16 bytes at a time");
    for (int i = threadIdx.x; i < (int)(sizeof(math::LogTables) / sizeof(double2)); i += kT) ld[i] = ls[i];
}

template <bool kUniformEst, bool kShiftFirst, bool kPeer>
__global__ void __launch_bounds__(kThreads, KFBO_PHILOX_MIN_BLOCKS)
rollout_philox(const Case *__restrict__ cases, const __grid_constant__ SampleTimeConsts st,
               const double2 *__restrict__ gu_tables, const math::LogTables *__restrict__ logtab,
               const __grid_constant__ PhiloxKeys key, uint32_t trial0, uint32_t ntrials,
               int tiles, double *__restrict__ per_trial, size_t per_trial_ld, double *__restrict__ partials,
               unsigned *__restrict__ tickets, double *__restrict__ sums)
{
    extern __shared__ __align__(16) unsigned char s_tables_raw[];
    math::LogTables &s_log = *reinterpret_cast<math::LogTables *>(s_tables_raw);
    const int case_idx = blockIdx.y * gridDim.z + blockIdx.z, tile = blockIdx.x;
    Case m = cases[case_idx];
    const int T = m.T;
    const uint32_t zero = (uint32_t)T >> 30;  // 0 (T <= 2000), opaque to the compiler: KFBO_PHILOX_SPLIT_MUL experiment only
    const unsigned kk = st.k0 + blockIdx.z;   // == m.k, but provably uniform
    if (kUniformEst) {                        // one candidate per launch: its Qd / R come by value, like f and the truth factor
        m.q00 = pick4(st.q00, kk); m.q01 = pick4(st.q01, kk); m.q11 = pick4(st.q11, kk); m.r_est = pick4(st.r_est, kk);
    }
    const double f = pick4(st.f, kk), l00 = pick4(st.l00, kk), l10 = pick4(st.l10, kk), l11 = pick4(st.l11, kk),
                 sr = pick4(st.sr, kk);
    const double2 *__restrict__ gu = gu_tables + (size_t)kk * kMaxSteps;
    load_noise_tables(&s_log, logtab);
    __syncthreads();

    const uint32_t local = (uint32_t)tile * kThreads + threadIdx.x;
    const uint32_t trial = trial0 + local;   // past-the-end trajectories of the last tile are computed and discarded
    Traj t;
    {
        double a, b;
        const uint4 w = philox_at(trial, m.stream, m.stream_x, 0u, 0u, key);
        math::normal_pair(w.x, w.y, s_log, a, b);
        traj_init<kShiftFirst>(t, a, b);
    }
    // Whole groups of four steps: ONE basic block per group (no branch inside), so the scheduler is free to overlap the
    // table lookups and the log/sqrt chains of one step's normals with the filter recursion of the previous step.
    const int nfull = T >> 2;
    GroupWords r = philox_group(trial, m.stream, m.stream_x, 0u, key);
#pragma unroll KFBO_PHILOX_GROUP_UNROLL_V
    for (int g = 0; g < nfull; ++g) {
        const GroupWords q = philox_group(trial, m.stream, m.stream_x, (uint32_t)g + 1u, key, zero);
        const double2 *gs = gu + 4 * g;
        const double2 g0 = __ldg(gs), g1 = __ldg(gs + 1), g2 = __ldg(gs + 2), g3 = __ldg(gs + 3);
        double p0a, p0b, p1a, p1b, p2a, p2b, p3a, p3b, o0, o1, o2, o3;
        math::normal_pair(r.c.x, r.c.y, s_log, o0, o1);
        math::normal_pair(r.a.x, r.a.y, s_log, p0a, p0b);
        math::normal_pair(r.a.z, r.a.w, s_log, p1a, p1b);
        filter_step<false, true, kShiftFirst>(t, m, f, g0, StepNoise{p0a, p0b, o0}, l00, l10, l11, sr, nullptr, g == 0);
        math::normal_pair(r.c.z, r.c.w, s_log, o2, o3);
        math::normal_pair(r.b.x, r.b.y, s_log, p2a, p2b);
        filter_step<false, true, kShiftFirst>(t, m, f, g1, StepNoise{p1a, p1b, o1}, l00, l10, l11, sr);
        math::normal_pair(r.b.z, r.b.w, s_log, p3a, p3b);
        filter_step<false, true, kShiftFirst>(t, m, f, g2, StepNoise{p2a, p2b, o2}, l00, l10, l11, sr);
        filter_step<false, true, kShiftFirst>(t, m, f, g3, StepNoise{p3a, p3b, o3}, l00, l10, l11, sr);
        r = q;
    }
    // the partial last group (T mod 4 steps); its unused normals are never formed
    const int rem = T & 3, sT = 4 * nfull;
    if (rem > 0) {
        double pa, pb, o0, o1;
        math::normal_pair(r.c.x, r.c.y, s_log, o0, o1);
        math::normal_pair(r.a.x, r.a.y, s_log, pa, pb);
        filter_step<false, true, kShiftFirst>(t, m, f, __ldg(gu + sT), StepNoise{pa, pb, o0}, l00, l10, l11, sr, nullptr, nfull == 0);
        if (rem > 1) {
            math::normal_pair(r.a.z, r.a.w, s_log, pa, pb);
            filter_step<false, true, kShiftFirst>(t, m, f, __ldg(gu + sT + 1), StepNoise{pa, pb, o1}, l00, l10, l11, sr);
        }
        if (rem > 2) {
            math::normal_pair(r.c.z, r.c.w, s_log, o0, o1);
            math::normal_pair(r.b.x, r.b.y, s_log, pa, pb);
            filter_step<false, true, kShiftFirst>(t, m, f, __ldg(gu + sT + 2), StepNoise{pa, pb, o0}, l00, l10, l11, sr);
        }
    }
    double out[4] = {0.0, 0.0, 0.0, 0.0};
    {
        double o[4];
        traj_finish(t, T, o);
        if (local < ntrials) {
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = o[k];
            if (per_trial != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) per_trial[((size_t)m.out_slot * 4 + k) * per_trial_ld + local] = o[k];
            }
        }
    }
    // every thread is past its last table lookup: the table space doubles as the reduction scratch
    static_assert(sizeof(math::LogTables) >= sizeof(double) * 4 * kThreads, "tables too small for the reduction scratch");
    __syncthreads();
    // kPeer: a compile-time switch and not a run-time test, so that the single-GPU kernel is the same machine code with
    // or without the exchange (ptxas' schedule of the rollout loop moves with what follows it)
    block_reduce_and_publish(out, case_idx, tile, tiles, m.out_slot, partials, tickets, sums, reinterpret_cast<double (*)[4]>(s_tables_raw),
                             kPeer ? peer_ctx_of(cases) : nullptr);
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
                    if (s == 0) filter_step<true, false>(t, m, f, s_gu[s], StepNoise{cw0[j], cw1[j], cv[j]});
                    else        filter_step<false, false>(t, m, f, s_gu[s], StepNoise{cw0[j], cw1[j], cv[j]});
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
    __shared__ double s_red[kThreads][4];
    block_reduce_and_publish(out, case_idx, tile, tiles, m.out_slot, partials, tickets, sums, s_red);
}

// -----------------------------------------------------------------------------------------
// Supplied-noise rollout, bulk-copy variant (the default when the streams are 16-byte aligned and
// B is even).  The noise of a 128-trial tile is staged through shared memory by the bulk-copy
// engine (cp.async.bulk, completion on an mbarrier) kBulkSteps steps at a time in a ring of
// kBulkStages stages: no per-thread address arithmetic (the 64-bit IMAD.WIDE address products of
// the LDG variant sit on the pipe FP64 shares), no prefetch registers (106 -> ~70 registers, 6 CTAs
// per SM instead of 4), and the bytes in flight are held by the copy engine, not by the warps.
//   stage = { w[kBulkSteps][2][128], v[kBulkSteps][128], gu[kBulkSteps] }  (12 KB + 64 B at 4 steps)
//   full[s]  : 1 arrival (the expect_tx of the issuing lane) + the bytes of the stage
//   empty[s] : one arrival per warp once all its lanes have consumed the stage
// Warp 0 is also the producer: before it consumes chunk c it refills the stage chunk c-1 used.
// Measured on B200, 1M trials x 1000 steps (profiles/NOTES_r1.md): 4 steps x 2 stages (24 KB, 8 CTAs
// = 32 warps per SM) streams 5.85 TB/s = 89% of the measured HBM peak; the LDG variant 5.25 TB/s;
// deeper rings or larger chunks lower the CTAs per SM and lose more than they gain.
// -----------------------------------------------------------------------------------------
#ifndef KFBO_BULK_STEPS
#define KFBO_BULK_STEPS 4
#endif
#ifndef KFBO_BULK_STAGES
#define KFBO_BULK_STAGES 2
#endif
constexpr int kBulkSteps = KFBO_BULK_STEPS;
constexpr int kBulkStages = KFBO_BULK_STAGES;
static_assert(kMaxSteps % kBulkSteps == 0, "the g*u table is read in whole chunks");
static_assert(3 * kBulkSteps + 1 <= 32, "one lane of warp 0 per bulk copy of a stage");

struct __align__(128) BulkStage {
    double w[kBulkSteps][2][kThreads];
    double v[kBulkSteps][kThreads];
    double2 gu[kBulkSteps];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned); read-once data: L2 evict_first
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__global__ void __launch_bounds__(kThreads)
rollout_supplied_bulk(const Case *__restrict__ cases, const __grid_constant__ SampleTimeConsts st,
                      const double2 *__restrict__ gu_tables, const double *__restrict__ x0noise,
                      const double *__restrict__ w, const double *__restrict__ v, size_t B, int tiles,
                      double *__restrict__ per_trial, double *__restrict__ partials, unsigned *__restrict__ tickets,
                      double *__restrict__ sums)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    BulkStage *stage = reinterpret_cast<BulkStage *>(s_raw);
    __shared__ __align__(8) uint64_t s_full[kBulkStages], s_empty[kBulkStages];
    const int case_idx = blockIdx.y * gridDim.z + blockIdx.z, tile = blockIdx.x;
    const Case m = cases[case_idx];
    const int T = m.T;
    const unsigned kk = st.k0 + blockIdx.z;
    const double f = pick4(st.f, kk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t col0 = (size_t)tile * kThreads;
    const uint32_t ncols = (uint32_t)(B - col0 < (size_t)kThreads ? B - col0 : (size_t)kThreads);
    const uint32_t row_bytes = ncols * (uint32_t)sizeof(double);
    const int nchunks = (T + kBulkSteps - 1) / kBulkSteps;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kBulkStages; ++s) { mbar_init(&s_full[s], 1u); mbar_init(&s_empty[s], kThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint64_t policy = 0;
    const double2 *gu_src = gu_tables + (size_t)kk * kMaxSteps;
    // lane l of warp 0 issues copy l of a stage: l < 2*nv -> w row, l < 3*nv -> v row, l == 3*nv -> g*u chunk
    auto issue = [&](int chunk) {
        BulkStage &sg = stage[chunk % kBulkStages];
        uint64_t *bar = &s_full[chunk % kBulkStages];
        const int s0 = chunk * kBulkSteps;
        const int nv = T - s0 < kBulkSteps ? T - s0 : kBulkSteps;        // steps of this chunk that exist
        if (lane == 0) mbar_arrive_expect_tx(bar, 3u * (uint32_t)nv * row_bytes + (uint32_t)(kBulkSteps * sizeof(double2)));
        if (lane < 2 * nv)
            bulk_g2s(&sg.w[lane >> 1][lane & 1][0], w + ((size_t)(2 * s0 + lane)) * B + col0, row_bytes, bar, policy);
        else if (lane < 3 * nv)
            bulk_g2s(&sg.v[lane - 2 * nv][0], v + ((size_t)(s0 + lane - 2 * nv)) * B + col0, row_bytes, bar, policy);
        else if (lane == 3 * nv)
            bulk_g2s(&sg.gu[0], gu_src + s0, (uint32_t)(kBulkSteps * sizeof(double2)), bar, policy);
    };
    if (warp == 0) {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        for (int c = 0; c < kBulkStages - 1 && c < nchunks; ++c) issue(c);
    }

    const size_t i = col0 + threadIdx.x;
    const bool active = threadIdx.x < ncols;
    Traj t;
    traj_init(t, active ? ld_stream(x0noise + i) : 0.0, active ? ld_stream(x0noise + B + i) : 0.0);
    for (int c = 0; c < nchunks; ++c) {
        if (warp == 0) {
            const int nx = c + kBulkStages - 1;       // refill the stage chunk c-1 lived in
            if (nx < nchunks) {
                if (nx >= kBulkStages) mbar_wait(&s_empty[nx % kBulkStages], (uint32_t)((nx / kBulkStages - 1) & 1));
                issue(nx);
            }
        }
        const int sidx = c % kBulkStages;
        mbar_wait(&s_full[sidx], (uint32_t)((c / kBulkStages) & 1));
        const BulkStage &sg = stage[sidx];
        const int s0 = c * kBulkSteps;
#pragma unroll
        for (int j = 0; j < kBulkSteps; ++j) {
            if (s0 + j < T) {
                const double w0 = sg.w[j][0][threadIdx.x], w1 = sg.w[j][1][threadIdx.x], vv = sg.v[j][threadIdx.x];
                const double2 g = sg.gu[j];
                if (j == 0 && c == 0) filter_step<true, false>(t, m, f, g, StepNoise{w0, w1, vv});
                else                  filter_step<false, false>(t, m, f, g, StepNoise{w0, w1, vv});
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sidx]);
    }
    double out[4] = {0.0, 0.0, 0.0, 0.0};
    if (active) {
        traj_finish(t, T, out);
        if (per_trial != nullptr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) per_trial[((size_t)m.out_slot * 4 + k) * B + i] = out[k];
        }
    }
    // every chunk that was issued has been consumed by every thread: the ring is free to reuse
    static_assert(sizeof(BulkStage) * kBulkStages >= sizeof(double) * 4 * kThreads, "ring too small for the reduction scratch");
    __syncthreads();
    block_reduce_and_publish(out, case_idx, tile, tiles, m.out_slot, partials, tickets, sums, reinterpret_cast<double (*)[4]>(s_raw));
}

// -----------------------------------------------------------------------------------------
// The same rollout for SMALL launches -- the reference's own default workload is 200 trials x 201 steps x 2 sample times
// (config/vehicleParams.yaml:9-16): 4 CTAs of rollout_philox, one warp per scheduler, and the evaluation takes as long as
// ONE warp needs for T strictly sequential steps with the noise generator and the filter interleaved in its instruction
// stream (~410 cycles per step: 47 us of the 63 us call).  When the launch cannot fill the GPU anyway, latency is what
// counts, and the two halves of a step are independent until the noise is consumed: here a CTA is 32 trajectories,
// warp 0 runs the filter recursion (the consumer; its loop-carried chain is the covariance recursion, ~110 cycles per
// step) and warps 1-3 -- on the other three schedulers of the SM -- each make one of the three Philox calls of a step
// group and its two Box-Muller pairs (the producers), one group ahead, through a double buffer in shared memory.
// One __syncthreads per group of four steps orders the two sides; no spinning anywhere.
//   SAME noise streams, SAME filter_step, and the per-CTA partial sums are combined four at a time exactly as the four
//   warps of a rollout_philox CTA combine theirs: the cost sums are THE SAME BITS as rollout_philox's, so which kernel
//   runs is invisible (single evaluation == candidate of a batch, one GPU == several).
// tiles = ceil(ntrials / 32); partials hold [cases][tiles][4].
// -----------------------------------------------------------------------------------------
constexpr int kSmallTile = 32;

template <bool kUniformEst, bool kShiftFirst>
__global__ void __launch_bounds__(kThreads)
rollout_philox_small(const Case *__restrict__ cases, const __grid_constant__ SampleTimeConsts st,
                     const double2 *__restrict__ gu_tables, const math::LogTables *__restrict__ logtab,
                     const __grid_constant__ PhiloxKeys key, uint32_t trial0, uint32_t ntrials,
                     int tiles, double *__restrict__ per_trial, size_t per_trial_ld, double *__restrict__ partials,
                     unsigned *__restrict__ tickets, double *__restrict__ sums)
{
    extern __shared__ __align__(16) unsigned char s_tables_raw[];
    math::LogTables &s_log = *reinterpret_cast<math::LogTables *>(s_tables_raw);
    __shared__ double s_ring[2][12][kSmallTile];      // [buffer][normal of the group][lane]: A -> 0..3, B -> 4..7, C -> 8..11
    __shared__ bool s_last;
    const int case_idx = blockIdx.y * gridDim.z + blockIdx.z, tile = blockIdx.x;
    Case m = cases[case_idx];
    const int T = m.T;
    const unsigned kk = st.k0 + blockIdx.z;
    if (kUniformEst) {
        m.q00 = pick4(st.q00, kk); m.q01 = pick4(st.q01, kk); m.q11 = pick4(st.q11, kk); m.r_est = pick4(st.r_est, kk);
    }
    const double f = pick4(st.f, kk), l00 = pick4(st.l00, kk), l10 = pick4(st.l10, kk), l11 = pick4(st.l11, kk),
                 sr = pick4(st.sr, kk);
    const double2 *__restrict__ gu = gu_tables + (size_t)kk * kMaxSteps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t local = (uint32_t)tile * kSmallTile + lane;
    const uint32_t trial = trial0 + local;
    // the noise tables by ONE bulk copy (48.4 KB, 16-byte multiple) instead of 24 rounds of per-thread loads: the launch's fixed
    // cost is what a small evaluation mostly consists of
    __shared__ __align__(8) uint64_t s_tab_bar;
    if (threadIdx.x == 0) {
        mbar_init(&s_tab_bar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
        mbar_arrive_expect_tx(&s_tab_bar, (uint32_t)sizeof(math::LogTables));
        bulk_g2s(&s_log, logtab, (uint32_t)sizeof(math::LogTables), &s_tab_bar, policy);
    }
    __syncthreads();                         // the barrier's initialisation is visible to everyone who waits on it
    const int ngroups = (T + 3) >> 2, nfull = T >> 2;
    // the consumer's control inputs g*u of a group, fetched one group ahead (an L2 round trip must not sit in its step chain)
    double2 gn0 = make_double2(0.0, 0.0), gn1 = gn0, gn2 = gn0, gn3 = gn0;
    if (warp == 0) { gn0 = __ldg(gu); gn1 = __ldg(gu + 1); gn2 = __ldg(gu + 2); gn3 = __ldg(gu + 3); }
    mbar_wait(&s_tab_bar, 0u);
    // producer `warp` makes call `warp - 1` of group g (counter block g + 1) and its two normal pairs
    auto produce = [&](int g) {
        const uint4 w = philox_at(trial, m.stream, m.stream_x, (uint32_t)g + 1u, (uint32_t)warp - 1u, key);
        double a, b, c, d;
        math::normal_pair(w.x, w.y, s_log, a, b);
        math::normal_pair(w.z, w.w, s_log, c, d);
        double (*dst)[kSmallTile] = &s_ring[g & 1][4 * (warp - 1)];
        dst[0][lane] = a; dst[1][lane] = b; dst[2][lane] = c; dst[3][lane] = d;
    };
    Traj t;
    if (warp == 0) {
        double a, b;
        const uint4 w = philox_at(trial, m.stream, m.stream_x, 0u, 0u, key);
        math::normal_pair(w.x, w.y, s_log, a, b);
        traj_init<kShiftFirst>(t, a, b);
    } else {
        produce(0);
    }
    __syncthreads();
    for (int g = 0; g < ngroups; ++g) {
        if (warp == 0) {
            const double (*n)[kSmallTile] = s_ring[g & 1];
            // everything the group's four steps read, up front: the normals of the group (whole groups are always made)
            // and, for the group after, g*u
            const double2 g0 = gn0, g1 = gn1, g2 = gn2, g3 = gn3;
            const double a0 = n[0][lane], b0 = n[1][lane], a1 = n[2][lane], b1 = n[3][lane], a2 = n[4][lane], b2 = n[5][lane],
                         a3 = n[6][lane], b3 = n[7][lane], o0 = n[8][lane], o1 = n[9][lane], o2 = n[10][lane], o3 = n[11][lane];
            if (g + 1 < ngroups) {           // the table has kMaxSteps entries and 4 (g + 1) + 3 < 4 ngroups <= kMaxSteps + 3 ...
                const double2 *nx = gu + 4 * (g + 1);
                const int left = kMaxSteps - 4 * (g + 1);          // ... so only the very last group of T = kMaxSteps - 1..3 could run over
                gn0 = __ldg(nx); gn1 = __ldg(nx + (left > 1 ? 1 : 0)); gn2 = __ldg(nx + (left > 2 ? 2 : 0)); gn3 = __ldg(nx + (left > 3 ? 3 : 0));
            }
            if (g < nfull) {
                filter_step<false, true, kShiftFirst>(t, m, f, g0, StepNoise{a0, b0, o0}, l00, l10, l11, sr, nullptr, g == 0);
                filter_step<false, true, kShiftFirst>(t, m, f, g1, StepNoise{a1, b1, o1}, l00, l10, l11, sr);
                filter_step<false, true, kShiftFirst>(t, m, f, g2, StepNoise{a2, b2, o2}, l00, l10, l11, sr);
                filter_step<false, true, kShiftFirst>(t, m, f, g3, StepNoise{a3, b3, o3}, l00, l10, l11, sr);
            } else {                         // the partial last group: T mod 4 steps
                const int rem = T & 3;
                filter_step<false, true, kShiftFirst>(t, m, f, g0, StepNoise{a0, b0, o0}, l00, l10, l11, sr, nullptr, g == 0);
                if (rem > 1) filter_step<false, true, kShiftFirst>(t, m, f, g1, StepNoise{a1, b1, o1}, l00, l10, l11, sr);
                if (rem > 2) filter_step<false, true, kShiftFirst>(t, m, f, g2, StepNoise{a2, b2, o2}, l00, l10, l11, sr);
            }
        } else if (g + 1 < ngroups) {
            produce(g + 1);
        }
        __syncthreads();
    }
    // ---- results of the tile's trajectories (warp 0), then the reduction: warp sum -> partial of this 32-trajectory tile
    double r4[4] = {0.0, 0.0, 0.0, 0.0};
    if (warp == 0) {
        double o[4], out[4] = {0.0, 0.0, 0.0, 0.0};
        traj_finish(t, T, o);
        if (local < ntrials) {
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = o[k];
            if (per_trial != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) per_trial[((size_t)m.out_slot * 4 + k) * per_trial_ld + local] = o[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) r4[k] = warp_sum(out[k]);
        if (lane < 4) {
            partials[((size_t)case_idx * tiles + tile) * 4 + lane] = lane == 0 ? r4[0] : lane == 1 ? r4[1] : lane == 2 ? r4[2] : r4[3];
            __threadfence();
        }
        __syncwarp();
        if (lane == 0) {
            const unsigned tk = atomicAdd(&tickets[case_idx], 1u);
            s_last = (tk == (unsigned)tiles - 1u);
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // The last CTA of the case: the 32-trajectory partials four at a time, in order, from 0.0 -- what the four warps of a
    // rollout_philox CTA add up to for their 128-trajectory tile -- then that kernel's fixed-order sum over the tiles.
    double (*s_red)[4] = reinterpret_cast<double (*)[4]>(s_tables_raw);     // the tables are no longer needed
    const int tiles128 = (tiles + 3) >> 2;
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int tt = threadIdx.x; tt < tiles128; tt += kThreads) {
        double p[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < 4; ++w) {
            const int t32 = 4 * tt + w;
            if (t32 < tiles) {
                const double *src = partials + ((size_t)case_idx * tiles + t32) * 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) p[k] += __ldcg(src + k);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] += p[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s_red[threadIdx.x][k] = a[k];
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
#pragma unroll
            for (int k = 0; k < 4; ++k) s_red[threadIdx.x][k] += s_red[threadIdx.x + o][k];
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) sums[(size_t)m.out_slot * 4 + threadIdx.x] = s_red[0][threadIdx.x];
    if (threadIdx.x == 0) tickets[case_idx] = 0u;  // ready for the next launch
}

// -----------------------------------------------------------------------------------------
// Single-trial trace (the Nsimruns == 1 path of trial.cpp:69-78,116-129, which plots the estimation
// error inside its +-2 sigma band): one thread replays ONE trajectory of the Philox rollout -- same
// streams, same filter_step -- and records per step
//   {time, e1, e2, lb1, ub1, lb2, ub2, nees, nis},  lb/ub = -/+ 2 sqrt(P_ii).
// -----------------------------------------------------------------------------------------
__global__ void trace_philox(const Case *__restrict__ cases, const __grid_constant__ SampleTimeConsts st,
                             const double2 *__restrict__ gu_tables, const math::LogTables *__restrict__ logtab,
                             const __grid_constant__ PhiloxKeys key, uint32_t trial, double dt, double *__restrict__ trace)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Case m = cases[0];
    const unsigned kk = st.k0;
    const double f = st.f[kk], l00 = st.l00[kk], l10 = st.l10[kk], l11 = st.l11[kk], sr = st.sr[kk];
    const double2 *gu = gu_tables + (size_t)kk * kMaxSteps;
    const math::LogTables &tab = *logtab;
    Traj t;
    {
        double a, b;
        const uint4 w = philox_at(trial, m.stream, m.stream_x, 0u, 0u, key);
        math::normal_pair(w.x, w.y, tab, a, b);
        traj_init(t, a, b);
    }
    for (int s = 0; s < m.T; ++s) {
        // the normals of step s out of its group's words (kfbo_device.cuh: philox_group), formed anew every step
        const GroupWords w = philox_group(trial, m.stream, m.stream_x, (uint32_t)(s >> 2), key);
        const int q = s & 3;
        const uint4 pw = q < 2 ? w.a : w.b;
        double pa, pb, oa, ob;
        if (q & 1) math::normal_pair(pw.z, pw.w, tab, pa, pb); else math::normal_pair(pw.x, pw.y, tab, pa, pb);
        if (q < 2) math::normal_pair(w.c.x, w.c.y, tab, oa, ob); else math::normal_pair(w.c.z, w.c.w, tab, oa, ob);
        StepProbe probe;
        if (s == 0) filter_step<true, true>(t, m, f, gu[s], StepNoise{pa, pb, (q & 1) ? ob : oa}, l00, l10, l11, sr, &probe);
        else        filter_step<false, true>(t, m, f, gu[s], StepNoise{pa, pb, (q & 1) ? ob : oa}, l00, l10, l11, sr, &probe);
        double *row = trace + (size_t)s * 9;
        const double b1 = 2.0 * sqrt(t.p00), b2 = 2.0 * sqrt(t.p11);
        row[0] = s * dt;                       // trial.cpp:70
        row[1] = t.x0 - t.xh0; row[2] = t.x1 - t.xh1;
        row[3] = -b1; row[4] = b1; row[5] = -b2; row[6] = b2;
        row[7] = probe.nees; row[8] = probe.nis;
    }
}

// -----------------------------------------------------------------------------------------
// A candidate batch without the host on its critical path.
//   build_cases_kernel     16 bytes per candidate in -> the Case of every (candidate, sample time): continuousToDiscrete
//                          (continuous2discrete.hpp:8-32) and R = onoise/dt (system_models.hpp:143) with the HOST's
//                          arithmetic, operation for operation (kfbo_host.h: RoundedOps): bit-identical constants
//   assemble_costs_kernel  trial_node.cpp:70-119 per candidate: the four averages, the configured cost per sample time,
//                          max_element and its (first) index
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) build_cases_kernel(const double2 *__restrict__ cand, int ncand, const __grid_constant__ CaseBuildConsts cb,
                                                          Case *__restrict__ cases)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= ncand * cb.D) return;
    const int c = i / cb.D, k = i - c * cb.D;
    const double2 qr = cand[c];
    double Fd[4], Qd[4];
    host::discretize(qr.x, cb.dt[k], Fd, Qd);
    Case m;
    m.f = Fd[1];
    m.q00 = Qd[0]; m.q01 = Qd[1]; m.q10 = Qd[2]; m.q11 = Qd[3];
    m.r_est = __ddiv_rn(qr.y, cb.dt[k]);
    m.l00 = cb.l00[k]; m.l10 = cb.l10[k]; m.l11 = cb.l11[k]; m.sr = cb.sr[k];
    m.T = cb.T[k];
    const uint64_t stream = cb.stream0 + (uint64_t)(cb.c0 + c) * (uint64_t)cb.D + (uint64_t)k;
    m.stream = (uint32_t)stream;
    m.stream_x = (uint32_t)((stream >> 32) & 0x3FFFFFFFull) << 2;
    m.out_slot = (cb.c0 + c) * cb.D + k;
    cases[i] = m;
}

cudaError_t launch_build_cases(const double2 *d_cand, int ncand, const CaseBuildConsts &cb, Case *d_cases, cudaStream_t stream)
{
    if (ncand <= 0) return cudaSuccess;
    build_cases_kernel<<<(ncand * cb.D + 127) / 128, 128, 0, stream>>>(d_cand, ncand, cb, d_cases);
    return cudaGetLastError();
}

size_t compact_costs_bytes(int C, int D) { return (size_t)C * ((size_t)D + 1) * sizeof(double) + (size_t)C * sizeof(int32_t); }

__global__ void __launch_bounds__(128) assemble_costs_kernel(const double *__restrict__ sums, int C, const __grid_constant__ kfbo_params p,
                                                             double *__restrict__ cost, double *__restrict__ max_cost, int32_t *__restrict__ index_max)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= C) return;
    const int D = p.n_sample_times;
    double best = 0.0;
    int arg = 0;
    for (int k = 0; k < D; ++k) {
        const double4 s4 = *reinterpret_cast<const double4 *>(sums + ((size_t)c * D + k) * 4);
        const double s[4] = {s4.x, s4.y, s4.z, s4.w};
        double avg[4];
        host::averages(p, k, s, avg);
        const double j = host::cost_of(p, avg);
        cost[(size_t)c * D + k] = j;
        // std::max_element (trial_node.cpp:118-119): the FIRST maximal element; a later one replaces it only if the
        // current best compares less than it (NaN never does)
        if (k == 0 || best < j) { best = j; arg = k; }
    }
    max_cost[c] = best;
    index_max[c] = arg;
}

cudaError_t launch_assemble_costs(const double *d_sums, int C, const void *params, void *d_out, cudaStream_t stream)
{
    const kfbo_params &p = *static_cast<const kfbo_params *>(params);
    double *cost = static_cast<double *>(d_out), *mx = cost + (size_t)C * p.n_sample_times;
    int32_t *idx = reinterpret_cast<int32_t *>(mx + C);
    assemble_costs_kernel<<<(C + 127) / 128, 128, 0, stream>>>(d_sums, C, p, cost, mx, idx);
    return cudaGetLastError();
}

// -----------------------------------------------------------------------------------------
// DFMA-chain microbenchmark: 8 independent FMA chains per thread -- the FP64 roofline denominator.
//   kRegMultiplier = true  (the peak): x = fma(x, y, U), the multiplier y in a vector register of its own per chain and the
//       addend a uniform / parameter-bank operand: 2.01 cycles per warp instruction on an SMSP (tools/microbench/rf_ports.cu,
//       profiles/rf_ports_r1k.txt) = the pipe's 16 lanes x 2 passes;
//   kRegMultiplier = false: x = fma(x, U, U), both other operands uniform -- 2.22 cycles: the form round 1 reported its
//       fraction against (which flattered it by 10 %); kept as `peak_uniform_operands` for comparison.
// -----------------------------------------------------------------------------------------
template <bool kRegMultiplier>
__global__ void __launch_bounds__(256) dfma_chain(double *out, int iters, double a, double b)
{
    double x[8], y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x + j; y[j] = a - 1e-12 * (threadIdx.x + j); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = kRegMultiplier ? fma(x[j], y[j], b) : fma(x[j], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + y[j];
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

// ------------------------------- launchers (host) -------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel) instead of once per launch (1-2 us each).
// The attribute lives in the device's context; a race between two threads sets it twice to the same value.
static cudaError_t ensure_dynamic_smem(const void *kernel, int /*slot: unused, the table is keyed by the kernel itself*/, size_t smem)
{
    // a small open-addressed table keyed by (device, kernel): lock-free reads, entries only ever grow
    constexpr int kEntries = 256;
    struct Entry { std::atomic<const void *> key{nullptr}; std::atomic<int> dev{-1}; std::atomic<size_t> have{0}; };
    static Entry table[kEntries];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    size_t i = ((reinterpret_cast<uintptr_t>(kernel) >> 4) * 31u + (unsigned)dev) % kEntries;
    Entry *slot = nullptr;
    for (int probe = 0; probe < kEntries; ++probe, i = (i + 1) % kEntries) {
        const void *k = table[i].key.load(std::memory_order_acquire);
        if (k == kernel && table[i].dev.load(std::memory_order_acquire) == dev) { slot = &table[i]; break; }
        if (k == nullptr) {
            const void *expected = nullptr;
            if (table[i].key.compare_exchange_strong(expected, kernel, std::memory_order_acq_rel)) {
                table[i].dev.store(dev, std::memory_order_release);
                slot = &table[i];
                break;
            }
            if (expected == kernel && table[i].dev.load(std::memory_order_acquire) == dev) { slot = &table[i]; break; }
        }
    }
    if (slot && slot->have.load(std::memory_order_acquire) >= smem) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && slot) slot->have.store(smem, std::memory_order_release);
    return e;
}

PhiloxKeys expand_philox_key(uint64_t seed)
{
    PhiloxKeys k;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = k0; k.k1[r] = k1;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;   // Weyl sequence of Philox4x32
    }
    return k;
}

cudaError_t launch_rollout_philox(const Case *d_cases, int ncases, const SampleTimeConsts &st, const double2 *d_gu, const void *d_logtab, int min_T,
                                  uint64_t seed, uint32_t trial0, uint32_t ntrials, double *d_per_trial,
                                  size_t per_trial_ld, double *d_partials, unsigned *d_tickets, double *d_sums,
                                  cudaStream_t stream, const Case *h_cases, bool with_peer, PhiloxLaunchRecord *rec, int small_max_ctas)
{
    if (ncases <= 0 || ntrials == 0) return cudaSuccess;
    const int tiles = rollout_tiles(ntrials);
    const size_t smem = sizeof(math::LogTables);
    const dim3 grid((unsigned)tiles, (unsigned)(ncases / st.nk), (unsigned)st.nk);   // ncases = candidates x nk
    // the noise tables (48.4 KB) pass the default 48 KB of shared memory: opt in (per device, a few microseconds)
    const bool uniform_est = h_cases != nullptr && ncases == st.nk;
    // shift policy of the one-pass time variance (kfbo_device.cuh): the constant shift only when no trial is very short
    const bool shift_first = min_T < kMinStepsConstShift;
    auto *kernel = with_peer
                       ? (uniform_est ? (shift_first ? rollout_philox<true, true, true> : rollout_philox<true, false, true>)
                                      : (shift_first ? rollout_philox<false, true, true> : rollout_philox<false, false, true>))
                       : (uniform_est ? (shift_first ? rollout_philox<true, true, false> : rollout_philox<true, false, false>)
                                      : (shift_first ? rollout_philox<false, true, false> : rollout_philox<false, false, false>));
    const cudaError_t attr = ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), (with_peer ? 4 : 0) + (uniform_est ? 2 : 0) + (shift_first ? 1 : 0), smem);
    if (attr != cudaSuccess) return attr;
    SampleTimeConsts stc = st;
    if (uniform_est) set_uniform_estimator(&stc, h_cases);
    PhiloxLaunchRecord local;
    PhiloxLaunchRecord &r = rec ? *rec : local;
    r.func = reinterpret_cast<const void *>(kernel); r.grid = grid; r.block = kThreads; r.smem = smem;
    int tiles_arg = tiles;
    // a launch too small to fill the GPU runs as the latency form (rollout_philox_small: 32 trajectories per CTA, the noise
    // made by three producer warps) -- same bits, a third of the time
    const int tiles32 = (int)((ntrials + kSmallTile - 1) / kSmallTile);
    if (!with_peer && small_max_ctas > 0 && (long)tiles32 * ncases <= (long)small_max_ctas) {
        auto *small = uniform_est ? (shift_first ? rollout_philox_small<true, true> : rollout_philox_small<true, false>)
                                  : (shift_first ? rollout_philox_small<false, true> : rollout_philox_small<false, false>);
        const cudaError_t a2 = ensure_dynamic_smem(reinterpret_cast<const void *>(small), 16 + (uniform_est ? 2 : 0) + (shift_first ? 1 : 0), smem);
        if (a2 != cudaSuccess) return a2;
        r.func = reinterpret_cast<const void *>(small);
        r.grid = dim3((unsigned)tiles32, (unsigned)(ncases / st.nk), (unsigned)st.nk);
        tiles_arg = tiles32;
    }
    r.cases = d_cases; r.st = stc; r.gu = d_gu; r.logtab = d_logtab; r.key = expand_philox_key(seed); r.trial0 = trial0; r.ntrials = ntrials;
    r.tiles = tiles_arg; r.per_trial = d_per_trial; r.per_trial_ld = per_trial_ld; r.partials = d_partials; r.tickets = d_tickets; r.sums = d_sums;
    r.bind();
    return cudaLaunchKernel(r.func, r.grid, dim3(r.block), r.argv, r.smem, stream);
}

void set_uniform_estimator(SampleTimeConsts *st, const Case *h_cases)
{
    for (int i = 0; i < st->nk; ++i) {
        const int k = st->k0 + i;
        st->q00[k] = h_cases[i].q00; st->q01[k] = h_cases[i].q01; st->q11[k] = h_cases[i].q11; st->r_est[k] = h_cases[i].r_est;
    }
}

cudaError_t launch_peer_exchange(const PeerCtx *d_peer, int own_doubles, cudaStream_t stream)
{
    // one CTA per 8 KB of this rank's share (1024 doubles), at most 32
    const int nparts = std::max(1, std::min(32, (own_doubles + 1023) / 1024));
    peer_exchange_kernel<<<nparts, kThreads, 0, stream>>>(d_peer);
    return cudaGetLastError();
}

cudaError_t launch_trace_philox(const Case *d_case, const SampleTimeConsts &st, const double2 *d_gu, const void *d_logtab, uint64_t seed,
                                uint32_t trial, double dt, double *d_trace, cudaStream_t stream)
{
    trace_philox<<<1, 32, 0, stream>>>(d_case, st, d_gu, static_cast<const math::LogTables *>(d_logtab), expand_philox_key(seed), trial, dt, d_trace);
    return cudaGetLastError();
}

cudaError_t launch_rollout_supplied(const Case *d_cases, int ncases, const SampleTimeConsts &st, const double2 *d_gu, int max_T, const double *d_x0noise,
                                    const double *d_w, const double *d_v, size_t B, double *d_per_trial,
                                    double *d_partials, unsigned *d_tickets, double *d_sums, int variant, int smem_pad,
                                    cudaStream_t stream)
{
    if (ncases <= 0 || B == 0) return cudaSuccess;
    const int tiles = (int)((B + kThreads - 1) / kThreads);
    const dim3 grid((unsigned)tiles, (unsigned)(ncases / st.nk), (unsigned)st.nk);
    if (supplied_variant_for(variant, d_w, d_v, B) == kSuppliedBulk) {
        // per device (the attribute lives in the current context), a few microseconds per call
        const size_t smem = kBulkStages * sizeof(BulkStage) + (size_t)(smem_pad > 0 ? smem_pad : 0);   // pad: caps the CTAs per SM
        const cudaError_t attr = cudaFuncSetAttribute(rollout_supplied_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (attr != cudaSuccess) return attr;
        rollout_supplied_bulk<<<grid, kThreads, smem, stream>>>(
            d_cases, st, d_gu, d_x0noise, d_w, d_v, B, tiles, d_per_trial, d_partials, d_tickets, d_sums);
    } else {
        const size_t smem = (size_t)max_T * sizeof(double2);
        rollout_supplied<<<grid, kThreads, smem, stream>>>(
            d_cases, st, d_gu, d_x0noise, d_w, d_v, B, tiles, d_per_trial, d_partials, d_tickets, d_sums);
    }
    return cudaGetLastError();
}

int supplied_variant_for(int requested, const double *d_w, const double *d_v, size_t B)
{
    // bulk copies need 16-byte aligned rows: aligned bases and an even number of trials per row
    const bool can_bulk = (reinterpret_cast<uintptr_t>(d_w) % 16 == 0) && (reinterpret_cast<uintptr_t>(d_v) % 16 == 0) && (B % 2 == 0);
    if (requested == kSuppliedLdg || !can_bulk) return kSuppliedLdg;
    return kSuppliedBulk;
}

cudaError_t launch_dfma_chain(double *d_out, int blocks, int iters, bool reg_multiplier, cudaStream_t stream)
{
    if (reg_multiplier) dfma_chain<true><<<blocks, 256, 0, stream>>>(d_out, iters, 0.999999, 1e-9);
    else                dfma_chain<false><<<blocks, 256, 0, stream>>>(d_out, iters, 0.999999, 1e-9);
    return cudaGetLastError();
}

int rollout_tiles(size_t ntrials) { return (int)((ntrials + kThreads - 1) / kThreads); }
int small_tiles(size_t ntrials) { return (int)((ntrials + kSmallTile - 1) / kSmallTile); }

}  // namespace kfbo

#include "kfbo_aug.cuh"
