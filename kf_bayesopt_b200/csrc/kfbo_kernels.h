// kfbo_kernels.h -- host-visible launch interface of kfbo_kernels.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace kfbo {

#ifndef KFBO_THREADS
#define KFBO_THREADS 128
#endif
constexpr int kThreads = KFBO_THREADS;   // threads per CTA (128 = 4 warps, one per SMSP)
constexpr int kMaxSteps = 2000;      // include/system_models.hpp:51

// Constants of one (candidate, sample time) pair, built on the host (kfbo_api.cu: build_case).
// The robot1d model has Fd = [[1, f],[0, 1]] and H = [1 0] exactly
// (include/system_models.hpp:35-37,141-142 through continuous2discrete.hpp:8-32), so the
// products with the structural 0/1 entries are dropped: x*1 and x+0*y are exact in IEEE-754,
// the results are bit-identical to the dense 2x2 products for finite values.
struct Case {
    double f;                    // Fd(0,1) = dt
    double q00, q01, q10, q11;   // estimator Qd (kalman_filter.cpp:23)
    double r_est;                // estimator R = onoise_est/dt (system_models.hpp:143)
    double l00, l10, l11;        // Cholesky factor of the truth Qd  (replaces eigenmvn.h:126-130)
    double sr;                   // sqrt(truth R)
    int32_t T;                   // max_iterations_
    uint32_t stream_x;           // bits 32..61 of the Philox stream id, shifted left by 2: XORed into the counter word that
                                 // holds the call index `sub` (0..2) -- kfbo_device.cuh: philox_at
    uint32_t stream;             // bits 0..31 of the Philox stream id of this (evaluation, candidate, sample time)
    int32_t out_slot;            // row of the [n][4] sums / per-trial output this case writes
};

// The constants of a Case that depend on the sample time only, handed to the kernels BY VALUE so
// they sit in the parameter bank / uniform registers: f and l11 are multiplicands of DFMAs whose
// other two sources are vector registers, and a DFMA reading three distinct vector registers costs
// 3 FP64-pipe cycles instead of 2 (tools/microbench/rf_ports.cu).
struct SampleTimeConsts {
    double f[4], l00[4], l10[4], l11[4], sr[4];   // indexed by the sample-time index k < KFBO_MAX_SAMPLE_TIMES
    int32_t k0, nk;                               // a launch covers sample times [k0, k0+nk): grid.z = nk, cases laid out [candidate][k]
    // A launch of ONE candidate (the "noise" -> "cost" evaluation) also carries that candidate's estimator constants
    // per sample time, so that they too are uniform-register / parameter-bank operands instead of four doubles loaded
    // into every thread's vector registers (the kernel sits at its 128-register cap).
    double q00[4], q01[4], q11[4], r_est[4];
};

// The 10 Philox4x32 round keys (key + r*Weyl), expanded on the host so the kernel reads them from
// the parameter bank instead of re-deriving them per call.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};
PhiloxKeys expand_philox_key(uint64_t seed);

// Cross-GPU exchange of the cost sums through peer memory (NVLink / NVSwitch): this GPU's share of the [C][D][4] sums
// is stored into an exchange buffer on every GPU of the job, one flag per peer is raised, the peers' flags are awaited
// and the total formed -- no NCCL launch and no host round trip between the rollout and the all-reduce.  It runs as
// peer_exchange_kernel behind the rollout, or (opt-in, slower: see KFBO_OPT_PEER_FUSE in kfbo.h) in the tail of the
// rollout kernel, by the CTA that finishes the launch's last case.  One PeerCtx per evaluation sits in front of the
// case constants in device memory (one host-to-device copy brings both).
//   trials sharded     every rank owns every slot; rank r's values land in slab r of each peer, and every rank adds the
//                      slabs in rank order: the total is the same bit pattern on all ranks and from run to run.
//   candidates sharded every slot has one owner, who stores it straight into its place of each peer's gathered block;
//                      nothing is added.
// Two generations (parity of the evaluation number) of slabs and flags: a rank that is one evaluation ahead of a peer
// writes the other generation, and it cannot get two ahead (it needs that peer's flag of the evaluation between).
constexpr int kMaxPeers = 8;              // the GPUs of one NVSwitch box
struct PeerCtx {
    double *vals[kMaxPeers];              // rank r's exchange buffer as mapped HERE: [2 generations][slabs][cap] doubles
    unsigned *flags[kMaxPeers];           // rank r's flags as mapped here: [2 generations][kMaxPeers], value = evaluation number
    int32_t begin[kMaxPeers], end[kMaxPeers];   // the doubles of the sums that rank r owns
    int32_t nranks, rank;
    int32_t gather;                       // 1: candidates sharded (one owner per slot), 0: trials sharded (slabs, summed)
    int32_t nsums;                        // C * D * 4
    uint32_t cap;                         // doubles per slab
    uint32_t epoch;                       // evaluation number, from 1
    uint32_t ncases;                      // cases of the launch that carries the exchange (its last case's CTA runs it)
    uint32_t pad0;
    unsigned *done;                       // local: cases finished in this launch
    unsigned *host_flag;                  // local, mapped host memory (or NULL): set to `epoch` once the total has been stored, so
                                          // that the host can poll for the result instead of synchronising the stream
    int32_t *status;                      // local: set to 1 when a peer's flag did not arrive within timeout_ns
    double *sums;                         // local: this rank's sums (the rollout's output)
    double *out;                          // trials sharded: where the total goes (device memory or mapped host memory)
    unsigned long long timeout_ns;
};
constexpr size_t kPeerCtxBytes = 512;     // the header in front of the case constants
static_assert(sizeof(PeerCtx) <= kPeerCtxBytes, "PeerCtx must fit the header in front of the case constants");
// the same exchange as a kernel of its own, launched behind the rollout (own_doubles: the size of this rank's share; the
// stores are spread over a few CTAs, the last of which raises the flags and waits)
cudaError_t launch_peer_exchange(const PeerCtx *d_peer, int own_doubles, cudaStream_t stream);

// d_cases[ncases]; d_gu[D][kMaxSteps] = (g0*u, g1*u); d_logtab = math::LogTables on the device;
// trials [trial0, trial0+ntrials).  d_per_trial (optional) is [slots][4][per_trial_ld];
// d_partials [ncases][tiles][4]; d_tickets [ncases] zero-initialised once; d_sums [slots][4].
// with_peer: the launch ends with the cross-GPU exchange of the sums; its PeerCtx sits in the kPeerCtxBytes in front of
// d_cases (no kernel parameter of its own: one more parameter changes ptxas' schedule of the rollout loop).
// h_cases (optional): the host copy of d_cases; given, a single-candidate launch passes the estimator constants by value.
// min_T: the smallest step count among the cases of the launch (selects the shift policy of the time variance, kfbo_device.cuh).
// rec (optional): a copy of the launch -- function, grid, argument values -- for a caller that replays it from a CUDA graph
// and only swaps argument values between replays (kfbo_api.cu: the single-candidate evaluation).
struct PhiloxLaunchRecord {
    const void *func = nullptr;
    dim3 grid;
    unsigned block = 0;
    size_t smem = 0;
    // the argument list of rollout_philox, in order
    const Case *cases; SampleTimeConsts st; const double2 *gu; const void *logtab; PhiloxKeys key; uint32_t trial0, ntrials; int tiles;
    double *per_trial; size_t per_trial_ld; double *partials; unsigned *tickets; double *sums;
    void *argv[13];
    void bind()
    {
        void *a[13] = {&cases, &st, &gu, &logtab, &key, &trial0, &ntrials, &tiles, &per_trial, &per_trial_ld, &partials, &tickets, &sums};
        for (int i = 0; i < 13; ++i) argv[i] = a[i];
    }
};
// The by-value estimator constants of a single-candidate launch (SampleTimeConsts::q00 ...), refreshed from its cases.
void set_uniform_estimator(SampleTimeConsts *st, const Case *h_cases);
cudaError_t launch_rollout_philox(const Case *d_cases, int ncases, const SampleTimeConsts &st, const double2 *d_gu, const void *d_logtab, int min_T,
                                  uint64_t seed, uint32_t trial0, uint32_t ntrials, double *d_per_trial,
                                  size_t per_trial_ld, double *d_partials, unsigned *d_tickets, double *d_sums,
                                  cudaStream_t stream, const Case *h_cases = nullptr, bool with_peer = false, PhiloxLaunchRecord *rec = nullptr,
                                  int small_max_ctas = 0);
// small_max_ctas > 0: a launch of at most that many 32-trajectory CTAs runs as the latency form of the kernel
// (rollout_philox_small; same bits); d_partials must then hold [ncases][small_tiles(ntrials)][4].
int small_tiles(size_t ntrials);

cudaError_t launch_rollout_supplied(const Case *d_cases, int ncases, const SampleTimeConsts &st, const double2 *d_gu, int max_T, const double *d_x0noise,
                                    const double *d_w, const double *d_v, size_t B, double *d_per_trial,
                                    double *d_partials, unsigned *d_tickets, double *d_sums, int variant, int smem_pad,
                                    cudaStream_t stream);

// Which supplied-noise kernel runs: kSuppliedAuto picks the bulk-copy (cp.async.bulk + mbarrier)
// variant whenever the streams allow it (16-byte aligned, even B), else the LDG variant.
enum { kSuppliedAuto = 0, kSuppliedLdg = 1, kSuppliedBulk = 2 };
int supplied_variant_for(int requested, const double *d_w, const double *d_v, size_t B);

// One trajectory replayed step by step: d_trace[T][9] = {time, e1, e2, lb1, ub1, lb2, ub2, nees, nis}; st.k0 = sample time.
cudaError_t launch_trace_philox(const Case *d_case, const SampleTimeConsts &st, const double2 *d_gu, const void *d_logtab, uint64_t seed,
                                uint32_t trial, double dt, double *d_trace, cudaStream_t stream);

// reg_multiplier: x = fma(x, y, U) (the full-rate operand form, the peak) or x = fma(x, U, U) (10 % slower; see kfbo_kernels.cu)
// ---- a candidate batch without the host on the critical path (kfbo_eval_batch with more than one candidate, kfbo_eval_batch_costs)
// build_cases_kernel: the per-(candidate, sample time) constants from 16 bytes per candidate, on the device, with the
// host's arithmetic operation for operation (kfbo_host.h: RoundedOps) -- the Case table is bit for bit what
// build_case (kfbo_api.cu) makes on the host.
struct CaseBuildConsts {
    double dt[4], l00[4], l10[4], l11[4], sr[4];   // per sample time: dt and the truth noise transform
    int32_t T[4];
    int32_t D;                                     // sample times per candidate
    int32_t c0;                                    // index of the first candidate of d_cand in the whole batch (out_slot, stream)
    uint64_t stream0;                              // stream id of candidate 0, sample time 0 of this evaluation
};
cudaError_t launch_build_cases(const double2 *d_cand /* (pnoise, onoise) */, int ncand, const CaseBuildConsts &cb, Case *d_cases, cudaStream_t stream);

// assemble_costs_kernel: trial_node.cpp:70-119 per candidate on the device -- averages, the configured cost per sample
// time, the first-max selection -- from the [C][D][4] sums; out = cost[C][D] doubles, then max_cost[C] doubles, then
// index_max[C] int32 (one contiguous block, copied to the host in one piece).
size_t compact_costs_bytes(int C, int D);
cudaError_t launch_assemble_costs(const double *d_sums, int C, const void *params /* const kfbo_params* (host) */, void *d_out, cudaStream_t stream);

// ---- the state-augmented variants (N = 3, 4 states; kfbo_aug.cuh): one covariance row per lane of a 4-lane group ----
constexpr int kAugGroup = 4;          // lanes per trajectory = the largest supported state dimension
struct AugCase {                      // per (candidate, sample time); matrices row-major 4 x 4, zero-padded beyond N
    double qd[16];                    // estimator Qd
    double r_est;                     // estimator R = onoise_est/dt
    int32_t T;
    uint32_t stream_x, stream;        // Philox stream id, as in Case
    int32_t out_slot;
};
struct AugSampleTime {                // per sample time: depends on dt and the truth side only
    double F[16];                     // Fd
    double g[4];                      // g_i = dt^(N-i)/(N-i)!
    double L[16];                     // lower Cholesky factor of the truth Qd
    double sr;                        // sqrt(truth R)
};
int aug_tiles(size_t ntrials);        // CTAs per case: 32 trajectories per 128-thread CTA
// supplied: x0noise[N][B], w[T][N][B], v[T][B] (device), trials [0, ntrials) of them; else on-device Philox noise.
// d_u[D][kMaxSteps]: the control table per sample time; cases laid out [candidate][sample time k0 .. k0+nk).
cudaError_t launch_rollout_aug(int N, bool supplied, const AugCase *d_cases, int ncases, int nk, int k0, const AugSampleTime *d_sts,
                               const double *d_u, const void *d_logtab, uint64_t seed, uint32_t trial0, uint32_t ntrials,
                               const double *d_x0noise, const double *d_w, const double *d_v, size_t B, double *d_per_trial,
                               size_t per_trial_ld, double *d_partials, unsigned *d_tickets, double *d_sums, cudaStream_t stream,
                               const AugSampleTime *h_sts = nullptr /* host copy of d_sts: the per-thread layout takes a sample time's constants by value */,
                               int layout = 0);
constexpr int kAugLayoutLanes = 0, kAugLayoutThread = 1;   // KFBO_AUG_LAYOUT_* of kfbo.h

cudaError_t launch_dfma_chain(double *d_out, int blocks, int iters, bool reg_multiplier, cudaStream_t stream);

int rollout_tiles(size_t ntrials);   // CTAs per case of the rollout kernels (128-thread CTAs)

// The Philox stream id is 62 bits wide (kfbo.h: KFBO_STREAM_BITS): where its halves go in a Case.
inline void set_case_stream(Case *c, uint64_t stream)
{
    c->stream = (uint32_t)stream;
    c->stream_x = (uint32_t)((stream >> 32) & 0x3FFFFFFFull) << 2;
}

}  // namespace kfbo
