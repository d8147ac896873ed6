/*
 * kfbo.h -- C ABI of the B200-native Monte Carlo cost evaluator for kf_bayesopt.
 *
 * Drop-in boundary for ONE hot path of arpg/kf_bayesopt: the Monte Carlo cost
 * evaluation (truth + measurement rollout, Kalman predict/update, per-trial
 * NEES/NIS, J_NEES/J_NIS over the sample times).  Every entry point cites the
 * reference interface (file:line, relative to the reference tree) it replaces.
 *
 * Plain C: pointers and sizes only, no C++ / torch types.  All functions return
 * 0 on success, <0 for an invalid argument, >0 for a CUDA / NCCL failure, and never
 * throw.  One evaluation in flight per handle (like the single ros::spin() thread
 * that drives RunTrial::ProcessNoiseCallback, trial_node.cpp:154).
 *
 * There is NO CPU fallback: without a CUDA device kfbo_create fails.
 *
 * The CUDA device that is current on the calling thread is PRESERVED by every entry point (a handle may live on another
 * device than the caller's other CUDA work, e.g. torch's).
 */
#ifndef KFBO_H_
#define KFBO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KFBO_MAX_SAMPLE_TIMES 4   /* reference: it_num = 2 (trial_node.cpp:49) */
#define KFBO_MAX_STEPS 2000       /* control table size (include/system_models.hpp:51) */
#define KFBO_STATE_DIM 2          /* #define N 2  (include/system_models.hpp:11) */
#define KFBO_OBS_DIM 1            /* #define Me 1 (include/system_models.hpp:12) */

/* rv.cost_choice_ (read_para_vehicle.cpp:157; used at trial_node.cpp:70,86) */
enum { KFBO_JNEES = 0, KFBO_CNEES = 1, KFBO_JNIS = 2, KFBO_CNIS = 3 };

/* error codes (negative = caller error) */
enum {
    KFBO_OK = 0,
    KFBO_EINVAL = -1,      /* bad argument (NULL pointer, size mismatch, ...) */
    KFBO_ERANGE = -2,      /* max_iterations outside [2, 2000], too many sample times ... */
    KFBO_ENODEVICE = -3,   /* no usable CUDA device: there is no CPU fallback */
    KFBO_ESTATE = -4,      /* call not valid in this state (comm already initialised; a multi-GPU handle / group after a
                              failed evaluation: its ranks are out of step, destroy it and start over) */
    KFBO_ECUDA = 1,        /* a CUDA runtime call failed; see kfbo_last_error */
    KFBO_ENCCL = 2,        /* an NCCL call failed; see kfbo_last_error */
    KFBO_EPEER = 3         /* the peer-memory exchange of the cost sums timed out waiting for a rank */
};

/* how a multi-rank evaluation is partitioned (kfbo_comm_init) */
enum { KFBO_SHARD_TRIALS = 0, KFBO_SHARD_CANDIDATES = 1 };

/*
 * The fields of ReadParaVehicle (include/read_para_vehicle.h:9-49) that the hot path
 * reads, for the TRUTH side (rv_gt in trial_node.cpp:26); the estimator's pnoise/onoise
 * arrive per evaluation like the "noise" message (trial_node.cpp:29-37).
 * The two-sample-time loop (trial_node.cpp:52,105-114) is data here: pass k uses
 * dt[k], max_iterations[k].  Reference defaults: {0.1,201},{1.0,201}.
 */
typedef struct kfbo_params {
    int32_t n_sample_times;                          /* it_num (trial_node.cpp:49) */
    double dt[KFBO_MAX_SAMPLE_TIMES];                /* dt_ per pass */
    int32_t max_iterations[KFBO_MAX_SAMPLE_TIMES];   /* max_iterations_ per pass (read_para_vehicle.cpp:145) */
    int32_t nsimruns;                                /* Nsimruns: Monte Carlo trials per sample time */
    int32_t cpu_core_number;                         /* CPUcoreNumber: kept ONLY for the reference's integer
                                                        semantics (trial.cpp:19,28); 1 = no truncation */
    int32_t state_dof;                               /* stateDOF (trial_node.cpp:77) */
    int32_t observation_dof;                         /* observationDOF (trial_node.cpp:93) */
    double pnoise_truth;                             /* rv_gt.pnoise_[0] (system_models.hpp:39) */
    double onoise_truth;                             /* rv_gt.onoise_[0] (system_models.hpp:143) */
    int32_t cost_choice;                             /* KFBO_JNEES ... */
    int32_t device;                                  /* CUDA device ordinal, -1 = current device */
    uint64_t seed;                                   /* Philox key; replaces the wall-clock seeds
                                                        (system_models.hpp:74,147, simulator.cpp:9) */
    int32_t max_candidates;                          /* capacity of kfbo_eval_batch (>= 1) */
    int32_t state_dim;                               /* 0 or 2: the reference's model (#define N 2, system_models.hpp:11).
                                                        3 or 4: the STATE-AUGMENTED variant north_star names -- the same
                                                        recursion on a chain of 3 / 4 integrators (position, velocity,
                                                        acceleration [, jerk]; noise and control drive the highest
                                                        derivative; H = [1 0 ..]).  The reference has no such variant:
                                                        oracle/kfbo_oracle_aug.c states it.  Set state_dof to match.
                                                        Noise arrays of the parity entry points then have state_dim rows
                                                        where the text below says 2. */
} kfbo_params;

/* What RunTrial::ProcessNoiseCallback computes per "noise" message (trial_node.cpp:70-119). */
typedef struct kfbo_result {
    double average_nees[KFBO_MAX_SAMPLE_TIMES];      /* mean of Trial::get_average_nees (trial_node.cpp:74) */
    double average_nis[KFBO_MAX_SAMPLE_TIMES];       /* (:90) */
    double average_var_nees[KFBO_MAX_SAMPLE_TIMES];  /* (:75) */
    double average_var_nis[KFBO_MAX_SAMPLE_TIMES];   /* (:91) */
    double cost[KFBO_MAX_SAMPLE_TIMES];              /* pic_max[k] for the configured cost_choice */
    double max_cost;                                 /* *max_element(pic_max) (:118) -- the "cost" message */
    int32_t index_max;                               /* first maximal index (:119) */
    int32_t reserved;
} kfbo_result;

typedef struct kfbo_handle kfbo_handle;

/* Fill *p with the reference defaults (config/vehicleParams.yaml + trial_node.cpp:105-114). */
void kfbo_default_params(kfbo_params *p);

/* Replaces the per-callback construction of vector<Trial> and the models inside
 * (trial_node.cpp:55-58 -> trial.cpp:12 -> system_models.hpp:30-78,137-150).  Allocates all
 * device buffers, tables and streams once; kfbo_eval* never allocate. */
int kfbo_create(const kfbo_params *p, kfbo_handle **out);
void kfbo_destroy(kfbo_handle *h);

/* Text of the last failure on this handle (h == NULL: last kfbo_create failure of this thread). */
const char *kfbo_last_error(const kfbo_handle *h);

/* One "noise" message -> one "cost": replaces the body of RunTrial::ProcessNoiseCallback
 * (trial_node.cpp:29-119).  pn[npn], on[non] are the message halves (only [0] is read, like
 * system_models.hpp:39,143).  On-device Philox noise; each call draws fresh streams.
 * Domain of a candidate (pn[0], on[0]): finite and > 0.  Outside it (NaN, inf, <= 0) the call still SUCCEEDS and the
 * result is all NaN (max_cost NaN, index_max 0) -- what the reference observably does: its filter degenerates and
 * std::abs(std::log(.)) (trial_node.cpp:77,93) publishes a NaN "cost", which ends the optimiser's wait loop
 * (robot1d_bayesopt.cpp:107).  Positive values outside [1e-100, 1e100] return KFBO_ERANGE. */
int kfbo_eval(kfbo_handle *h, const double *pn, int32_t npn, const double *on, int32_t non,
              kfbo_result *out);

/* n_candidates evaluations in one launch (a batched ProcessNoiseCallback): q_est[c], r_est[c]
 * are pnoise_[0], onoise_[0] of candidate c; out[n_candidates]. */
int kfbo_eval_batch(kfbo_handle *h, int32_t n_candidates, const double *q_est, const double *r_est,
                    kfbo_result *out);

/* The same batch with the COST ASSEMBLY ON THE DEVICE (trial_node.cpp:70-119 per candidate: averages, the configured
 * cost per sample time, max_element and its first index) and a compact answer: cost[n_candidates][D], max_cost and
 * index_max[n_candidates] (host arrays).  What a cost-surface sweep wants: per candidate 16 bytes go to the device and
 * 8 (D + 1) + 4 come back, and the host does no per-candidate arithmetic.  Same divisions as kfbo_finalize in the same
 * order; log() is CUDA's (within 1 ulp of libm's), so costs agree with kfbo_eval_batch's to ~1e-15 relative and the
 * selection is the same unless two sample times' costs tie to the last bit.  kfbo_last_sums is not available after it. */
int kfbo_eval_batch_costs(kfbo_handle *h, int32_t n_candidates, const double *q_est, const double *r_est,
                          double *cost, double *max_cost, int32_t *index_max);

/* Exact-parity entry: sample time `k`, n_candidates estimators against ONE set of supplied
 * post-transform noise draws (host pointers, trial index fastest):
 *   x0noise[2][B]  draw of N(0,P0) added to x0          (simulator.cpp:12)
 *   w[T][2][B]     process-noise draws N(0,Qd_truth)    (system_models.hpp:87,96)
 *   v[T][B]        observation-noise draws N(0,R_truth) (system_models.hpp:158)
 * Outputs (either may be NULL): per_trial[n_candidates][4][B] = {row_nees, row_nis, var_nees,
 * var_nis} of trial.cpp:62-63,95-96; sums[n_candidates][4] = their sums over the B trials.
 * Copies host->device and back inside the call. */
int kfbo_eval_supplied(kfbo_handle *h, int32_t k, int32_t n_candidates, const double *q_est,
                       const double *r_est, int64_t B, const double *x0noise, const double *w,
                       const double *v, double *per_trial, double *sums);

/* Same with DEVICE pointers for the noise (already resident in HBM) and, optionally, for the
 * per-trial output; sums is a host pointer.  Used for the HBM-roofline measurement. */
int kfbo_eval_supplied_dev(kfbo_handle *h, int32_t k, int32_t n_candidates, const double *q_est,
                           const double *r_est, int64_t B, const double *d_x0noise, const double *d_w,
                           const double *d_v, double *d_per_trial, double *sums);

/* Philox-mode per-trial outputs for draw-for-draw checks: trials [trial0, trial0+ntrials) of
 * sample time k on Philox stream `stream`; per_trial[4][ntrials] (host). */
int kfbo_eval_philox_trials(kfbo_handle *h, int32_t k, double q_est, double r_est, uint64_t stream,
                            uint32_t trial0, int64_t ntrials, double *per_trial);

/* The Nsimruns == 1 path of Trial::Run (trial.cpp:69-78,116-129: the error inside its 95 % band): trajectory
 * `trial` of Philox stream `stream` at sample time k, replayed step by step.
 * trace[T][9] (host) = {time, e1, e2, lb1, ub1, lb2, ub2, nees, nis}, lb/ub = -/+ 2 sqrt(P_ii) of the posterior. */
#define KFBO_TRACE_COLUMNS 9
int kfbo_trace_trial(kfbo_handle *h, int32_t k, double q_est, double r_est, uint64_t stream, uint32_t trial, double *trace);

/* Philox stream bookkeeping: evaluation e of candidate c at sample time k uses stream
 * base + (e*max_candidates + c)*n_sample_times + k, a KFBO_STREAM_BITS-bit id (modulo 2^62): it never wraps in
 * practice (a 4096-candidate handle could run 5e14 evaluations).  kfbo_eval* advance e by one per call.
 * Callers that hand out stream bases themselves keep them apart by the top bits: the Trial adapters of
 * include/kfbo_trial.hpp start at KFBO_TRIAL_STREAM_BASE, RunTrial / kfbo_eval at 0. */
#define KFBO_STREAM_BITS 62
#define KFBO_TRIAL_STREAM_BASE (1ull << 61)
int kfbo_set_stream_base(kfbo_handle *h, uint64_t base);
int kfbo_set_eval_counter(kfbo_handle *h, uint64_t e);

/* Raw cost sums of the last kfbo_eval / kfbo_eval_batch: sums[n_candidates][D][4] =
 * {sum row_nees, sum row_nis, sum var_nees, sum var_nis} over the trials (all ranks). */
int kfbo_last_sums(const kfbo_handle *h, double *sums, int32_t n_candidates);

/* Device time (ms, CUDA events on the handle's stream) of the last rollout kernel launch(es)
 * and how many kernels the last evaluation launched. */
int kfbo_last_kernel_ms(const kfbo_handle *h, double *ms, int32_t *launches);

/* Bytes the last kfbo_eval* moved host -> device (candidate constants, exchange header) and device -> host (sums or
 * compact costs; sums the kernel stores straight into mapped host memory count too). */
int kfbo_last_transfer_bytes(const kfbo_handle *h, int64_t *h2d, int64_t *d2h);

/* Sum of the rollout-kernel device times of the evaluations since the last reset, and how many (a benchmark loop reads
 * this once at its end instead of calling kfbo_last_kernel_ms between evaluations).  The first call switches the
 * handle to accumulating. */
int kfbo_kernel_ms_total(kfbo_handle *h, double *ms_total, int64_t *evaluations, int32_t reset);

/* Host wall time (microseconds) of the phases of the last kfbo_eval / kfbo_eval_batch on this rank:
 * us[0] building the per-(candidate, sample time) constants, us[1] enqueueing copies and launches,
 * us[2] waiting for the device (kernels, all-reduce, result copy), us[3] cost assembly (trial_node.cpp:70-119);
 * us[4] is device time: the all-reduce of the cost sums (CUDA events around it; 0 without a communicator).
 * The reference has no counterpart; this is what separates host overhead from kernel time in bench.py. */
#define KFBO_HOST_PHASES 5
int kfbo_last_host_us(const kfbo_handle *h, double us[KFBO_HOST_PHASES]);

/* Run subsequent launches on this CUDA stream (a cudaStream_t); NULL = the handle's own. */
int kfbo_set_cuda_stream(kfbo_handle *h, void *cuda_stream);

/* Tuning knobs (none changes results).  KFBO_OPT_SUPPLIED_KERNEL: which supplied-noise kernel
 * kfbo_eval_supplied* launch -- AUTO = the bulk-copy (cp.async.bulk + mbarrier) kernel whenever the
 * streams are 16-byte aligned and B is even, else the per-thread-load kernel. */
enum { KFBO_OPT_SUPPLIED_KERNEL = 1, KFBO_OPT_BULK_SMEM_PAD = 2 /* extra dynamic shared memory (bytes) per CTA of the
                                        bulk-copy kernel: lowers the CTAs resident per SM */,
       KFBO_OPT_ROLLOUT_FORM = 3,
       KFBO_OPT_REDUCE = 4 /* how the cost sums of a multi-GPU evaluation meet, KFBO_REDUCE_* below */,
       KFBO_OPT_PEER_FUSE = 5 /* KFBO_PEER_FUSE_* below */,
       KFBO_OPT_PEER_TIMEOUT_MS = 6 /* how long the exchange waits for a rank before KFBO_EPEER (default 10000) */,
       KFBO_OPT_CUDA_GRAPH = 8 /* 1 (default): a single-candidate evaluation -- constants copy, rollout, exchange, timing events --
                                        is replayed as ONE CUDA graph launch; 0: enqueued call by call */,
       KFBO_OPT_HOST_POLL = 9 /* 1 (default): where the kernel stores the total into mapped host memory (multi-GPU, trials
                                        sharded) the host polls that memory for it; 0: cudaStreamSynchronize */,
       KFBO_OPT_SMALL_LAUNCH = 10 /* 1 (default): an on-device-noise launch of at most 2 CTAs per SM -- the reference's default
                                        200 x 201 x 2 workload is 14 -- runs as the LATENCY form of the rollout: 32 trajectories
                                        per CTA, one warp runs the filter recursion, three make its noise a step group
                                        ahead.  Same streams, same arithmetic, same order of the sums: same bits. 0: never */,
       KFBO_OPT_AUG_LAYOUT = 11 /* state_dim 3, 4 with on-device noise: KFBO_AUG_LAYOUT_THREAD (default) one trajectory per
                                        thread, the whole covariance in its registers; KFBO_AUG_LAYOUT_LANES one covariance
                                        row per lane of a 4-lane group (the layout that scales beyond N = 4; supplied noise
                                        always runs it).  Same Philox words, same arithmetic per row: same bits. */,
       KFBO_OPT_GROUP_THREADS = 12 /* kfbo_group_set_option only. 1 (default): a device group runs the launch, the exchange and the
                                        wait of an evaluation on all its devices at once, one persistent host thread per device
                                        beyond the first (they spin for 200 us after an evaluation, then sleep); 0: the calling
                                        thread serves one device after the other */,
       KFBO_OPT_INJECT_FAULT = 13 /* fault injection for tests: 1 makes the NEXT evaluation of this handle fail with KFBO_ECUDA before
                                        anything is launched on its device -- a rank (or, through kfbo_group_handle, a device
                                        of a group) dropping out of a multi-GPU evaluation.  The others then see KFBO_EPEER
                                        after KFBO_OPT_PEER_TIMEOUT_MS and every handle involved refuses further evaluations
                                        with KFBO_ESTATE (tests/test_gpu_multi.py) */,
       KFBO_OPT_RESERVE_SUPPLIED_TRIALS = 7 /* pre-size the scratch of the parity entry points (kfbo_eval_supplied,
                                        kfbo_eval_philox_trials, kfbo_trace_trial) for this many trials x max_candidates
                                        candidates: after it they allocate nothing, like kfbo_eval* */ };
/* KFBO_OPT_ROLLOUT_FORM: only KFBO_FORM_REFERENCE exists -- every trajectory carries truth state, estimate and covariance and
 * runs simulator.cpp:15-19 + kalman_filter.cpp:18-38 + trial.cpp:53-60 step by step, as north_star prescribes.  (Round 1 also
 * had a "reduced" form -- gain recursion tabulated once per case, error-state trajectories; it was never more than 6 % faster
 * and was removed: DESIGN.md section 4a.  The option number stays reserved.) */
enum { KFBO_FORM_REFERENCE = 0 };
enum { KFBO_AUG_LAYOUT_LANES = 0, KFBO_AUG_LAYOUT_THREAD = 1 };
enum { KFBO_SUPPLIED_AUTO = 0, KFBO_SUPPLIED_LDG = 1, KFBO_SUPPLIED_BULK = 2 };
/* KFBO_OPT_REDUCE (set it to the same value on every rank):
 *   KFBO_REDUCE_AUTO (default): the peer-memory exchange (below) when the ranks are attached to each other, else NCCL;
 *   KFBO_REDUCE_NCCL: always one ncclAllReduce.
 * KFBO_OPT_PEER_FUSE: where the exchange runs -- KFBO_PEER_FUSE_AUTO (default) and KFBO_PEER_FUSE_NEVER: as a small
 *   kernel launched behind the rollout on the same stream; KFBO_PEER_FUSE_ALWAYS: in the tail of the rollout kernel itself
 *   (the CTA that finishes the evaluation's last case runs it).  The fused form is the slower one on B200 -- the rollout
 *   loop is FP64-issue bound and ptxas schedules it differently when the exchange follows in the same kernel (+0.9 % /
 *   +2.0 % kernel time for one candidate / a batch, against 2-3 us for the second launch; profiles/NOTES_r1.md) -- so
 *   it stays an option, and the single-GPU kernel is the same machine code either way. */
enum { KFBO_REDUCE_AUTO = 0, KFBO_REDUCE_NCCL = 1 };
enum { KFBO_PEER_FUSE_NEVER = 0, KFBO_PEER_FUSE_AUTO = 1, KFBO_PEER_FUSE_ALWAYS = 2 };
int kfbo_set_option(kfbo_handle *h, int32_t option, int64_t value);

/* ---- multi-GPU: one process per GPU; the cost sums of an evaluation meet in ONE exchange (peer memory or NCCL) ---- */
#define KFBO_NCCL_ID_BYTES 128
int kfbo_nccl_unique_id(void *id_bytes /*[KFBO_NCCL_ID_BYTES]*/);
int kfbo_comm_init(kfbo_handle *h, const void *id_bytes, int32_t rank, int32_t nranks, int32_t shard_mode);
/* The exchange of the cost sums over peer memory (NVLink / NVSwitch) instead of the NCCL call: behind the rollout, each
 * rank stores its share of the [C][D][4] sums into an exchange buffer on every GPU of the box, raises a flag there,
 * waits for the other ranks' flags and -- trials sharded -- adds the ranks' shares in rank order (the total is the same
 * bit pattern on every rank and from run to run; candidates sharded: every slot has one owner and nothing is added).
 * No NCCL launch, no host round trip; two generations of buffers, so consecutive evaluations need no barrier.
 * Ranks on one box (at most KFBO_MAX_PEERS), after kfbo_comm_init:
 *   kfbo_peer_export  allocates this rank's exchange buffer and returns its CUDA IPC handle;
 *   kfbo_peer_attach  takes the handles of ALL ranks in rank order (exchanged by the caller, e.g. an all-gather over
 *                     torch.distributed -- kf_bayesopt_b200/dist.py) and maps the peers' buffers.
 * Every rank must make the same sequence of kfbo_eval* calls (as for NCCL); a rank that waits longer than
 * KFBO_OPT_PEER_TIMEOUT_MS for another returns KFBO_EPEER instead of hanging (the ranks are then out of step with each
 * other: the handle -- or the whole device group, after every device has been drained -- refuses further evaluations
 * with KFBO_ESTATE; destroy it and start over, as after an NCCL failure).  A device group (below) attaches its
 * devices by itself.  Trial-sharded evaluations of more than 1024 (candidate, sample time) pairs keep the NCCL call. */
#define KFBO_MAX_PEERS 8
#define KFBO_PEER_HANDLE_BYTES 64
int kfbo_peer_export(kfbo_handle *h, void *handle_bytes /*[KFBO_PEER_HANDLE_BYTES]*/);
int kfbo_peer_attach(kfbo_handle *h, const void *all_handle_bytes /*[n_handles][KFBO_PEER_HANDLE_BYTES]*/,
                     int32_t n_handles /* must equal the rank count given to kfbo_comm_init */);
/* 0: the last kfbo_eval / kfbo_eval_batch ran on one rank; 1: its sums met in an ncclAllReduce; 2: in the peer-memory
 * exchange as a kernel of its own; 3: in the peer-memory exchange fused into the rollout kernel. */
int kfbo_last_reduce_path(const kfbo_handle *h);
/* [begin,end) of n units owned by `rank` of `nranks` (contiguous, remainder to the low ranks). */
int kfbo_shard_range(int64_t n, int32_t rank, int32_t nranks, int64_t *begin, int64_t *end);

/* ---- multi-GPU from ONE process: a device group ----
 * What a single-process caller (the ROS node robot1d_kf, src/test_trial.cpp) needs to spread an evaluation over
 * several GPUs of the box: one host thread drives n_devices devices (devices[i], or 0..n_devices-1 when NULL); each
 * evaluation launches its shard on every device; the [C][D][4] cost sums meet in the peer-memory exchange above (the
 * devices are attached to each other at creation when they can access each other's memory) or in ONE grouped
 * ncclAllReduce; then the costs are assembled from the first device's copy.  Sharding, Philox streams and therefore the results are those of
 * the one-process-per-GPU mode with n_devices ranks.  n_devices == 1 needs no NCCL. */
typedef struct kfbo_group kfbo_group;
int kfbo_group_create(const kfbo_params *p, int32_t n_devices, const int32_t *devices, int32_t shard_mode, kfbo_group **out);
void kfbo_group_destroy(kfbo_group *g);
const char *kfbo_group_last_error(const kfbo_group *g);      /* g == NULL: last kfbo_group_create failure of this thread */
int32_t kfbo_group_size(const kfbo_group *g);
kfbo_handle *kfbo_group_handle(kfbo_group *g, int32_t i);    /* the handle of device i (owned by the group) */
int kfbo_group_eval(kfbo_group *g, const double *pn, int32_t npn, const double *on, int32_t non, kfbo_result *out);
int kfbo_group_eval_batch(kfbo_group *g, int32_t n_candidates, const double *q_est, const double *r_est, kfbo_result *out);
int kfbo_group_eval_batch_costs(kfbo_group *g, int32_t n_candidates, const double *q_est, const double *r_est,
                                double *cost, double *max_cost, int32_t *index_max);   /* kfbo_eval_batch_costs on a group */
int kfbo_group_last_sums(const kfbo_group *g, double *sums, int32_t n_candidates);
int kfbo_group_last_kernel_ms(const kfbo_group *g, double *ms_max /* slowest device */, int32_t *launches /* all devices */);
int kfbo_group_set_option(kfbo_group *g, int32_t option, int64_t value);
int kfbo_group_set_stream_base(kfbo_group *g, uint64_t base);
int kfbo_group_set_eval_counter(kfbo_group *g, uint64_t e);

/* ---- host-side pieces of the path, exported so they can be tested without a GPU ---- */
/* continuousToDiscrete for the robot1d model (continuous2discrete.hpp:8-32, system_models.hpp:33-39,58). */
int kfbo_discretize(double pnoise0, double dt, double *Fd /*[4]*/, double *Qd /*[4]*/);
/* u_[i] = 2cos(0.75 j), j += dt (system_models.hpp:49-56). */
int kfbo_control_table(double dt, int32_t n, double *u);
/* Thread averages + costs + max of trial.cpp:19,110-113 and trial_node.cpp:70-119 from the raw sums
 * sums[D][4] over the (nsimruns/cpu_core_number)*cpu_core_number trials actually run. */
int kfbo_finalize(const kfbo_params *p, const double *sums, kfbo_result *out);
/* Trials actually run per sample time: (nsimruns / cpu_core_number) * cpu_core_number (trial.cpp:28). */
int64_t kfbo_effective_trials(const kfbo_params *p);

/* DFMA-chain microbenchmark: sustained FP64 FMA throughput of `device` in TFLOP/s (2 flops per FMA)
 * over about `millis` ms -- the roofline denominator for the Philox path.  *tflops: chains x = fma(x, y, U) with the
 * multiplier in a register (the full-rate operand form, 2.0 pipe cycles per warp instruction); *tflops_uniform_operands
 * (optional): x = fma(x, U, U), both other operands in the parameter bank, which the pipe runs ~10 % slower. */
int kfbo_measure_fp64_peak(int32_t device, double millis, double *tflops, double *tflops_uniform_operands);

const char *kfbo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KFBO_H_ */
