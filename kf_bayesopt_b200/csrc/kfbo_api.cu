// kfbo_api.cu -- the C ABI (include/kfbo.h): handle, host orchestration, the cross-GPU exchange of the cost sums
// (peer memory over NVLink / NVSwitch: kfbo_peer_*, kfbo_group_create; NCCL all-reduce as the fallback).
//
// Replaces the CPUcoreNumber std::thread pool of trial_node.cpp:54-67 with a device-side batch
// launcher over trials x candidates x sample times.  No CPU fallback: every evaluation runs the
// sm_100a kernels of kfbo_kernels.cu or fails with an error code.
#include <dlfcn.h>
#include <sched.h>
#include <nccl.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <functional>
#include <limits>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "kfbo_host.h"
#include "kfbo_kernels.h"
#include "kfbo_math.cuh"

using kfbo::Case;

namespace {

thread_local std::string g_create_error;

// ---- NCCL, loaded on first use so single-GPU users need no libnccl at all ----------------
struct NcclApi {
    void *so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        a.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.so) a.so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.so) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.so, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.so, "ncclCommInitRank");
        a.AllReduce = (decltype(a.AllReduce))dlsym(a.so, "ncclAllReduce");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.so, "ncclCommDestroy");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.so, "ncclGetErrorString");
        a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.so, "ncclCommInitAll");
        a.GroupStart = (decltype(a.GroupStart))dlsym(a.so, "ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.so, "ncclGroupEnd");
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString && a.CommInitAll &&
               a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}

constexpr size_t kPartialsBudgetBytes = 64u << 20;  // per-launch partial-sum scratch cap
constexpr size_t kXchgFlagBytes = 256;              // flags [2][kMaxPeers] in front of the exchange slabs
constexpr size_t kZeroCopyMaxDoubles = 1024;        // results up to this size are stored straight into mapped host memory by the kernel
constexpr size_t kPeerMaxSummedDoubles = 4096;      // trials sharded: one CTA adds nranks slabs of this many doubles at most
static_assert(KFBO_MAX_PEERS == kfbo::kMaxPeers, "kfbo.h and kfbo_kernels.h disagree on the peer count");
static_assert(KFBO_PEER_HANDLE_BYTES == sizeof(cudaIpcMemHandle_t), "cudaIpcMemHandle_t size");
static_assert(KFBO_AUG_LAYOUT_LANES == kfbo::kAugLayoutLanes && KFBO_AUG_LAYOUT_THREAD == kfbo::kAugLayoutThread, "kfbo.h and kfbo_kernels.h disagree on the layouts");
static_assert(2 * kfbo::kMaxPeers * sizeof(unsigned) <= kXchgFlagBytes, "flag block");

}  // namespace

struct kfbo_handle {
    kfbo_params p{};
    int device = 0;
    int64_t b_eff = 0;
    int D = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    // tables
    double2 *d_gu = nullptr;          // [D][kMaxSteps] = (g0*u, g1*u)
    kfbo::math::LogTables *d_logtab = nullptr;   // -2 ln(u) tables of the noise generator
    double truth_l[KFBO_MAX_SAMPLE_TIMES][3]{};
    double truth_sr[KFBO_MAX_SAMPLE_TIMES]{};
    double fd01[KFBO_MAX_SAMPLE_TIMES]{};
    int max_T = 0, min_T = kfbo::kMaxSteps;
    kfbo::SampleTimeConsts st{};       // per-sample-time constants, passed to the kernels by value
    // per-evaluation buffers (sized at create).  The case constants sit behind a kPeerCtxBytes header that carries the
    // PeerCtx of a multi-GPU evaluation (one host-to-device copy brings both): d_cases = d_case_buf + kPeerCtxBytes
    // A batch of more than one candidate sends 16 bytes per candidate instead -- (pnoise, onoise), right-aligned against
    // the header so that header and candidates are still one copy -- and build_cases_kernel makes the cases on the device:
    //   [ candidates: max_candidates x double2, used from the END ][ PeerCtx header ][ cases ]
    unsigned char *d_case_buf = nullptr, *h_case_buf = nullptr;
    unsigned char *d_hdr = nullptr, *h_hdr = nullptr;     // the header inside the two buffers
    Case *d_cases = nullptr, *h_cases = nullptr;
    // compact cost output of kfbo_eval_batch_costs: cost[C][D], max_cost[C], index_max[C], assembled on the device
    void *d_compact = nullptr, *h_compact = nullptr;
    bool compact = false;                  // the evaluation in flight ends in assemble_costs_kernel instead of the host's finalize
    int64_t last_h2d = 0, last_d2h = 0;    // bytes the last evaluation moved between host and device
    int cases_cap = 0;
    double *d_partials = nullptr;
    size_t partials_cap = 0;          // doubles
    unsigned *d_tickets = nullptr;
    double *d_sums = nullptr, *h_sums = nullptr;
    double *h_sums_dev = nullptr;     // h_sums as the device sees it (mapped pinned memory): small results are stored there by the kernel
    // grow-only scratch for the supplied-noise / per-trial entry points
    double *d_scratch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_cap[5] = {0, 0, 0, 0, 0};
    // memo of the estimator-side discretisation by (pnoise, sample time): a candidate sweep repeats each
    // pnoise value many times (a 64 x 64 grid has 64 of them), and the 4x4 matrix exponential is the only
    // non-trivial host work per case
    struct DiscMemo { uint64_t qbits = 0; int k = -1; double f = 0, Qd[4] = {0, 0, 0, 0}; };
    DiscMemo disc_memo[256];
    // the state-augmented variants (kfbo_params.state_dim = 3, 4; kfbo_aug.cuh): their own case / table types
    int N = 2;                                   // state dimension
    int aug_layout = KFBO_AUG_LAYOUT_THREAD;     // KFBO_OPT_AUG_LAYOUT: which kernel runs an on-device-noise evaluation of N = 3, 4
    kfbo::AugSampleTime h_aug_st[KFBO_MAX_SAMPLE_TIMES]{};
    kfbo::AugSampleTime *d_aug_st = nullptr;     // [D]
    double *d_u = nullptr;                       // [D][kMaxSteps] control table u_[i] per sample time
    kfbo::AugCase *d_aug_cases = nullptr, *h_aug_cases = nullptr;   // [cases_cap]
    std::vector<kfbo::AugCase> sup_aug_cases;
    // NCCL
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, shard_mode = KFBO_SHARD_TRIALS;
    int host_procs = 1;                    // processes that share this box's host cores with this one (kfbo_comm_init: nranks; a group: 1)
    // peer-memory exchange of the cost sums (kfbo_peer_export / kfbo_peer_attach, or kfbo_group_create)
    unsigned char *d_xchg = nullptr;       // [kXchgFlagBytes: flags [2][kMaxPeers]] [vals [2][nranks][xchg_cap] doubles]
    size_t xchg_cap = 0;                   // doubles per slab = cases_cap * 4
    unsigned char *peer_base[kfbo::kMaxPeers] = {};   // rank r's d_xchg as mapped in this process
    bool peer_opened[kfbo::kMaxPeers] = {};           // mapped through cudaIpcOpenMemHandle (to be closed)
    bool peer_ready = false;
    unsigned *d_done = nullptr;            // cases finished in the launch that carries the exchange
    int32_t *h_status = nullptr;           // pinned, written by the kernel: 1 + the rank whose flag did not arrive
    uint32_t peer_epoch = 0;
    int reduce_pref = KFBO_REDUCE_AUTO, peer_fuse = KFBO_PEER_FUSE_AUTO;
    int64_t peer_timeout_ms = 10000;
    int reduce_path = 0;                   // of the evaluation in flight / the last one: kfbo_last_reduce_path
    const double *result_src = nullptr;    // where the evaluation in flight leaves the total sums on this device
    // Philox stream bookkeeping (62-bit stream ids, kfbo.h: KFBO_STREAM_BITS)
    uint64_t stream_base = 0, eval_counter = 0;
    // candidates of the evaluation in flight whose (pnoise, onoise) lie outside the domain (non-finite or <= 0): their
    // costs are NaN, like the "cost" message the reference publishes for them (eval_check / finalize_range)
    std::vector<uint8_t> cand_bad;
    int n_bad = 0;
    // set when a multi-rank evaluation failed part-way (KFBO_EPEER, a CUDA / NCCL error after the launch): the ranks'
    // evaluation counters and exchange generations may no longer agree, so every later evaluation fails fast
    bool poisoned = false;
    std::string poison_why;
    bool inject_fault = false;             // KFBO_OPT_INJECT_FAULT (tests): the next launch of this handle fails before anything is enqueued
    std::vector<Case> sup_cases;           // host scratch of the supplied-noise entry points (grow-only)
    std::vector<double> sup_sums;
    int supplied_kernel = kfbo::kSuppliedAuto;
    int bulk_smem_pad = 0;
    // The single-candidate evaluation ("noise" -> "cost") replayed from a CUDA graph: constants copy, rollout, exchange and the
    // three timing events are ONE cudaGraphLaunch; between replays only the by-value estimator constants of the rollout
    // kernel node change (cudaGraphExecKernelNodeSetParams).  Rebuilt when anything else about the launch changes (gsig).
    bool use_graph = true, capturing = false, graph_used = false;
    cudaGraphExec_t gexec = nullptr;
    cudaGraphNode_t gnode = nullptr;           // the rollout kernel node
    kfbo::PhiloxLaunchRecord grec;
    struct GraphSig {
        cudaStream_t stream = nullptr; const void *sums = nullptr; int ncases = -1; uint32_t ntrials = 0, t0 = 0; bool peer = false, zero_copy = false;
        bool operator==(const GraphSig &o) const
        {
            return stream == o.stream && sums == o.sums && ncases == o.ncases && ntrials == o.ntrials && t0 == o.t0 && peer == o.peer &&
                   zero_copy == o.zero_copy;
        }
    } gsig;
    int sm_count = 0;                          // a launch of at most 2 CTAs per SM runs as the latency form of the rollout (rollout_philox_small)
    bool use_small = true;
    // the host polls this word of mapped pinned memory for the exchange kernel's "total stored" instead of synchronising the stream
    unsigned *h_flag = nullptr, *h_flag_dev = nullptr;
    uint32_t flag_expect = 0;
    bool flag_armed = false, use_poll = true;
    // kernel timing is read back lazily (cudaEventElapsedTime costs a microsecond each): on query, or -- accumulating -- at the next launch
    bool timing_pending = false, timing_accumulate = false;
    double kernel_ms_total = 0.0;
    int64_t kernel_ms_evals = 0;
    // last evaluation
    int last_ncand = 0, last_launches = 0;
    double last_ms = 0.0;
    struct kfbo_assembly_pool *pool = nullptr;   // host threads for the cost assembly of large batches (created on first use)
    std::chrono::steady_clock::time_point tp[3];   // phase marks of the evaluation in flight
    double last_host_us[KFBO_HOST_PHASES] = {0, 0, 0, 0, 0};   // build cases, enqueue, wait for the device, cost assembly, all-reduce (device)
    std::string err;
};

namespace {

int fail(kfbo_handle *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define KFBO_CUDA(h, call)                                                                          \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail((h), KFBO_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// The library is meant to sit next to other CUDA users of the calling thread (torch: kfbo_set_cuda_stream and the
// *_dev entry points take its pointers): every entry point that makes the handle's device current puts the caller's
// device back on every return path.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            switched = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define KFBO_ON_DEVICE(h)                                                                           \
    DeviceGuard dg_((h)->device);                                                                   \
    if (dg_.err != cudaSuccess)                                                                     \
        return fail((h), KFBO_ECUDA, "cudaSetDevice(%d) failed: %s", (h)->device, cudaGetErrorString(dg_.err))

#define KFBO_NCCL(h, call)                                                                          \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess)                                                                      \
            return fail((h), KFBO_ENCCL, "%s failed: %s (%s:%d)", #call, nccl().GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

int ensure_scratch(kfbo_handle *h, int slot, size_t doubles)
{
    if (h->scratch_cap[slot] >= doubles) return KFBO_OK;
    if (h->d_scratch[slot]) KFBO_CUDA(h, cudaFree(h->d_scratch[slot]));
    h->d_scratch[slot] = nullptr;
    h->scratch_cap[slot] = 0;
    KFBO_CUDA(h, cudaMalloc(&h->d_scratch[slot], doubles * sizeof(double)));
    h->scratch_cap[slot] = doubles;
    return KFBO_OK;
}

// scratch slot 3 of the supplied-noise entry points: partials, sums, cases, tickets of C candidates
constexpr size_t kCaseDoubles = (std::max(sizeof(Case), sizeof(kfbo::AugCase)) + 7) / 8;   // room for either case type
size_t supplied_scratch3_doubles(int C, int tiles)
{
    return (size_t)C * tiles * 4 + (size_t)C * 4 + (size_t)C * kCaseDoubles + C;
}

// The constants of (candidate, sample time k) of a state-augmented evaluation.
int build_aug_case(kfbo_handle *h, int k, double q_est, double r_est, uint64_t stream, int out_slot, kfbo::AugCase *c)
{
    double Fd[16];
    kfbo::host::discretize_chain(h->N, q_est, h->p.dt[k], Fd, c->qd);
    c->r_est = r_est / h->p.dt[k];
    c->T = h->p.max_iterations[k];
    c->stream = (uint32_t)stream;
    c->stream_x = (uint32_t)((stream >> 32) & 0x3FFFFFFFull) << 2;
    c->out_slot = out_slot;
    return KFBO_OK;
}

// Constants of (candidate q_est/r_est, sample time k): ProcessModel / ObservationModel of the
// estimator (kalman_filter.cpp:6-10) plus the truth-side noise transform (simulator.cpp:5).
int build_case(kfbo_handle *h, int k, double q_est, double r_est, uint64_t stream, int out_slot, Case *c)
{
    uint64_t qbits;
    std::memcpy(&qbits, &q_est, sizeof qbits);
    kfbo_handle::DiscMemo &memo = h->disc_memo[((qbits * 0x9E3779B97F4A7C15ull) >> 56) ^ (uint64_t)k];
    if (memo.k != k || memo.qbits != qbits) {
        double Fd[4];
        kfbo::host::discretize(q_est, h->p.dt[k], Fd, memo.Qd);
        if (Fd[0] != 1.0 || Fd[2] != 0.0 || Fd[3] != 1.0)
            return fail(h, KFBO_EINVAL, "discretised Fd lost its [[1,f],[0,1]] structure");
        memo.f = Fd[1]; memo.k = k; memo.qbits = qbits;
    }
    const double *Qd = memo.Qd;
    c->f = memo.f;
    c->q00 = Qd[0]; c->q01 = Qd[1]; c->q10 = Qd[2]; c->q11 = Qd[3];
    c->r_est = r_est / h->p.dt[k];                       // system_models.hpp:143
    c->l00 = h->truth_l[k][0]; c->l10 = h->truth_l[k][1]; c->l11 = h->truth_l[k][2];
    c->sr = h->truth_sr[k];
    c->T = h->p.max_iterations[k];
    kfbo::set_case_stream(c, stream);
    c->out_slot = out_slot;
    return KFBO_OK;
}

// Cost assembly of a batch (trial_node.cpp:70-119 per candidate).  ~30 ns per candidate (two logs per sample
// time), which for a 4096-candidate sweep is 0.13 ms after an evaluation of 1.5-11 ms.  Large batches are split over a
// few PERSISTENT host threads owned by the handle (candidates are independent; each thread writes its own slice of
// `out`); spawning threads per call costs more than it saves (measured: 147 us against 126 us single-threaded).
void nan_result(int D, kfbo_result *out)
{
    *out = kfbo_result{};
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int k = 0; k < D; ++k)
        out->average_nees[k] = out->average_nis[k] = out->average_var_nees[k] = out->average_var_nis[k] = out->cost[k] = nan;
    out->max_cost = nan;      // *max_element of all-NaN costs is the first one (trial_node.cpp:118-119)
    out->index_max = 0;
}

void finalize_range(const kfbo_handle *h, kfbo_result *out, int c0, int c1)
{
    const int D = h->D;
    for (int c = c0; c < c1; ++c) {
        if (h->n_bad && h->cand_bad[c]) nan_result(D, &out[c]);
        else kfbo::host::finalize(h->p, h->h_sums + (size_t)c * D * 4, &out[c]);
    }
}

}  // namespace

struct kfbo_assembly_pool {
    static constexpr int kMaxWorkers = 3;
    const int nworkers;
    std::thread th[kMaxWorkers];
    std::mutex mu;
    std::condition_variable cv;
    uint64_t posted = 0;               // job generation (guarded by mu)
    bool stop = false;
    std::atomic<int> done{0};
    const kfbo_handle *h = nullptr;    // the job
    kfbo_result *out = nullptr;
    int C = 0;

    explicit kfbo_assembly_pool(int n) : nworkers(std::max(0, std::min(n, (int)kMaxWorkers)))
    {
        for (int i = 0; i < nworkers; ++i) th[i] = std::thread([this, i] { work(i + 1); });
    }
    ~kfbo_assembly_pool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (int i = 0; i < nworkers; ++i) th[i].join();
    }
    void slice(int CC, int part, int *c0, int *c1) const
    {
        const int per = (CC + nworkers) / (nworkers + 1);
        *c0 = std::min(CC, part * per);
        *c1 = std::min(CC, (part + 1) * per);
    }
    void work(int part)
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || posted != seen; });
                if (stop) return;
                seen = posted;
            }
            int c0, c1;
            slice(C, part, &c0, &c1);
            finalize_range(h, out, c0, c1);
            done.fetch_add(1, std::memory_order_release);
        }
    }
    void run(const kfbo_handle *hh, int CC, kfbo_result *o)
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            h = hh; C = CC; out = o;
            done.store(0, std::memory_order_relaxed);
            ++posted;
        }
        cv.notify_all();
        int c0, c1;
        slice(CC, 0, &c0, &c1);
        finalize_range(hh, o, c0, c1);
        while (done.load(std::memory_order_acquire) != nworkers) std::this_thread::yield();
    }
};

namespace {

void finalize_all(kfbo_handle *h, int C, kfbo_result *out)
{
    if (C < 1024) { finalize_range(h, out, 0, C); return; }
    if (!h->pool) {
        // Workers beside the calling thread: what this process' share of the host cores allows.  One process per GPU
        // means that many processes assemble at the same moment (the exchange ends on all of them together): on a
        // 16-core box with 8 ranks four threads each would be 32 runnable threads, the callers among them spinning.
        cpu_set_t set;
        int cores = (int)std::thread::hardware_concurrency();
        if (sched_getaffinity(0, sizeof set, &set) == 0) cores = CPU_COUNT(&set);
        const int share = std::max(1, cores / std::max(1, h->host_procs));
        h->pool = new (std::nothrow) kfbo_assembly_pool(share - 1);
    }
    if (!h->pool || h->pool->nworkers == 0) { finalize_range(h, out, 0, C); return; }
    h->pool->run(h, C, out);
}

int cases_per_launch(const kfbo_handle *h, int tiles)
{
    const size_t per_case = (size_t)tiles * 4;
    size_t n = h->partials_cap / per_case;
    if (n < 1) n = 1;
    if (n > (size_t)h->cases_cap) n = (size_t)h->cases_cap;
    // whole candidates only (grid.z = D sample times per candidate), grid.y <= 65535 candidates
    const size_t D = (size_t)h->D;
    if (n > 65535 * D) n = 65535 * D;
    n = n / D * D;
    if (n < D) n = D;
    return (int)n;
}

}  // namespace

extern "C" {

const char *kfbo_version(void) { return "kfbo-b200 0.1 (sm_100a)"; }

void kfbo_default_params(kfbo_params *p)
{
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->n_sample_times = 2;                       // it_num (trial_node.cpp:49)
    p->dt[0] = 0.1; p->max_iterations[0] = 201;  // vehicleParams.yaml:8-10, read_para_vehicle.cpp:145
    p->dt[1] = 1.0; p->max_iterations[1] = 201;  // trial_node.cpp:105-114
    p->nsimruns = 200;                           // vehicleParams.yaml:16
    p->cpu_core_number = 10;                     // vehicleParams.yaml:1
    p->state_dof = 2; p->observation_dof = 1;    // vehicleParams.yaml:4-5
    p->pnoise_truth = 1.0; p->onoise_truth = 0.1;  // vehicleParams.yaml:18-19
    p->cost_choice = KFBO_JNEES;                 // vehicleParams.yaml:25
    p->device = -1;
    p->seed = 0x4B46424F20260101ull;
    p->max_candidates = 1;
}

int64_t kfbo_effective_trials(const kfbo_params *p)
{
    if (!p || p->cpu_core_number < 1) return KFBO_EINVAL;
    return kfbo::host::effective_trials(*p);
}

int kfbo_discretize(double pnoise0, double dt, double *Fd, double *Qd)
{
    if (!Fd || !Qd) return KFBO_EINVAL;
    kfbo::host::discretize(pnoise0, dt, Fd, Qd);
    return KFBO_OK;
}

int kfbo_control_table(double dt, int32_t n, double *u)
{
    if (!u || n < 0) return KFBO_EINVAL;
    kfbo::host::control_table(dt, n, u);
    return KFBO_OK;
}

int kfbo_finalize(const kfbo_params *p, const double *sums, kfbo_result *out)
{
    if (!p || !sums || !out) return KFBO_EINVAL;
    const int rc = kfbo::host::validate(*p);
    if (rc) return rc;
    kfbo::host::finalize(*p, sums, out);
    return KFBO_OK;
}

int kfbo_shard_range(int64_t n, int32_t rank, int32_t nranks, int64_t *begin, int64_t *end)
{
    if (!begin || !end || n < 0 || nranks < 1 || rank < 0 || rank >= nranks) return KFBO_EINVAL;
    kfbo::host::shard_range(n, rank, nranks, begin, end);
    return KFBO_OK;
}

const char *kfbo_last_error(const kfbo_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void kfbo_destroy(kfbo_handle *h)
{
    if (!h) return;
    delete h->pool;
    DeviceGuard dg(h->device);           // the caller's current device is put back on return
    if (h->comm && nccl().ok) nccl().CommDestroy(h->comm);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    for (int i = 0; i < 5; ++i) if (h->d_scratch[i]) cudaFree(h->d_scratch[i]);
    if (h->d_gu) cudaFree(h->d_gu);
    if (h->d_logtab) cudaFree(h->d_logtab);
    for (int r = 0; r < kfbo::kMaxPeers; ++r)
        if (h->peer_opened[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
    if (h->d_xchg) cudaFree(h->d_xchg);
    if (h->d_done) cudaFree(h->d_done);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->d_case_buf) cudaFree(h->d_case_buf);
    if (h->h_case_buf) cudaFreeHost(h->h_case_buf);
    if (h->d_aug_st) cudaFree(h->d_aug_st);
    if (h->d_u) cudaFree(h->d_u);
    if (h->d_aug_cases) cudaFree(h->d_aug_cases);
    if (h->h_aug_cases) cudaFreeHost(h->h_aug_cases);
    if (h->d_compact) cudaFree(h->d_compact);
    if (h->h_compact) cudaFreeHost(h->h_compact);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    if (h->h_flag) cudaFreeHost(h->h_flag);
    if (h->d_partials) cudaFree(h->d_partials);
    if (h->d_tickets) cudaFree(h->d_tickets);
    if (h->d_sums) cudaFree(h->d_sums);
    if (h->h_sums) cudaFreeHost(h->h_sums);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int kfbo_create(const kfbo_params *p, kfbo_handle **out)
{
    if (!p || !out) return fail(nullptr, KFBO_EINVAL, "kfbo_create: NULL argument");
    *out = nullptr;
    int rc = kfbo::host::validate(*p);
    if (rc) return fail(nullptr, rc, "kfbo_create: invalid kfbo_params (code %d)", rc);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev < 1)
        return fail(nullptr, KFBO_ENODEVICE, "kfbo_create: no CUDA device (%s); there is no CPU fallback",
                    ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    kfbo_handle *h = new (std::nothrow) kfbo_handle;
    if (!h) return fail(nullptr, KFBO_EINVAL, "kfbo_create: out of host memory");
    h->p = *p;
    h->D = p->n_sample_times;
    h->b_eff = kfbo::host::effective_trials(*p);
    auto bail = [&](int code) { g_create_error = h->err; kfbo_destroy(h); return code; };
    auto cu = [&](cudaError_t e, const char *what) {
        if (e == cudaSuccess) return false;
        fail(h, KFBO_ECUDA, "kfbo_create: %s failed: %s", what, cudaGetErrorString(e));
        return true;
    };
    if (p->device >= 0) h->device = p->device;
    else if (cu(cudaGetDevice(&h->device), "cudaGetDevice")) return bail(KFBO_ECUDA);
    if (h->device >= ndev) { fail(h, KFBO_ENODEVICE, "kfbo_create: device %d of %d", h->device, ndev); return bail(KFBO_ENODEVICE); }
    DeviceGuard dg(h->device);           // the caller's current device is put back on return
    if (cu(dg.err, "cudaSetDevice")) return bail(KFBO_ECUDA);
    if (cu(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return bail(KFBO_ECUDA);
    h->stream = h->own_stream;
    if (cu(cudaEventCreate(&h->ev0), "cudaEventCreate") || cu(cudaEventCreate(&h->ev1), "cudaEventCreate") ||
        cu(cudaEventCreate(&h->ev2), "cudaEventCreate")) return bail(KFBO_ECUDA);

    // ---- per-sample-time tables: g*u (system_models.hpp:46-56,82) and the truth noise transform
    std::vector<double2> gu((size_t)h->D * kfbo::kMaxSteps);
    std::vector<double> u(kfbo::kMaxSteps);
    for (int k = 0; k < h->D; ++k) {
        const double dt = p->dt[k];
        const double g0 = 0.5 * dt * dt, g1 = dt;   // system_models.hpp:46-47
        kfbo::host::control_table(dt, kfbo::kMaxSteps, u.data());
        for (int i = 0; i < kfbo::kMaxSteps; ++i) gu[(size_t)k * kfbo::kMaxSteps + i] = make_double2(g0 * u[i], g1 * u[i]);
        double Fd[4], Qd[4];
        kfbo::host::discretize(p->pnoise_truth, dt, Fd, Qd);
        h->fd01[k] = Fd[1];
        // lower Cholesky factor of Qd_truth; any T with T T^T = Qd is statistically equivalent to
        // eigenmvn.h:126-130's V sqrt(Lambda)
        h->truth_l[k][0] = std::sqrt(Qd[0]);
        h->truth_l[k][1] = Qd[2] / h->truth_l[k][0];
        h->truth_l[k][2] = std::sqrt(Qd[3] - h->truth_l[k][1] * h->truth_l[k][1]);
        h->truth_sr[k] = std::sqrt(p->onoise_truth / dt);  // R = onoise/dt (system_models.hpp:143)
        h->st.f[k] = Fd[1]; h->st.l00[k] = h->truth_l[k][0]; h->st.l10[k] = h->truth_l[k][1];
        h->st.l11[k] = h->truth_l[k][2]; h->st.sr[k] = h->truth_sr[k];
        h->max_T = std::max(h->max_T, p->max_iterations[k]);
        h->min_T = std::min(h->min_T, p->max_iterations[k]);
    }
    if (cu(cudaMalloc(&h->d_gu, gu.size() * sizeof(double2)), "cudaMalloc(gu)")) return bail(KFBO_ECUDA);
    if (cu(cudaMemcpy(h->d_gu, gu.data(), gu.size() * sizeof(double2), cudaMemcpyHostToDevice), "cudaMemcpy(gu)")) return bail(KFBO_ECUDA);
    {
        // built once per process (thread-safe: Trial::Run creates handles from several threads)
        static const kfbo::math::LogTables tabs = [] { kfbo::math::LogTables t; kfbo::math::build_log_tables(&t); return t; }();
        if (cu(cudaMalloc(&h->d_logtab, sizeof tabs), "cudaMalloc(logtab)")) return bail(KFBO_ECUDA);
        if (cu(cudaMemcpy(h->d_logtab, &tabs, sizeof tabs, cudaMemcpyHostToDevice), "cudaMemcpy(logtab)")) return bail(KFBO_ECUDA);
    }

    // ---- evaluation buffers
    h->cases_cap = p->max_candidates * h->D;
    h->cand_bad.assign((size_t)p->max_candidates, 0);
    h->N = p->state_dim >= 2 ? p->state_dim : 2;
    if (h->N > 2) {
        // the chain-of-integrators model per sample time (kfbo_host.h: discretize_chain)
        std::vector<double> ut((size_t)h->D * kfbo::kMaxSteps);
        for (int k = 0; k < h->D; ++k) {
            kfbo::AugSampleTime &a = h->h_aug_st[k];
            double Qd[16];
            kfbo::host::discretize_chain(h->N, p->pnoise_truth, p->dt[k], a.F, Qd);
            kfbo::host::control_vector_chain(h->N, p->dt[k], a.g);
            if (!kfbo::host::cholesky_chain(h->N, Qd, a.L)) { fail(h, KFBO_EINVAL, "kfbo_create: the truth Qd is not positive definite"); return bail(KFBO_EINVAL); }
            a.sr = std::sqrt(p->onoise_truth / p->dt[k]);
            kfbo::host::control_table(p->dt[k], kfbo::kMaxSteps, ut.data() + (size_t)k * kfbo::kMaxSteps);
        }
        if (cu(cudaMalloc(&h->d_aug_st, sizeof h->h_aug_st), "cudaMalloc(aug tables)") ||
            cu(cudaMemcpy(h->d_aug_st, h->h_aug_st, sizeof h->h_aug_st, cudaMemcpyHostToDevice), "cudaMemcpy(aug tables)") ||
            cu(cudaMalloc(&h->d_u, ut.size() * sizeof(double)), "cudaMalloc(u)") ||
            cu(cudaMemcpy(h->d_u, ut.data(), ut.size() * sizeof(double), cudaMemcpyHostToDevice), "cudaMemcpy(u)") ||
            cu(cudaMalloc(&h->d_aug_cases, (size_t)h->cases_cap * sizeof(kfbo::AugCase)), "cudaMalloc(aug cases)") ||
            cu(cudaMallocHost(&h->h_aug_cases, (size_t)h->cases_cap * sizeof(kfbo::AugCase)), "cudaMallocHost(aug cases)"))
            return bail(KFBO_ECUDA);
    }
    const int tiles = h->N > 2 ? kfbo::aug_tiles((size_t)h->b_eff) : kfbo::rollout_tiles((size_t)h->b_eff);
    size_t part = (size_t)h->cases_cap * tiles * 4;
    const size_t budget = kPartialsBudgetBytes / sizeof(double);
    if (part > budget) part = std::max(budget, (size_t)tiles * 4 * h->D);   // at least one whole candidate per launch
    {   // the latency form of the rollout (at most 2 CTAs of 32 trajectories per SM) keeps one partial per CTA
        if (cu(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device), "cudaDeviceGetAttribute")) return bail(KFBO_ECUDA);
        part = std::max(part, (size_t)2 * h->sm_count * 4);
    }
    h->partials_cap = part;
    const size_t cand_bytes = (size_t)p->max_candidates * sizeof(double2);
    const size_t case_bytes = cand_bytes + kfbo::kPeerCtxBytes + h->cases_cap * sizeof(Case);
    const size_t compact_bytes = kfbo::compact_costs_bytes(p->max_candidates, h->D);
    if (cu(cudaMalloc(&h->d_case_buf, case_bytes), "cudaMalloc(cases)") ||
        cu(cudaMallocHost(&h->h_case_buf, case_bytes), "cudaMallocHost(cases)") ||
        cu(cudaMalloc(&h->d_compact, compact_bytes), "cudaMalloc(compact costs)") ||
        cu(cudaMallocHost(&h->h_compact, compact_bytes), "cudaMallocHost(compact costs)") ||
        cu(cudaMalloc(&h->d_partials, h->partials_cap * sizeof(double)), "cudaMalloc(partials)") ||
        cu(cudaMalloc(&h->d_tickets, h->cases_cap * sizeof(unsigned)), "cudaMalloc(tickets)") ||
        cu(cudaMemset(h->d_tickets, 0, h->cases_cap * sizeof(unsigned)), "cudaMemset(tickets)") ||
        cu(cudaMalloc(&h->d_sums, (size_t)h->cases_cap * 4 * sizeof(double)), "cudaMalloc(sums)") ||
        cu(cudaMallocHost(&h->h_sums, (size_t)h->cases_cap * 4 * sizeof(double)), "cudaMallocHost(sums)"))
        return bail(KFBO_ECUDA);
    if (cu(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->h_sums_dev), h->h_sums, 0), "cudaHostGetDevicePointer(sums)")) return bail(KFBO_ECUDA);
    if (cu(cudaMallocHost(&h->h_flag, 64), "cudaMallocHost(flag)") ||
        cu(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->h_flag_dev), h->h_flag, 0), "cudaHostGetDevicePointer(flag)")) return bail(KFBO_ECUDA);
    std::memset(h->h_flag, 0, 64);
    h->d_hdr = h->d_case_buf + cand_bytes;
    h->h_hdr = h->h_case_buf + cand_bytes;
    std::memset(h->h_hdr, 0, kfbo::kPeerCtxBytes);
    h->d_cases = reinterpret_cast<Case *>(h->d_hdr + kfbo::kPeerCtxBytes);
    h->h_cases = reinterpret_cast<Case *>(h->h_hdr + kfbo::kPeerCtxBytes);
    *out = h;
    return KFBO_OK;
}

int kfbo_set_stream_base(kfbo_handle *h, uint64_t base) { if (!h) return KFBO_EINVAL; h->stream_base = base; return KFBO_OK; }
int kfbo_set_eval_counter(kfbo_handle *h, uint64_t e) { if (!h) return KFBO_EINVAL; h->eval_counter = e; return KFBO_OK; }

int kfbo_set_option(kfbo_handle *h, int32_t option, int64_t value)
{
    if (!h) return KFBO_EINVAL;
    switch (option) {
    case KFBO_OPT_SUPPLIED_KERNEL:
        if (value < KFBO_SUPPLIED_AUTO || value > KFBO_SUPPLIED_BULK) return fail(h, KFBO_EINVAL, "kfbo_set_option: supplied kernel %lld", (long long)value);
        h->supplied_kernel = (int)value;
        return KFBO_OK;
    case KFBO_OPT_ROLLOUT_FORM:
        // the reduced (tabulated-gain, error-state) form was removed in round 2: it never beat the reference form by more
        // than 6 % (bound by shared-memory wavefronts of the noise tables, profiles/variants_reduced_r2a.txt)
        if (value != KFBO_FORM_REFERENCE) return fail(h, KFBO_EINVAL, "kfbo_set_option: only KFBO_FORM_REFERENCE exists (the reduced form was removed)");
        return KFBO_OK;
    case KFBO_OPT_BULK_SMEM_PAD:
        if (value < 0 || value > (128 << 10)) return fail(h, KFBO_EINVAL, "kfbo_set_option: smem pad %lld", (long long)value);
        h->bulk_smem_pad = (int)value;
        return KFBO_OK;
    case KFBO_OPT_REDUCE:
        if (value != KFBO_REDUCE_AUTO && value != KFBO_REDUCE_NCCL) return fail(h, KFBO_EINVAL, "kfbo_set_option: reduce %lld", (long long)value);
        h->reduce_pref = (int)value;
        return KFBO_OK;
    case KFBO_OPT_PEER_FUSE:
        if (value < KFBO_PEER_FUSE_NEVER || value > KFBO_PEER_FUSE_ALWAYS) return fail(h, KFBO_EINVAL, "kfbo_set_option: peer fuse %lld", (long long)value);
        h->peer_fuse = (int)value;
        return KFBO_OK;
    case KFBO_OPT_RESERVE_SUPPLIED_TRIALS: {
        // everything kfbo_eval_supplied / kfbo_eval_philox_trials / kfbo_trace_trial would otherwise allocate on first use,
        // for up to `value` trials, max_candidates candidates and the longest sample time
        if (value < 1 || value > 0x7fffffff) return fail(h, KFBO_EINVAL, "kfbo_set_option: reserve %lld trials", (long long)value);
        KFBO_ON_DEVICE(h);
        const size_t B = (size_t)value, T = (size_t)h->max_T, C = (size_t)h->p.max_candidates, NS = (size_t)h->N;
        int rc;
        if ((rc = ensure_scratch(h, 0, NS * B)) || (rc = ensure_scratch(h, 1, NS * T * B)) || (rc = ensure_scratch(h, 2, T * B)) ||
            (rc = ensure_scratch(h, 3, supplied_scratch3_doubles((int)C, h->N > 2 ? kfbo::aug_tiles(B) : kfbo::small_tiles(B)))) ||
            (rc = ensure_scratch(h, 4, std::max(C * 4 * B, (size_t)KFBO_TRACE_COLUMNS * T))))
            return rc;
        h->sup_cases.resize(C);
        if (h->N > 2) h->sup_aug_cases.resize(C);
        h->sup_sums.resize(C * 4);
        return KFBO_OK;
    }
    case KFBO_OPT_CUDA_GRAPH:
        h->use_graph = value != 0;
        if (!h->use_graph && h->gexec) { KFBO_ON_DEVICE(h); cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
        return KFBO_OK;
    case KFBO_OPT_SMALL_LAUNCH:
        h->use_small = value != 0;
        return KFBO_OK;
    case KFBO_OPT_HOST_POLL:
        h->use_poll = value != 0;
        return KFBO_OK;
    case KFBO_OPT_INJECT_FAULT:
        h->inject_fault = value != 0;
        return KFBO_OK;
    case KFBO_OPT_AUG_LAYOUT:
        if (value != KFBO_AUG_LAYOUT_LANES && value != KFBO_AUG_LAYOUT_THREAD) return fail(h, KFBO_EINVAL, "kfbo_set_option: aug layout %lld", (long long)value);
        h->aug_layout = (int)value;
        return KFBO_OK;
    case KFBO_OPT_PEER_TIMEOUT_MS:
        if (value < 1 || value > 3600000) return fail(h, KFBO_EINVAL, "kfbo_set_option: peer timeout %lld ms", (long long)value);
        h->peer_timeout_ms = value;
        return KFBO_OK;
    default:
        return fail(h, KFBO_EINVAL, "kfbo_set_option: unknown option %d", option);
    }
}

int kfbo_set_cuda_stream(kfbo_handle *h, void *cuda_stream)
{
    if (!h) return KFBO_EINVAL;
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return KFBO_OK;
}

// The domain of a candidate (pnoise_est, onoise_est).  Outside it -- non-finite or <= 0 -- the reference still publishes a
// "cost" message: its covariance recursion degenerates (R = 0: P -> 0, det(P) = 0; Qd < 0: P indefinite) and the
// std::abs(std::log(.)) of trial_node.cpp:77,93 yields NaN, which the optimiser accepts as an answer (its wait loop
// ends on anything that is not -1, robot1d_bayesopt.cpp:107).  The same is observable here: such a candidate gets a
// NaN kfbo_result (max_cost NaN, index_max 0 = what max_element returns for all-NaN costs) and the call succeeds.
// Positive values outside [1e-100, 1e100] are refused with KFBO_ERANGE: the reciprocals of the filter step
// (kfbo_device.cuh: KFBO_RCP) assume normal, well-scaled S and det(P).
constexpr double kCandMin = 1e-100, kCandMax = 1e100;
static int classify_candidate(double q, double r)   // 0: in the domain, 1: NaN cost, < 0: error code
{
    if (!(q > 0.0) || !(r > 0.0) || !std::isfinite(q) || !std::isfinite(r)) return 1;
    if (q < kCandMin || q > kCandMax || r < kCandMin || r > kCandMax) return KFBO_ERANGE;
    return 0;
}

static int check_point_candidate(kfbo_handle *h, const char *who, double q, double r)   // the parity / trace entry points
{
    if (classify_candidate(q, r) != 0)
        return fail(h, KFBO_EINVAL, "%s: pnoise_est %g / onoise_est %g outside the domain (finite, in [1e-100, 1e100])", who, q, r);
    return KFBO_OK;
}

// An evaluation in three phases, so that a device group (kfbo_group_*) can run each phase on all its devices before
// the next:  launch (constants, copies, rollout kernels)  ->  reduce (the all-reduce of the cost sums)  ->
// collect (result copy, wait, cost assembly).  kfbo_eval_batch is the three in a row on one handle.
// Whether the cost sums of an evaluation with `nsums` doubles will meet over peer memory (else: one ncclAllReduce).  The same
// answer on every rank / device: it depends on the batch size and on options only.
static bool sums_meet_in_peer_memory(const kfbo_handle *h, size_t nsums)
{
    return h->comm && h->peer_ready && h->reduce_pref != KFBO_REDUCE_NCCL && nsums <= h->xchg_cap &&
           (h->shard_mode == KFBO_SHARD_CANDIDATES || nsums <= kPeerMaxSummedDoubles);
}

static int eval_check(kfbo_handle *h, int32_t C, const double *q_est, const double *r_est, const void *out)
{
    if (h->poisoned)
        return fail(h, KFBO_ESTATE, "an earlier multi-GPU evaluation on this handle failed part-way (%s): the ranks are out of step -- "
                                    "destroy the handles and start over", h->poison_why.c_str());
    if (!q_est || !r_est || !out) return fail(h, KFBO_EINVAL, "kfbo_eval_batch: NULL argument");
    if (C < 1 || C > h->p.max_candidates)
        return fail(h, KFBO_EINVAL, "kfbo_eval_batch: n_candidates %d outside [1, max_candidates=%d]", C, h->p.max_candidates);
    h->n_bad = 0;
    for (int c = 0; c < C; ++c) {
        const int cl = classify_candidate(q_est[c], r_est[c]);
        if (cl < 0)
            return fail(h, cl, "kfbo_eval_batch: candidate %d (pnoise %g, onoise %g) outside [1e-100, 1e100]", c, q_est[c], r_est[c]);
        h->cand_bad[c] = (uint8_t)cl;
        h->n_bad += cl;
    }
    return KFBO_OK;
}

// After a failure between launch and collect of a multi-rank evaluation: wait for whatever was enqueued (the exchange
// kernel gives up after its timeout), clear the exchange's status word and refuse further evaluations.
static void poison(kfbo_handle *h)
{
    if (!h->comm) return;
    DeviceGuard dg(h->device);
    cudaStreamSynchronize(h->stream);
    cudaGetLastError();
    if (h->h_status) *h->h_status = 0;
    if (!h->poisoned) h->poison_why = h->err;
    h->poisoned = true;
}

// Timing of the last evaluation, read back lazily: ev0..ev1 around the rollout, ev1..ev2 around the exchange of the sums.
static void resolve_timing(kfbo_handle *h)
{
    if (!h->timing_pending) return;
    h->timing_pending = false;
    DeviceGuard dg(h->device);
    float ms = 0.f, ar = 0.f;
    cudaEvent_t last = h->comm ? h->ev2 : h->ev1;
    if (cudaEventSynchronize(last) != cudaSuccess || cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) { cudaGetLastError(); return; }
    h->last_ms = ms;
    h->kernel_ms_total += ms;
    ++h->kernel_ms_evals;
    h->last_host_us[4] = 0.0;
    if (h->comm && cudaEventElapsedTime(&ar, h->ev1, h->ev2) == cudaSuccess) h->last_host_us[4] = 1e3 * ar;
    cudaGetLastError();
}

static cudaError_t record_event(kfbo_handle *h, cudaEvent_t ev)
{
    // inside a stream capture an event that is to be read with cudaEventElapsedTime must be recorded as an "external" node
    return cudaEventRecordWithFlags(ev, h->stream, h->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}

// bad: the classification of the candidates by eval_check (of this handle, or of a group's first handle)
// compact: the evaluation ends in the device-side cost assembly (kfbo_eval_batch_costs) instead of the host's
static int eval_launch(kfbo_handle *h, int32_t C, const double *q_est, const double *r_est, const uint8_t *bad, bool compact = false)
{
    using clk = std::chrono::steady_clock;
    KFBO_ON_DEVICE(h);
    if (h->inject_fault) {      // a rank / device that drops out of a multi-GPU evaluation, for the tests of what the others do then
        h->inject_fault = false;
        return fail(h, KFBO_ECUDA, "injected fault (KFBO_OPT_INJECT_FAULT): nothing was launched on device %d", h->device);
    }
    const int D = h->D;
    int64_t c0 = 0, c1 = C, t0 = 0, t1 = h->b_eff;
    if (h->comm) {
        if (h->shard_mode == KFBO_SHARD_CANDIDATES) kfbo::host::shard_range(C, h->rank, h->nranks, &c0, &c1);
        else kfbo::host::shard_range(h->b_eff, h->rank, h->nranks, &t0, &t1);
    }
    const int ncand = (int)(c1 - c0), ncases = ncand * D;
    h->compact = compact;
    if (h->timing_accumulate) resolve_timing(h);   // the previous evaluation's events are about to be recorded again
    h->timing_pending = false;
    h->tp[0] = clk::now();
    // The per-(candidate, sample time) constants.  One candidate (the "noise" -> "cost" call): built here, they also ride
    // in the kernel's parameter bank.  A batch: 16 bytes per candidate go to the device and build_cases_kernel makes
    // the same constants there, bit for bit (kfbo_host.h: RoundedOps) -- no per-candidate host work, a sixth of the bytes.
    const bool aug = h->N > 2;             // state-augmented variant: its own case type, built here for any batch size
    const bool device_build = C > 1 && !aug;
    const uint64_t stream0 = h->stream_base + h->eval_counter * (uint64_t)h->p.max_candidates * (uint64_t)D;
    double2 *const h_cand = reinterpret_cast<double2 *>(h->h_hdr) - ncand;      // right-aligned against the header
    if (device_build) {
        // a candidate outside the domain still occupies its slots (with any valid constants); its costs become NaN
        for (int64_t c = c0; c < c1; ++c) h_cand[c - c0] = bad[c] ? make_double2(1.0, 1.0) : make_double2(q_est[c], r_est[c]);
    } else {
        for (int64_t c = c0; c < c1; ++c)
            for (int k = 0; k < D; ++k) {
                const double q = bad[c] ? 1.0 : q_est[c], r = bad[c] ? 1.0 : r_est[c];
                const int rc = aug ? build_aug_case(h, k, q, r, stream0 + (uint64_t)c * D + k, (int)(c * D + k), &h->h_aug_cases[(c - c0) * D + k])
                                   : build_case(h, k, q, r, stream0 + (uint64_t)c * D + k, (int)(c * D + k), &h->h_cases[(c - c0) * D + k]);
                if (rc) return rc;
            }
    }
    const size_t nsums = (size_t)C * D * 4;
    // ---- how the sums of the ranks will meet (the same decision on every rank: it depends on C, D and options only)
    const bool gather = h->shard_mode == KFBO_SHARD_CANDIDATES;
    const bool use_peer = sums_meet_in_peer_memory(h, nsums);
    const uint32_t ntrials = (uint32_t)(t1 - t0);
    const int tiles = aug ? kfbo::aug_tiles(ntrials) : kfbo::rollout_tiles(ntrials);
    int step = cases_per_launch(h, std::max(tiles, 1));
    const bool has_work = ncases > 0 && ntrials > 0;
    // Opt-in (KFBO_OPT_PEER_FUSE): the exchange rides in the tail of the rollout kernel when the evaluation is one launch.
    // Off by default because it is slower, measured: the rollout loop is FP64-issue bound and ptxas schedules it
    // differently when the exchange follows it in the same kernel (+0.9 % kernel time for one candidate, +2.0 % for a
    // batch, i.e. 15-60 us), which is more than the 2-3 us a second launch on the same stream costs.
    const bool fused = use_peer && has_work && ncases <= step && h->peer_fuse == KFBO_PEER_FUSE_ALWAYS && !aug;
    h->reduce_path = !h->comm ? 0 : !use_peer ? 1 : fused ? 3 : 2;
    // Where the total lands.  A small result is stored by the kernel straight into the mapped pinned buffer the costs are
    // assembled from (no device-to-host copy on the critical path of a "noise" -> "cost" evaluation): by the rollout
    // itself on a single rank, by the exchange when the trials are sharded.  Otherwise it stays on the device (the
    // all-reduce buffer d_sums, or the gathered block of the exchange buffer) and one copy brings it over -- or, compact,
    // assemble_costs_kernel reads it there.
    const bool zero_copy = !compact && nsums <= kZeroCopyMaxDoubles && (!h->comm || (use_peer && !gather));
    double *const rollout_sums = (!h->comm && zero_copy) ? h->h_sums_dev : h->d_sums;
    h->result_src = zero_copy ? nullptr : h->d_sums;
    if (use_peer) {
        kfbo::PeerCtx *pc = reinterpret_cast<kfbo::PeerCtx *>(h->h_hdr);
        const uint32_t epoch = ++h->peer_epoch;
        const size_t slabs = gather ? 1u : (size_t)h->nranks;
        for (int r = 0; r < h->nranks; ++r) {
            pc->flags[r] = reinterpret_cast<unsigned *>(h->peer_base[r]);
            pc->vals[r] = reinterpret_cast<double *>(h->peer_base[r] + kXchgFlagBytes);
            int64_t b = 0, e = C;
            if (gather) kfbo::host::shard_range(C, r, h->nranks, &b, &e);
            pc->begin[r] = (int32_t)(b * D * 4); pc->end[r] = (int32_t)(e * D * 4);
        }
        pc->nranks = h->nranks; pc->rank = h->rank; pc->gather = gather ? 1 : 0; pc->nsums = (int32_t)nsums;
        pc->cap = (uint32_t)h->xchg_cap; pc->epoch = epoch; pc->ncases = (uint32_t)ncases;
        pc->done = h->d_done; pc->status = h->h_status; pc->sums = h->d_sums;
        pc->out = zero_copy ? h->h_sums_dev : h->d_sums;
        // a total that lands in mapped host memory is announced there too: the host polls for it (eval_collect)
        h->flag_armed = zero_copy && !fused && h->use_poll;
        pc->host_flag = h->flag_armed ? h->h_flag_dev : nullptr;
        h->flag_expect = epoch;
        pc->timeout_ns = (unsigned long long)h->peer_timeout_ms * 1000000ull;
        if (gather) h->result_src = reinterpret_cast<const double *>(h->d_xchg + kXchgFlagBytes) + (size_t)(epoch & 1u) * slabs * h->xchg_cap;
    }
    if (!use_peer) h->flag_armed = false;
    h->tp[1] = clk::now();
    // ---- the single-candidate evaluation as ONE graph launch (see kfbo_handle::gexec)
    h->graph_used = false;
    // Measured (profiles/NOTES_r2.md): where the evaluation is a copy and ONE kernel (a single rank) the replay costs more host
    // time than the two calls it replaces (10.8 against 8.5 us); with the exchange kernel and its event behind the rollout it
    // is ahead (13.4 against 16.4 us) -- so: multi-GPU evaluations whose sums meet over peer memory.
    const bool graphable = h->use_graph && C == 1 && !aug && has_work && ncases <= step && !fused && use_peer;
    if (graphable) {
        kfbo_handle::GraphSig sig;
        sig.stream = h->stream; sig.sums = rollout_sums; sig.ncases = ncases; sig.ntrials = ntrials; sig.t0 = (uint32_t)t0; sig.peer = use_peer; sig.zero_copy = zero_copy;
        if (h->gexec && sig == h->gsig) {
            h->grec.st = h->st;
            h->grec.st.k0 = 0; h->grec.st.nk = D;
            kfbo::set_uniform_estimator(&h->grec.st, h->h_cases);
            cudaKernelNodeParams kp{};
            kp.func = const_cast<void *>(h->grec.func); kp.gridDim = h->grec.grid; kp.blockDim = dim3(h->grec.block);
            kp.sharedMemBytes = (unsigned)h->grec.smem; kp.kernelParams = h->grec.argv; kp.extra = nullptr;
            if (cudaGraphExecKernelNodeSetParams(h->gexec, h->gnode, &kp) == cudaSuccess && cudaGraphLaunch(h->gexec, h->stream) == cudaSuccess) {
                h->last_h2d = (int64_t)((size_t)ncases * sizeof(Case) + (use_peer ? kfbo::kPeerCtxBytes : 0));
                h->last_launches = use_peer ? 2 : 1;
                h->graph_used = true;
                return KFBO_OK;
            }
            cudaGetLastError();                 // the replay failed: drop the graph, enqueue directly from now on
            cudaGraphExecDestroy(h->gexec);
            h->gexec = nullptr;
            h->use_graph = false;
        } else {
            if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
            if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                h->capturing = true;
                h->gsig = sig;
            } else {
                cudaGetLastError();
                h->use_graph = false;
            }
        }
    }
    // whatever happens below, a capture that was begun is ended (a failed enqueue must not leave the stream capturing)
    struct CaptureEnd {
        kfbo_handle *h;
        ~CaptureEnd()
        {
            if (!h->capturing) return;
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(h->stream, &g);
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            h->capturing = false;
            h->use_graph = false;
        }
    } capture_end{h};
    // Zeros only where the rollout does not write every slot that is read afterwards: a rank without work, and the slots
    // of other ranks' candidates under the NCCL all-reduce (the exchange moves only what a rank owns).
    if (!has_work || (h->comm && !use_peer && gather))
        KFBO_CUDA(h, cudaMemsetAsync(h->d_sums, 0, nsums * sizeof(double), h->stream));
    // ONE host-to-device copy: [candidates][header] or [header][cases], the header only when the exchange needs it
    h->last_h2d = 0;
    if (device_build) {
        const size_t bytes = (size_t)ncand * sizeof(double2) + (use_peer ? kfbo::kPeerCtxBytes : 0);
        const size_t off = (size_t)ncand * sizeof(double2);
        if (bytes) KFBO_CUDA(h, cudaMemcpyAsync(h->d_hdr - off, h->h_hdr - off, bytes, cudaMemcpyHostToDevice, h->stream));
        h->last_h2d = (int64_t)bytes;
        if (ncand > 0) {
            kfbo::CaseBuildConsts cb{};
            for (int k = 0; k < D; ++k) {
                cb.dt[k] = h->p.dt[k]; cb.T[k] = h->p.max_iterations[k];
                cb.l00[k] = h->truth_l[k][0]; cb.l10[k] = h->truth_l[k][1]; cb.l11[k] = h->truth_l[k][2]; cb.sr[k] = h->truth_sr[k];
            }
            cb.D = D; cb.c0 = (int32_t)c0; cb.stream0 = stream0;
            KFBO_CUDA(h, kfbo::launch_build_cases(reinterpret_cast<const double2 *>(h->d_hdr) - ncand, ncand, cb, h->d_cases, h->stream));
        }
    } else if (aug) {
        if (use_peer) KFBO_CUDA(h, cudaMemcpyAsync(h->d_hdr, h->h_hdr, kfbo::kPeerCtxBytes, cudaMemcpyHostToDevice, h->stream));
        if (ncases > 0)
            KFBO_CUDA(h, cudaMemcpyAsync(h->d_aug_cases, h->h_aug_cases, (size_t)ncases * sizeof(kfbo::AugCase), cudaMemcpyHostToDevice, h->stream));
        h->last_h2d = (int64_t)((size_t)ncases * sizeof(kfbo::AugCase) + (use_peer ? kfbo::kPeerCtxBytes : 0));
    } else {
        const size_t bytes = (size_t)ncases * sizeof(Case) + (use_peer ? kfbo::kPeerCtxBytes : 0);
        if (use_peer) KFBO_CUDA(h, cudaMemcpyAsync(h->d_hdr, h->h_hdr, bytes, cudaMemcpyHostToDevice, h->stream));
        else if (ncases > 0) KFBO_CUDA(h, cudaMemcpyAsync(h->d_cases, h->h_cases, bytes, cudaMemcpyHostToDevice, h->stream));
        h->last_h2d = (int64_t)bytes;
    }
    h->last_launches = (device_build && ncand > 0) ? 1 : 0;
    // the latency form of the rollout for launches that cannot fill the GPU: its 32-trajectory tiles need four times the partials
    const int small_max = (h->use_small && !aug && ncases <= step && (size_t)ncases * kfbo::small_tiles(ntrials) * 4 <= h->partials_cap)
                              ? 2 * h->sm_count : 0;
    KFBO_CUDA(h, record_event(h, h->ev0));
    if (has_work) {
        for (int first = 0; first < ncases; first += step) {
            const int n = std::min(step, ncases - first);
            kfbo::SampleTimeConsts st = h->st;
            st.k0 = 0; st.nk = D;   // cases are laid out [candidate][sample time]; chunks hold whole candidates
            if (aug) {
                KFBO_CUDA(h, kfbo::launch_rollout_aug(h->N, false, h->d_aug_cases + first, n, D, 0, h->d_aug_st, h->d_u, h->d_logtab, h->p.seed, (uint32_t)t0,
                                                      ntrials, nullptr, nullptr, nullptr, 0, nullptr, 0, h->d_partials, h->d_tickets, rollout_sums, h->stream,
                                                      h->h_aug_st, h->aug_layout));
                h->last_launches += h->aug_layout == KFBO_AUG_LAYOUT_THREAD ? D : 1;   // the per-thread layout launches once per sample time
            } else {
                KFBO_CUDA(h, kfbo::launch_rollout_philox(h->d_cases + first, n, st, h->d_gu, h->d_logtab, h->min_T, h->p.seed, (uint32_t)t0, ntrials,
                                                         nullptr, 0, h->d_partials, h->d_tickets, rollout_sums, h->stream,
                                                         device_build ? nullptr : h->h_cases + first, fused, h->capturing ? &h->grec : nullptr, small_max));
                ++h->last_launches;
            }
        }
    }
    KFBO_CUDA(h, record_event(h, h->ev1));
    if (use_peer && !fused) {
        const kfbo::PeerCtx *pc = reinterpret_cast<const kfbo::PeerCtx *>(h->h_hdr);
        KFBO_CUDA(h, kfbo::launch_peer_exchange(reinterpret_cast<const kfbo::PeerCtx *>(h->d_hdr), pc->end[h->rank] - pc->begin[h->rank], h->stream));
        ++h->last_launches;
    }
    if (h->capturing) {
        // close the capture (the exchange's closing event belongs to it), instantiate, find the rollout's node, launch
        if (use_peer) KFBO_CUDA(h, record_event(h, h->ev2));
        cudaGraph_t g = nullptr;
        h->capturing = false;
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        if (e == cudaSuccess) e = cudaGraphInstantiate(&h->gexec, g, 0);
        if (e == cudaSuccess) {
            size_t n = 0;
            cudaGraphGetNodes(g, nullptr, &n);
            std::vector<cudaGraphNode_t> nodes(n);
            cudaGraphGetNodes(g, nodes.data(), &n);
            h->gnode = nullptr;
            for (cudaGraphNode_t nd : nodes) {
                cudaGraphNodeType ty;
                cudaKernelNodeParams kp{};
                if (cudaGraphNodeGetType(nd, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel &&
                    cudaGraphKernelNodeGetParams(nd, &kp) == cudaSuccess && kp.func == h->grec.func) h->gnode = nd;
            }
            if (!h->gnode) e = cudaErrorUnknown;
        }
        if (g) cudaGraphDestroy(g);       // the executable graph keeps what it needs (node handles stay valid for ExecKernelNodeSetParams)
        if (e == cudaSuccess) e = cudaGraphLaunch(h->gexec, h->stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
            h->use_graph = false;
            return fail(h, KFBO_ECUDA, "building the evaluation's CUDA graph failed: %s", cudaGetErrorString(e));
        }
        h->graph_used = true;
    }
    return KFBO_OK;
}

// in_group: the call sits between ncclGroupStart / ncclGroupEnd of a device group, where NCCL defers the operation to
// the end of the group -- the event that closes the all-reduce is then recorded by the group, after ncclGroupEnd.
static int eval_reduce(kfbo_handle *h, int32_t C, bool in_group)
{
    if (!h->comm) return KFBO_OK;
    KFBO_ON_DEVICE(h);
    if (h->reduce_path >= 2) {   // the sums met in peer memory: in the rollout kernel itself, or in the kernel behind it
        if (!in_group && !h->graph_used) KFBO_CUDA(h, cudaEventRecord(h->ev2, h->stream));   // a graph replay records it itself
        return KFBO_OK;
    }
    const size_t nsums = (size_t)C * h->D * 4;
    KFBO_NCCL(h, nccl().AllReduce(h->d_sums, h->d_sums, nsums, ncclDouble, ncclSum, h->comm, h->stream));
    if (!in_group) KFBO_CUDA(h, cudaEventRecord(h->ev2, h->stream));
    return KFBO_OK;
}

// out == NULL: wait for the device only (a group collects the results on its first device; after the all-reduce
// every device holds the same sums).  A compact evaluation (h->compact) hands back cost / max_cost / index_max instead.
static int eval_collect(kfbo_handle *h, int32_t C, kfbo_result *out, double *cost = nullptr, double *max_cost = nullptr, int32_t *index_max = nullptr)
{
    using clk = std::chrono::steady_clock;
    KFBO_ON_DEVICE(h);
    const int D = h->D;
    const size_t nsums = (size_t)C * D * 4;
    h->last_d2h = 0;
    if (h->compact && !cost) {
        // a device of a group that is not the collecting one: nothing to bring over
    } else if (h->compact) {
        // trial_node.cpp:70-119 for every candidate on the device, then C x (D + 1) doubles + C ints instead of the sums
        KFBO_CUDA(h, kfbo::launch_assemble_costs(h->result_src, C, &h->p, h->d_compact, h->stream));
        ++h->last_launches;
        const size_t bytes = kfbo::compact_costs_bytes(C, D);
        KFBO_CUDA(h, cudaMemcpyAsync(h->h_compact, h->d_compact, bytes, cudaMemcpyDeviceToHost, h->stream));
        h->last_d2h = (int64_t)bytes;
    } else if (out && h->result_src) {   // NULL: the kernel has stored the result in h_sums itself
        KFBO_CUDA(h, cudaMemcpyAsync(h->h_sums, h->result_src, nsums * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        h->last_d2h = (int64_t)(nsums * sizeof(double));
    } else if (out) {
        h->last_d2h = (int64_t)(nsums * sizeof(double));   // stored over PCIe by the kernel (mapped pinned memory)
    }
    const clk::time_point w2 = clk::now();
    if (h->flag_armed && !h->compact && !(out && h->result_src)) {
        // the exchange kernel stores the total into mapped host memory and then this word: poll for it (what a stream
        // synchronisation does inside the driver, minus the driver); every so often make sure the stream is still alive
        volatile unsigned *flag = h->h_flag;
        for (unsigned spins = 1; (int)(*flag - h->flag_expect) < 0; ++spins)
            if ((spins & 0x3FFFu) == 0 && cudaStreamQuery(h->stream) != cudaErrorNotReady) break;
        std::atomic_thread_fence(std::memory_order_acquire);
        if ((int)(*flag - h->flag_expect) < 0) KFBO_CUDA(h, cudaStreamSynchronize(h->stream));   // the stream ended without the flag: an error
    } else {
        KFBO_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    const clk::time_point w3 = clk::now();
    if (h->reduce_path >= 2 && *h->h_status != 0) {
        const int who = *h->h_status - 1;
        *h->h_status = 0;
        return fail(h, KFBO_EPEER, "peer-memory exchange: rank %d waited %lld ms for the cost sums of rank %d (did every rank make this call?)",
                    h->rank, (long long)h->peer_timeout_ms, who);
    }
    h->timing_pending = true;            // read back on query (kfbo_last_kernel_ms, kfbo_last_host_us) or, accumulating, at the next launch
    h->last_ncand = (out && !h->compact) ? C : 0;
    if (h->compact && cost) {
        const double *hc = static_cast<const double *>(h->h_compact), *hm = hc + (size_t)C * D;
        const int32_t *hi = reinterpret_cast<const int32_t *>(hm + C);
        std::memcpy(cost, hc, (size_t)C * D * sizeof(double));
        std::memcpy(max_cost, hm, (size_t)C * sizeof(double));
        std::memcpy(index_max, hi, (size_t)C * sizeof(int32_t));
        if (h->n_bad)
            for (int c = 0; c < C; ++c)
                if (h->cand_bad[c]) {
                    for (int k = 0; k < D; ++k) cost[(size_t)c * D + k] = std::numeric_limits<double>::quiet_NaN();
                    max_cost[c] = std::numeric_limits<double>::quiet_NaN();
                    index_max[c] = 0;
                }
    } else if (out) {
        finalize_all(h, C, out);
    }
    const clk::time_point w4 = clk::now();
    const clk::time_point w[5] = {h->tp[0], h->tp[1], w2, w3, w4};
    for (int i = 0; i < 4; ++i) h->last_host_us[i] = std::chrono::duration<double, std::micro>(w[i + 1] - w[i]).count();
    ++h->eval_counter;
    return KFBO_OK;
}

int kfbo_eval_batch(kfbo_handle *h, int32_t C, const double *q_est, const double *r_est, kfbo_result *out)
{
    if (!h) return KFBO_EINVAL;
    int rc = eval_check(h, C, q_est, r_est, out);
    if (rc) return rc;
    rc = eval_launch(h, C, q_est, r_est, h->cand_bad.data());
    if (rc == KFBO_OK) rc = eval_reduce(h, C, false);
    if (rc == KFBO_OK) rc = eval_collect(h, C, out);
    if (rc) poison(h);
    return rc;
}

int kfbo_eval(kfbo_handle *h, const double *pn, int32_t npn, const double *on, int32_t non, kfbo_result *out)
{
    if (!h) return KFBO_EINVAL;
    if (!pn || !on || !out || npn < 1 || non < 1)
        return fail(h, KFBO_EINVAL, "kfbo_eval: needs pn[>=1], on[>=1] (system_models.hpp:39,143 read element 0)");
    return kfbo_eval_batch(h, 1, pn, on, out);
}

int kfbo_eval_batch_costs(kfbo_handle *h, int32_t C, const double *q_est, const double *r_est, double *cost, double *max_cost, int32_t *index_max)
{
    if (!h) return KFBO_EINVAL;
    if (!cost || !max_cost || !index_max) return fail(h, KFBO_EINVAL, "kfbo_eval_batch_costs: NULL argument");
    int rc = eval_check(h, C, q_est, r_est, cost);
    if (rc) return rc;
    rc = eval_launch(h, C, q_est, r_est, h->cand_bad.data(), true);
    if (rc == KFBO_OK) rc = eval_reduce(h, C, false);
    if (rc == KFBO_OK) rc = eval_collect(h, C, nullptr, cost, max_cost, index_max);
    if (rc) poison(h);
    return rc;
}

int kfbo_last_transfer_bytes(const kfbo_handle *h, int64_t *h2d, int64_t *d2h)
{
    if (!h) return KFBO_EINVAL;
    if (h2d) *h2d = h->last_h2d;
    if (d2h) *d2h = h->last_d2h;
    return KFBO_OK;
}

int kfbo_last_sums(const kfbo_handle *h, double *sums, int32_t n_candidates)
{
    if (!h || !sums || n_candidates < 1 || n_candidates > h->last_ncand) return KFBO_EINVAL;
    std::memcpy(sums, h->h_sums, (size_t)n_candidates * h->D * 4 * sizeof(double));
    if (h->n_bad)     // a candidate outside the domain has no sums (its slots were run with placeholder constants)
        for (int c = 0; c < n_candidates; ++c)
            if (h->cand_bad[c])
                for (int i = 0; i < h->D * 4; ++i) sums[(size_t)c * h->D * 4 + i] = std::numeric_limits<double>::quiet_NaN();
    return KFBO_OK;
}

int kfbo_last_host_us(const kfbo_handle *hc, double us[KFBO_HOST_PHASES])
{
    if (!hc || !us) return KFBO_EINVAL;
    kfbo_handle *h = const_cast<kfbo_handle *>(hc);
    resolve_timing(h);
    for (int i = 0; i < KFBO_HOST_PHASES; ++i) us[i] = h->last_host_us[i];
    return KFBO_OK;
}

int kfbo_kernel_ms_total(kfbo_handle *h, double *ms_total, int64_t *evaluations, int32_t reset)
{
    if (!h) return KFBO_EINVAL;
    resolve_timing(h);
    if (ms_total) *ms_total = h->kernel_ms_total;
    if (evaluations) *evaluations = h->kernel_ms_evals;
    if (reset) { h->kernel_ms_total = 0.0; h->kernel_ms_evals = 0; }
    h->timing_accumulate = true;         // from now on every evaluation's time is read back (at the next launch at the latest)
    return KFBO_OK;
}

int kfbo_last_kernel_ms(const kfbo_handle *hc, double *ms, int32_t *launches)
{
    if (!hc) return KFBO_EINVAL;
    kfbo_handle *h = const_cast<kfbo_handle *>(hc);
    resolve_timing(h);
    if (ms) *ms = h->last_ms;
    if (launches) *launches = h->last_launches;
    return KFBO_OK;
}

static int supplied_impl(kfbo_handle *h, int32_t k, int32_t C, const double *q_est, const double *r_est, int64_t B,
                         const double *d_x0, const double *d_w, const double *d_v, double *d_per_trial,
                         double *h_per_trial, double *sums)
{
    const bool aug = h->N > 2;
    const int tiles = aug ? kfbo::aug_tiles((size_t)B) : kfbo::rollout_tiles((size_t)B);
    if (C > 65535) return fail(h, KFBO_ERANGE, "kfbo_eval_supplied: at most 65535 candidates per call");
    if (h->sup_cases.size() < (size_t)C) h->sup_cases.resize((size_t)C);   // grow-only (pre-sized by KFBO_OPT_RESERVE_SUPPLIED_TRIALS)
    if (aug && h->sup_aug_cases.size() < (size_t)C) h->sup_aug_cases.resize((size_t)C);
    if (h->sup_sums.size() < (size_t)C * 4) h->sup_sums.resize((size_t)C * 4);
    for (int c = 0; c < C; ++c) {
        const int rc = aug ? build_aug_case(h, k, q_est[c], r_est[c], 0u, c, &h->sup_aug_cases[c]) : build_case(h, k, q_est[c], r_est[c], 0u, c, &h->sup_cases[c]);
        if (rc) return rc;
    }
    int rc = ensure_scratch(h, 3, supplied_scratch3_doubles(C, tiles));
    if (rc) return rc;
    double *d_partials = h->d_scratch[3];
    double *d_sums = d_partials + (size_t)C * tiles * 4;
    void *d_cases = d_sums + (size_t)C * 4;
    unsigned *d_tickets = reinterpret_cast<unsigned *>(reinterpret_cast<double *>(d_cases) + (size_t)C * kCaseDoubles);
    if (aug) KFBO_CUDA(h, cudaMemcpyAsync(d_cases, h->sup_aug_cases.data(), (size_t)C * sizeof(kfbo::AugCase), cudaMemcpyHostToDevice, h->stream));
    else     KFBO_CUDA(h, cudaMemcpyAsync(d_cases, h->sup_cases.data(), (size_t)C * sizeof(Case), cudaMemcpyHostToDevice, h->stream));
    KFBO_CUDA(h, cudaMemsetAsync(d_tickets, 0, (size_t)C * sizeof(unsigned), h->stream));
    KFBO_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (aug) {
        KFBO_CUDA(h, kfbo::launch_rollout_aug(h->N, true, static_cast<const kfbo::AugCase *>(d_cases), C, 1, k, h->d_aug_st, h->d_u, h->d_logtab, h->p.seed,
                                              0u, (uint32_t)B, d_x0, d_w, d_v, (size_t)B, d_per_trial, (size_t)B, d_partials, d_tickets, d_sums, h->stream));
    } else {
        kfbo::SampleTimeConsts st = h->st;
        st.k0 = k; st.nk = 1;
        KFBO_CUDA(h, kfbo::launch_rollout_supplied(static_cast<const Case *>(d_cases), C, st, h->d_gu, h->p.max_iterations[k], d_x0, d_w, d_v, (size_t)B, d_per_trial,
                                                   d_partials, d_tickets, d_sums, h->supplied_kernel, h->bulk_smem_pad, h->stream));
    }
    KFBO_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    h->last_launches = 1;
    std::vector<double> &hs = h->sup_sums;
    KFBO_CUDA(h, cudaMemcpyAsync(hs.data(), d_sums, (size_t)C * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (h_per_trial)
        KFBO_CUDA(h, cudaMemcpyAsync(h_per_trial, d_per_trial, (size_t)C * 4 * B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    KFBO_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    KFBO_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    h->timing_pending = false;
    if (sums) std::memcpy(sums, hs.data(), (size_t)C * 4 * sizeof(double));
    return KFBO_OK;
}

static int supplied_check(kfbo_handle *h, int32_t k, int32_t C, const double *q, const double *r, int64_t B,
                          const void *x0, const void *w, const void *v)
{
    if (!h) return KFBO_EINVAL;
    if (!q || !r || !x0 || !w || !v) return fail(h, KFBO_EINVAL, "kfbo_eval_supplied: NULL argument");
    if (k < 0 || k >= h->D) return fail(h, KFBO_EINVAL, "kfbo_eval_supplied: sample time %d outside [0,%d)", k, h->D);
    if (C < 1 || B < 1 || B > 0x7fffffff) return fail(h, KFBO_EINVAL, "kfbo_eval_supplied: bad n_candidates/B");
    for (int c = 0; c < C; ++c) {
        const int rc = check_point_candidate(h, "kfbo_eval_supplied", q[c], r[c]);
        if (rc) return rc;
    }
    return KFBO_OK;
}

int kfbo_eval_supplied(kfbo_handle *h, int32_t k, int32_t C, const double *q_est, const double *r_est, int64_t B,
                       const double *x0noise, const double *w, const double *v, double *per_trial, double *sums)
{
    int rc = supplied_check(h, k, C, q_est, r_est, B, x0noise, w, v);
    if (rc) return rc;
    KFBO_ON_DEVICE(h);
    const size_t T = (size_t)h->p.max_iterations[k], NS = (size_t)h->N;      // noise rows per step = the state dimension
    if ((rc = ensure_scratch(h, 0, NS * (size_t)B))) return rc;
    if ((rc = ensure_scratch(h, 1, NS * T * (size_t)B))) return rc;
    if ((rc = ensure_scratch(h, 2, T * (size_t)B))) return rc;
    if (per_trial && (rc = ensure_scratch(h, 4, (size_t)C * 4 * (size_t)B))) return rc;
    KFBO_CUDA(h, cudaMemcpyAsync(h->d_scratch[0], x0noise, NS * (size_t)B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    KFBO_CUDA(h, cudaMemcpyAsync(h->d_scratch[1], w, NS * T * (size_t)B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    KFBO_CUDA(h, cudaMemcpyAsync(h->d_scratch[2], v, T * (size_t)B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return supplied_impl(h, k, C, q_est, r_est, B, h->d_scratch[0], h->d_scratch[1], h->d_scratch[2],
                         per_trial ? h->d_scratch[4] : nullptr, per_trial, sums);
}

int kfbo_eval_supplied_dev(kfbo_handle *h, int32_t k, int32_t C, const double *q_est, const double *r_est, int64_t B,
                           const double *d_x0noise, const double *d_w, const double *d_v, double *d_per_trial, double *sums)
{
    int rc = supplied_check(h, k, C, q_est, r_est, B, d_x0noise, d_w, d_v);
    if (rc) return rc;
    KFBO_ON_DEVICE(h);
    return supplied_impl(h, k, C, q_est, r_est, B, d_x0noise, d_w, d_v, d_per_trial, nullptr, sums);
}

int kfbo_eval_philox_trials(kfbo_handle *h, int32_t k, double q_est, double r_est, uint64_t stream, uint32_t trial0,
                            int64_t ntrials, double *per_trial)
{
    if (!h) return KFBO_EINVAL;
    if (!per_trial || k < 0 || k >= h->D || ntrials < 1 || ntrials > 0x7fffffff)
        return fail(h, KFBO_EINVAL, "kfbo_eval_philox_trials: bad argument");
    int rc = check_point_candidate(h, "kfbo_eval_philox_trials", q_est, r_est);
    if (rc) return rc;
    KFBO_ON_DEVICE(h);
    const bool aug = h->N > 2;
    Case c;
    kfbo::AugCase ca;
    rc = aug ? build_aug_case(h, k, q_est, r_est, stream, 0, &ca) : build_case(h, k, q_est, r_est, stream, 0, &c);
    if (rc) return rc;
    const int tiles = aug ? kfbo::aug_tiles((size_t)ntrials) : kfbo::small_tiles((size_t)ntrials);   // room for the smaller tiles of the latency form
    if ((rc = ensure_scratch(h, 3, supplied_scratch3_doubles(1, tiles)))) return rc;
    if ((rc = ensure_scratch(h, 4, 4 * (size_t)ntrials))) return rc;
    double *d_partials = h->d_scratch[3];
    double *d_sums = d_partials + (size_t)tiles * 4;
    void *d_case = d_sums + 4;
    unsigned *d_ticket = reinterpret_cast<unsigned *>(reinterpret_cast<double *>(d_case) + kCaseDoubles);
    if (aug) KFBO_CUDA(h, cudaMemcpyAsync(d_case, &ca, sizeof ca, cudaMemcpyHostToDevice, h->stream));
    else     KFBO_CUDA(h, cudaMemcpyAsync(d_case, &c, sizeof c, cudaMemcpyHostToDevice, h->stream));
    KFBO_CUDA(h, cudaMemsetAsync(d_ticket, 0, sizeof(unsigned), h->stream));
    if (aug) {
        KFBO_CUDA(h, kfbo::launch_rollout_aug(h->N, false, static_cast<const kfbo::AugCase *>(d_case), 1, 1, k, h->d_aug_st, h->d_u, h->d_logtab, h->p.seed,
                                              trial0, (uint32_t)ntrials, nullptr, nullptr, nullptr, 0, h->d_scratch[4], (size_t)ntrials, d_partials, d_ticket,
                                              d_sums, h->stream, h->h_aug_st, h->aug_layout));
    } else {
        kfbo::SampleTimeConsts st = h->st;
        st.k0 = k; st.nk = 1;
        KFBO_CUDA(h, kfbo::launch_rollout_philox(static_cast<const Case *>(d_case), 1, st, h->d_gu, h->d_logtab, h->min_T, h->p.seed, trial0, (uint32_t)ntrials,
                                                 h->d_scratch[4], (size_t)ntrials, d_partials, d_ticket, d_sums, h->stream, nullptr, false, nullptr,
                                                 h->use_small ? 2 * h->sm_count : 0));
    }
    KFBO_CUDA(h, cudaMemcpyAsync(per_trial, h->d_scratch[4], 4 * (size_t)ntrials * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    KFBO_CUDA(h, cudaStreamSynchronize(h->stream));
    h->last_launches = 1;
    return KFBO_OK;
}

int kfbo_trace_trial(kfbo_handle *h, int32_t k, double q_est, double r_est, uint64_t stream, uint32_t trial, double *trace)
{
    if (!h) return KFBO_EINVAL;
    if (!trace || k < 0 || k >= h->D) return fail(h, KFBO_EINVAL, "kfbo_trace_trial: bad argument");
    if (h->N > 2) return fail(h, KFBO_EINVAL, "kfbo_trace_trial: the single-trial trace exists for state_dim = 2 only (trial.cpp:69-78 plots two states)");
    int rc = check_point_candidate(h, "kfbo_trace_trial", q_est, r_est);
    if (rc) return rc;
    KFBO_ON_DEVICE(h);
    Case c;
    rc = build_case(h, k, q_est, r_est, stream, 0, &c);
    if (rc) return rc;
    const size_t T = (size_t)h->p.max_iterations[k];
    if ((rc = ensure_scratch(h, 3, kCaseDoubles + 1))) return rc;
    if ((rc = ensure_scratch(h, 4, 9 * T))) return rc;
    Case *d_case = reinterpret_cast<Case *>(h->d_scratch[3]);
    KFBO_CUDA(h, cudaMemcpyAsync(d_case, &c, sizeof c, cudaMemcpyHostToDevice, h->stream));
    kfbo::SampleTimeConsts st = h->st;
    st.k0 = k; st.nk = 1;
    KFBO_CUDA(h, kfbo::launch_trace_philox(d_case, st, h->d_gu, h->d_logtab, h->p.seed, trial, h->p.dt[k], h->d_scratch[4], h->stream));
    KFBO_CUDA(h, cudaMemcpyAsync(trace, h->d_scratch[4], 9 * T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    KFBO_CUDA(h, cudaStreamSynchronize(h->stream));
    h->last_launches = 1;
    return KFBO_OK;
}

int kfbo_nccl_unique_id(void *id_bytes)
{
    static_assert(sizeof(ncclUniqueId) == KFBO_NCCL_ID_BYTES, "ncclUniqueId size");
    if (!id_bytes) return KFBO_EINVAL;
    if (!nccl().ok) return fail(nullptr, KFBO_ENCCL, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, KFBO_ENCCL, "ncclGetUniqueId: %s", nccl().GetErrorString(r));
    std::memcpy(id_bytes, &id, sizeof id);
    return KFBO_OK;
}

int kfbo_comm_init(kfbo_handle *h, const void *id_bytes, int32_t rank, int32_t nranks, int32_t shard_mode)
{
    if (!h) return KFBO_EINVAL;
    if (!id_bytes || nranks < 1 || rank < 0 || rank >= nranks ||
        (shard_mode != KFBO_SHARD_TRIALS && shard_mode != KFBO_SHARD_CANDIDATES))
        return fail(h, KFBO_EINVAL, "kfbo_comm_init: bad argument");
    if (h->comm) return fail(h, KFBO_ESTATE, "kfbo_comm_init: communicator already initialised");
    if (!nccl().ok) return fail(h, KFBO_ENCCL, "libnccl.so.2 could not be loaded");
    KFBO_ON_DEVICE(h);
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof id);
    KFBO_NCCL(h, nccl().CommInitRank(&h->comm, nranks, id, rank));
    h->rank = rank; h->nranks = nranks; h->shard_mode = shard_mode;
    h->host_procs = nranks;               // one process per GPU of the box
    return KFBO_OK;
}

// ---- peer-memory exchange: buffers and attachment ------------------------------------------------------------------
static int peer_alloc(kfbo_handle *h)
{
    if (h->d_xchg) return KFBO_OK;
    if (h->nranks > kfbo::kMaxPeers) return fail(h, KFBO_ERANGE, "peer exchange: %d ranks, at most %d (one box)", h->nranks, kfbo::kMaxPeers);
    KFBO_ON_DEVICE(h);
    h->xchg_cap = (size_t)h->cases_cap * 4;
    const size_t bytes = kXchgFlagBytes + 2 * (size_t)h->nranks * h->xchg_cap * sizeof(double);
    KFBO_CUDA(h, cudaMalloc(&h->d_xchg, bytes));
    KFBO_CUDA(h, cudaMemset(h->d_xchg, 0, bytes));
    KFBO_CUDA(h, cudaMalloc(&h->d_done, sizeof(unsigned)));
    KFBO_CUDA(h, cudaMemset(h->d_done, 0, sizeof(unsigned)));
    KFBO_CUDA(h, cudaMallocHost(&h->h_status, sizeof(int32_t)));
    *h->h_status = 0;
    KFBO_CUDA(h, cudaDeviceSynchronize());   // zeroed before anyone can be told about the buffer
    return KFBO_OK;
}

int kfbo_peer_export(kfbo_handle *h, void *handle_bytes)
{
    if (!h) return KFBO_EINVAL;
    if (!handle_bytes) return fail(h, KFBO_EINVAL, "kfbo_peer_export: NULL argument");
    if (!h->comm) return fail(h, KFBO_ESTATE, "kfbo_peer_export: call kfbo_comm_init first (rank and rank count come from it)");
    if (h->peer_ready) return fail(h, KFBO_ESTATE, "kfbo_peer_export: peers already attached");
    const int rc = peer_alloc(h);
    if (rc) return rc;
    cudaIpcMemHandle_t ipc;
    KFBO_CUDA(h, cudaIpcGetMemHandle(&ipc, h->d_xchg));
    std::memcpy(handle_bytes, &ipc, sizeof ipc);
    return KFBO_OK;
}

int kfbo_peer_attach(kfbo_handle *h, const void *all_handle_bytes, int32_t n_handles)
{
    if (!h) return KFBO_EINVAL;
    if (!all_handle_bytes) return fail(h, KFBO_EINVAL, "kfbo_peer_attach: NULL argument");
    if (n_handles != h->nranks)
        return fail(h, KFBO_EINVAL, "kfbo_peer_attach: %d handles for %d ranks (one per rank, in rank order)", n_handles, h->nranks);
    if (!h->d_xchg) return fail(h, KFBO_ESTATE, "kfbo_peer_attach: call kfbo_peer_export first");
    if (h->peer_ready) return fail(h, KFBO_ESTATE, "kfbo_peer_attach: peers already attached");
    KFBO_ON_DEVICE(h);
    for (int r = 0; r < h->nranks; ++r) {
        if (r == h->rank) { h->peer_base[r] = h->d_xchg; continue; }
        cudaIpcMemHandle_t ipc;
        std::memcpy(&ipc, static_cast<const unsigned char *>(all_handle_bytes) + (size_t)r * sizeof ipc, sizeof ipc);
        void *base = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&base, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (h->peer_opened[q]) { cudaIpcCloseMemHandle(h->peer_base[q]); h->peer_opened[q] = false; }
            return fail(h, KFBO_ECUDA, "kfbo_peer_attach: cudaIpcOpenMemHandle of rank %d failed: %s", r, cudaGetErrorString(e));
        }
        h->peer_base[r] = static_cast<unsigned char *>(base);
        h->peer_opened[r] = true;
    }
    h->peer_ready = true;
    return KFBO_OK;
}

int kfbo_last_reduce_path(const kfbo_handle *h) { return h ? h->reduce_path : KFBO_EINVAL; }

// ------------------------------------------------------------------------------------------------------------------
// Device group: ONE host process (one thread) driving several GPUs of the box -- what a single-process caller such as
// the ROS node needs to use more than one GPU.  A group is one handle per device, joined by ncclCommInitAll; an
// evaluation launches on every device, then one grouped ncclAllReduce of the cost sums, then the first device's sums
// are assembled into costs.  Same sharding and the same Philox streams as the one-process-per-GPU mode, so the
// results are the same numbers.
// ------------------------------------------------------------------------------------------------------------------
struct kfbo_group {
    std::vector<kfbo_handle *> hs;
    std::string err;
    bool poisoned = false;      // an evaluation failed part-way: the devices' counters / exchange generations may disagree
    std::string poison_why;
    // One persistent host thread per device beyond the first (KFBO_OPT_GROUP_THREADS, default on): the launch, the exchange and
    // the wait of an evaluation run on all devices at once instead of one device after the other -- a single thread enqueues on 8
    // devices in ~120 us, which is the start-up skew of the last device and 7 % of a 1.7 ms evaluation.  A worker spins on the task
    // generation for kGroupSpinUs after its last task (back-to-back evaluations find it awake), then blocks on the condition
    // variable (an idle node burns no cores).  The caller's thread is device 0's worker.
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint32_t> task_gen{0};
    std::atomic<int> pending{0}, sleepers{0};
    std::atomic<bool> quit{false};
    std::function<int(kfbo_handle *)> task;
    std::vector<int> rcs;
    bool use_threads = true;
};

namespace {

constexpr int kGroupSpinUs = 200;

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    sched_yield();
#endif
}

void group_worker(kfbo_group *g, size_t i)
{
    using clk = std::chrono::steady_clock;
    cudaSetDevice(g->hs[i]->device);          // once: every CUDA call of this thread is for this device
    uint32_t seen = 0;
    for (;;) {
        const clk::time_point t0 = clk::now();
        unsigned spins = 0;
        while (g->task_gen.load() == seen && !g->quit.load()) {
            cpu_relax();
            if ((++spins & 0xFFu) == 0 && clk::now() - t0 > std::chrono::microseconds(kGroupSpinUs)) break;
        }
        if (g->task_gen.load() == seen && !g->quit.load()) {
            std::unique_lock<std::mutex> lk(g->mu);
            g->sleepers.fetch_add(1);
            g->cv.wait(lk, [&] { return g->task_gen.load() != seen || g->quit.load(); });
            g->sleepers.fetch_sub(1);
        }
        if (g->quit.load()) return;
        seen = g->task_gen.load();
        g->rcs[i] = g->task(g->hs[i]);
        g->pending.fetch_sub(1);
    }
}

// fn on every device of the group at once: devices 1.. on their workers, device 0 on the calling thread.  Returns when all are done.
void group_run(kfbo_group *g, const std::function<int(kfbo_handle *)> &fn)
{
    g->task = fn;
    g->pending.store((int)g->workers.size());
    g->task_gen.fetch_add(1);
    if (g->sleepers.load() > 0) {
        std::lock_guard<std::mutex> lk(g->mu);
        g->cv.notify_all();
    }
    g->rcs[0] = fn(g->hs[0]);
    while (g->pending.load() > 0) cpu_relax();
}

void group_stop_workers(kfbo_group *g)
{
    if (g->workers.empty()) return;
    g->quit.store(true);
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->cv.notify_all();
    }
    for (std::thread &t : g->workers) t.join();
    g->workers.clear();
}

}  // namespace

static int group_fail(kfbo_group *g, int code, const std::string &what)
{
    if (g) g->err = what; else g_create_error = what;
    return code;
}

int kfbo_group_create(const kfbo_params *p, int32_t n_devices, const int32_t *devices, int32_t shard_mode, kfbo_group **out)
{
    if (!p || !out) return fail(nullptr, KFBO_EINVAL, "kfbo_group_create: NULL argument");
    *out = nullptr;
    if (shard_mode != KFBO_SHARD_TRIALS && shard_mode != KFBO_SHARD_CANDIDATES)
        return fail(nullptr, KFBO_EINVAL, "kfbo_group_create: shard mode %d", shard_mode);
    int ndev = 0;
    const cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev < 1)
        return fail(nullptr, KFBO_ENODEVICE, "kfbo_group_create: no CUDA device (%s); there is no CPU fallback",
                    ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    if (n_devices < 1 || n_devices > ndev)
        return fail(nullptr, KFBO_ENODEVICE, "kfbo_group_create: %d devices requested, %d present", n_devices, ndev);
    std::vector<int> devs((size_t)n_devices);
    for (int i = 0; i < n_devices; ++i) {
        devs[i] = devices ? devices[i] : i;
        if (devs[i] < 0 || devs[i] >= ndev) return fail(nullptr, KFBO_ENODEVICE, "kfbo_group_create: device %d of %d", devs[i], ndev);
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return fail(nullptr, KFBO_EINVAL, "kfbo_group_create: device %d listed twice", devs[i]);
    }
    kfbo_group *g = new (std::nothrow) kfbo_group;
    if (!g) return fail(nullptr, KFBO_EINVAL, "kfbo_group_create: out of host memory");
    for (int i = 0; i < n_devices; ++i) {
        kfbo_params pi = *p;
        pi.device = devs[i];
        kfbo_handle *h = nullptr;
        const int rc = kfbo_create(&pi, &h);
        if (rc) { kfbo_group_destroy(g); return rc; }      // g_create_error holds kfbo_create's text
        g->hs.push_back(h);
    }
    if (n_devices > 1) {
        if (!nccl().ok) { kfbo_group_destroy(g); return fail(nullptr, KFBO_ENCCL, "libnccl.so.2 could not be loaded"); }
        std::vector<ncclComm_t> comms((size_t)n_devices);
        const ncclResult_t r = nccl().CommInitAll(comms.data(), n_devices, devs.data());
        if (r != ncclSuccess) {
            kfbo_group_destroy(g);
            return fail(nullptr, KFBO_ENCCL, "ncclCommInitAll: %s", nccl().GetErrorString(r));
        }
        for (int i = 0; i < n_devices; ++i) {
            g->hs[i]->comm = comms[i];
            g->hs[i]->rank = i; g->hs[i]->nranks = n_devices; g->hs[i]->shard_mode = shard_mode;
        }
        // Peer-memory exchange between the devices of the group, when every device can map every other's memory
        // (NVLink / NVSwitch boxes); otherwise the group keeps the NCCL call.  Same process: plain device pointers.
        bool p2p = n_devices <= kfbo::kMaxPeers;
        for (int i = 0; i < n_devices && p2p; ++i)
            for (int j = 0; j < n_devices && p2p; ++j) {
                int can = 0;
                if (i != j && (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess || !can)) p2p = false;
            }
        for (int i = 0; i < n_devices && p2p; ++i) {
            DeviceGuard dg(devs[i]);
            if (dg.err != cudaSuccess) { p2p = false; break; }
            for (int j = 0; j < n_devices && p2p; ++j) {
                if (i == j) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) { cudaGetLastError(); p2p = false; }
            }
            if (p2p && peer_alloc(g->hs[i]) != KFBO_OK) p2p = false;
        }
        if (p2p)
            for (int i = 0; i < n_devices; ++i) {
                for (int j = 0; j < n_devices; ++j) g->hs[i]->peer_base[j] = g->hs[j]->d_xchg;
                g->hs[i]->peer_ready = true;
            }
        g->rcs.assign((size_t)n_devices, KFBO_OK);
        try {
            for (int i = 1; i < n_devices; ++i) g->workers.emplace_back(group_worker, g, (size_t)i);
        } catch (...) {        // no threads to be had: the group works without them, one device after the other
            group_stop_workers(g);
        }
    }
    *out = g;
    return KFBO_OK;
}

void kfbo_group_destroy(kfbo_group *g)
{
    if (!g) return;
    group_stop_workers(g);
    for (kfbo_handle *h : g->hs) kfbo_destroy(h);
    delete g;
}

const char *kfbo_group_last_error(const kfbo_group *g) { return g ? g->err.c_str() : g_create_error.c_str(); }
int32_t kfbo_group_size(const kfbo_group *g) { return g ? (int32_t)g->hs.size() : 0; }
kfbo_handle *kfbo_group_handle(kfbo_group *g, int32_t i) { return (g && i >= 0 && i < (int32_t)g->hs.size()) ? g->hs[i] : nullptr; }

static int group_eval_impl(kfbo_group *g, int32_t C, const double *q_est, const double *r_est, kfbo_result *out, double *cost, double *max_cost,
                           int32_t *index_max)
{
    const bool compact = cost != nullptr;
    if (!g || g->hs.empty()) return KFBO_EINVAL;
    if (g->poisoned)
        return group_fail(g, KFBO_ESTATE, "an earlier evaluation of this device group failed part-way (" + g->poison_why +
                                          "): its devices are out of step -- destroy the group and create a new one");
    DeviceGuard caller(g->hs[0]->device);    // whatever device the phases below leave current, the caller gets its own back
    auto check = [g](kfbo_handle *h, int rc) { if (rc) g->err = h->err; return rc; };
    int rc = check(g->hs[0], eval_check(g->hs[0], C, q_est, r_est, compact ? (const void *)cost : (const void *)out));
    if (rc) return rc;                       // nothing was launched anywhere: the group stays usable
    // From here on a failure on one device leaves the others mid-evaluation (launched, possibly spinning on the failed
    // device's flag until the exchange times out): every device is drained and the group refuses further evaluations --
    // counters and exchange generations advance only when ALL devices got through.
    auto abort_all = [g](int code) {
        for (kfbo_handle *h : g->hs) poison(h);      // no-op for a group of one device (no communicator)
        if (g->hs.size() > 1 && !g->poisoned) { g->poisoned = true; g->poison_why = g->err; }
        return code;
    };
    const uint8_t *bad = g->hs[0]->cand_bad.data();
    for (kfbo_handle *h : g->hs) h->n_bad = g->hs[0]->n_bad;
    // All devices at once (one host thread per device) when the sums will meet in peer memory: launch, exchange kernel and
    // wait are independent per device there -- the exchange itself synchronises the devices, as between processes.  (The NCCL
    // fallback keeps the one-thread order below: its grouped call is made from one thread.)
    bool peer_path = true;
    for (kfbo_handle *h : g->hs) peer_path = peer_path && sums_meet_in_peer_memory(h, (size_t)C * h->D * 4);
    if (g->use_threads && !g->workers.empty() && peer_path) {
        group_run(g, [&](kfbo_handle *h) -> int {
            int r = eval_launch(h, C, q_est, r_est, bad, compact);
            if (!r && h->reduce_path < 2) r = fail(h, KFBO_ESTATE, "device group: the devices did not take the peer-memory path together");
            if (!r) r = eval_reduce(h, C, false);
            if (r) return r;       // this device is out; the others run into the exchange's timeout and report KFBO_EPEER
            return h == g->hs[0] ? eval_collect(h, C, out, cost, max_cost, index_max) : eval_collect(h, C, nullptr);
        });
        for (size_t i = 0; i < g->hs.size(); ++i)
            if (g->rcs[i]) return abort_all(check(g->hs[i], g->rcs[i]));
        return KFBO_OK;
    }
    for (kfbo_handle *h : g->hs)
        if ((rc = check(h, eval_launch(h, C, q_est, r_est, bad, compact)))) return abort_all(rc);
    if (g->hs.size() > 1 && g->hs[0]->reduce_path >= 2) {   // the sums met in peer memory (every device took the same path)
        for (kfbo_handle *h : g->hs)
            if ((rc = check(h, eval_reduce(h, C, false)))) return abort_all(rc);
    } else if (g->hs.size() > 1) {
        ncclResult_t r = nccl().GroupStart();
        if (r != ncclSuccess) return abort_all(group_fail(g, KFBO_ENCCL, std::string("ncclGroupStart: ") + nccl().GetErrorString(r)));
        for (kfbo_handle *h : g->hs)
            if ((rc = check(h, eval_reduce(h, C, true)))) break;
        r = nccl().GroupEnd();
        if (rc) return abort_all(rc);
        if (r != ncclSuccess) return abort_all(group_fail(g, KFBO_ENCCL, std::string("ncclGroupEnd: ") + nccl().GetErrorString(r)));
        for (kfbo_handle *h : g->hs) {
            DeviceGuard dg(h->device);
            if (dg.err != cudaSuccess || cudaEventRecord(h->ev2, h->stream) != cudaSuccess)
                return abort_all(group_fail(g, KFBO_ECUDA, "cudaEventRecord after the grouped all-reduce failed"));
        }
    }
    // every device is waited for even when one reports a failure (KFBO_EPEER on one device: the others have timed out or
    // will); the first failure is the one reported
    int first = KFBO_OK;
    for (size_t i = 0; i < g->hs.size(); ++i) {
        const int r1 = i == 0 ? eval_collect(g->hs[i], C, out, cost, max_cost, index_max) : eval_collect(g->hs[i], C, nullptr);
        if (r1 && !first) first = check(g->hs[i], r1);
    }
    return first ? abort_all(first) : KFBO_OK;
}

int kfbo_group_eval_batch(kfbo_group *g, int32_t C, const double *q_est, const double *r_est, kfbo_result *out)
{
    if (!out) return group_fail(g, KFBO_EINVAL, "kfbo_group_eval_batch: NULL argument");
    return group_eval_impl(g, C, q_est, r_est, out, nullptr, nullptr, nullptr);
}

int kfbo_group_eval_batch_costs(kfbo_group *g, int32_t C, const double *q_est, const double *r_est, double *cost, double *max_cost, int32_t *index_max)
{
    if (!cost || !max_cost || !index_max) return group_fail(g, KFBO_EINVAL, "kfbo_group_eval_batch_costs: NULL argument");
    return group_eval_impl(g, C, q_est, r_est, nullptr, cost, max_cost, index_max);
}

int kfbo_group_eval(kfbo_group *g, const double *pn, int32_t npn, const double *on, int32_t non, kfbo_result *out)
{
    if (!g) return KFBO_EINVAL;
    if (!pn || !on || !out || npn < 1 || non < 1)
        return group_fail(g, KFBO_EINVAL, "kfbo_group_eval: needs pn[>=1], on[>=1] (system_models.hpp:39,143 read element 0)");
    return kfbo_group_eval_batch(g, 1, pn, on, out);
}

int kfbo_group_last_sums(const kfbo_group *g, double *sums, int32_t n_candidates)
{
    return (g && !g->hs.empty()) ? kfbo_last_sums(g->hs[0], sums, n_candidates) : KFBO_EINVAL;
}

int kfbo_group_last_kernel_ms(const kfbo_group *g, double *ms_max, int32_t *launches)
{
    if (!g || g->hs.empty()) return KFBO_EINVAL;
    double mx = 0.0;
    int32_t n = 0;
    for (kfbo_handle *h : g->hs) { resolve_timing(h); mx = std::max(mx, h->last_ms); n += h->last_launches; }
    if (ms_max) *ms_max = mx;
    if (launches) *launches = n;
    return KFBO_OK;
}

int kfbo_group_set_option(kfbo_group *g, int32_t option, int64_t value)
{
    if (!g) return KFBO_EINVAL;
    if (option == KFBO_OPT_GROUP_THREADS) { g->use_threads = value != 0; return KFBO_OK; }
    for (kfbo_handle *h : g->hs) {
        const int rc = kfbo_set_option(h, option, value);
        if (rc) { g->err = h->err; return rc; }
    }
    return KFBO_OK;
}

int kfbo_group_set_stream_base(kfbo_group *g, uint64_t base)
{
    if (!g) return KFBO_EINVAL;
    for (kfbo_handle *h : g->hs) h->stream_base = base;
    return KFBO_OK;
}

int kfbo_group_set_eval_counter(kfbo_group *g, uint64_t e)
{
    if (!g) return KFBO_EINVAL;
    for (kfbo_handle *h : g->hs) h->eval_counter = e;
    return KFBO_OK;
}

int kfbo_measure_fp64_peak(int32_t device, double millis, double *tflops, double *tflops_uniform_operands)
{
    if (!tflops || !(millis > 0)) return KFBO_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(nullptr, KFBO_ENODEVICE, "no CUDA device");
    if (device < 0) cudaGetDevice(&device);
    DeviceGuard dg(device);
    if (dg.err != cudaSuccess) return fail(nullptr, KFBO_ECUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, iters = 2048;
    double *d = nullptr;
    cudaEvent_t e0, e1;
    cudaStream_t s;
    if (cudaMalloc(&d, 64) != cudaSuccess) return fail(nullptr, KFBO_ECUDA, "cudaMalloc failed");
    cudaStreamCreate(&s); cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double flops_per_launch = 2.0 * 8 * 16 * (double)iters * 256.0 * blocks;
    int rc = KFBO_OK;
    // the register-multiplier form is the peak (2.0 pipe cycles per warp instruction); the all-uniform form, which takes
    // 2.2, is reported beside it when asked for
    for (int form = 0; form < (tflops_uniform_operands ? 2 : 1) && rc == KFBO_OK; ++form) {
        const bool reg = form == 0;
        kfbo::launch_dfma_chain(d, blocks, iters, reg, s);  // warm-up
        cudaStreamSynchronize(s);
        double total_ms = 0.0, total_flops = 0.0;
        const double budget = reg ? millis : std::min(millis, 100.0);
        while (total_ms < budget) {
            cudaEventRecord(e0, s);
            for (int i = 0; i < 4; ++i) kfbo::launch_dfma_chain(d, blocks, iters, reg, s);
            cudaEventRecord(e1, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) { rc = fail(nullptr, KFBO_ECUDA, "dfma_chain failed"); break; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            total_ms += ms; total_flops += 4 * flops_per_launch;
        }
        if (rc == KFBO_OK) *(reg ? tflops : tflops_uniform_operands) = total_flops / (total_ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s); cudaFree(d);
    return rc;
}

}  // extern "C"
