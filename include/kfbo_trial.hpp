// kfbo_trial.hpp -- C++ host side of the B200 cost evaluator: the reference's own class interface
// (ReadParaVehicle, Trial, RunTrial and the two std_msgs types on the "noise"/"cost" topics) as a
// header-only layer over the C ABI of kfbo.h.  Link with -lkfbo (kf_bayesopt_b200/libkfbo.so).
//
// What each class replaces (file:line relative to the reference tree):
//   kfbo::ReadParaVehicle   include/read_para_vehicle.h:9-49, read_para_vehicle.cpp:119-158
//                           (filled from config/vehicleParams.yaml directly instead of the ROS
//                           parameter server; same member names)
//   kfbo::Trial             include/trial.h:7-59, trial.cpp:12-131: same constructor, Run() and the
//                           four getters; Run() is one std::thread's share of the Monte Carlo batch
//                           (nsimruns_/cpu_core_number_ trials, trial.cpp:28), rolled out on the GPU
//   kfbo::RunTrial          trial_node.cpp:13-142: ProcessNoiseCallback("noise") -> "cost"; the
//                           CPUcoreNumber thread pool (trial_node.cpp:54-67) becomes ONE batched launch
//                           over trials x sample times on the device
//   kfbo::msg::*            std_msgs/Float64MultiArray and std_msgs/Float64 as robot1d_bayesopt.cpp:53-92
//                           and trial_node.cpp:47,127-129 use them (field-for-field stand-ins; with
//                           -DKFBO_WITH_ROS the real types are used instead)
//
// Errors: the C ABI never throws; this layer turns a non-zero status into kfbo::Error.
#ifndef KFBO_TRIAL_HPP_
#define KFBO_TRIAL_HPP_

#include <atomic>
#include <array>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "kfbo.h"

#ifdef KFBO_WITH_ROS
#include <std_msgs/Float64.h>
#include <std_msgs/Float64MultiArray.h>
#endif

namespace kfbo {

class Error : public std::runtime_error {
public:
    Error(int code, const std::string &what) : std::runtime_error("kfbo error " + std::to_string(code) + ": " + what), code_(code) {}
    int code() const { return code_; }
private:
    int code_;
};

// ------------------------------------------------------------------------------------------
// A reader for the flat yaml the reference ships (config/vehicleParams.yaml, config/bayesParams.yaml):
// "key: scalar" and "key: [a, b, ...]" lines, '#' comments.  Values are kept as strings.
// ------------------------------------------------------------------------------------------
class FlatYaml {
public:
    static FlatYaml Load(const std::string &path)
    {
        std::ifstream f(path);
        if (!f) throw Error(KFBO_EINVAL, "cannot open " + path);
        FlatYaml y;
        std::string line;
        while (std::getline(f, line)) {
            // a '#' starts a comment at the start of the line or after whitespace (yaml rule)
            for (size_t i = 0; i < line.size(); ++i)
                if (line[i] == '#' && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) { line.erase(i); break; }
            const size_t c = line.find(':');
            if (c == std::string::npos) continue;
            const std::string key = Trim(line.substr(0, c)), val = Trim(line.substr(c + 1));
            if (!key.empty()) y.kv_[key] = val;
        }
        return y;
    }
    bool Has(const std::string &k) const { return kv_.count(k) != 0; }
    std::string Str(const std::string &k, const std::string &dflt) const
    {
        auto it = kv_.find(k);
        if (it == kv_.end()) return dflt;
        std::string v = it->second;
        if (v.size() >= 2 && (v.front() == '"' || v.front() == '\'') && v.back() == v.front()) v = v.substr(1, v.size() - 2);
        return v;
    }
    double Num(const std::string &k, double dflt) const
    {
        auto it = kv_.find(k);
        if (it == kv_.end() || it->second.empty()) return dflt;
        try { return std::stod(it->second); } catch (...) { throw Error(KFBO_EINVAL, "yaml key '" + k + "' is not a number: " + it->second); }
    }
    std::vector<double> List(const std::string &k, const std::vector<double> &dflt) const
    {
        auto it = kv_.find(k);
        if (it == kv_.end()) return dflt;
        std::string v = it->second;
        if (v.empty() || v.front() != '[' || v.back() != ']') throw Error(KFBO_EINVAL, "yaml key '" + k + "' is not a [list]");
        v = v.substr(1, v.size() - 2);
        std::vector<double> out;
        std::stringstream ss(v);
        std::string item;
        while (std::getline(ss, item, ',')) {
            item = Trim(item);
            if (!item.empty()) out.push_back(std::stod(item));
        }
        return out;
    }
private:
    static std::string Trim(const std::string &s)
    {
        const size_t b = s.find_first_not_of(" \t\r\n");
        if (b == std::string::npos) return "";
        return s.substr(b, s.find_last_not_of(" \t\r\n") - b + 1);
    }
    std::map<std::string, std::string> kv_;
};

// ------------------------------------------------------------------------------------------
// ReadParaVehicle: the members of include/read_para_vehicle.h:9-49, defaults of config/vehicleParams.yaml.
// ------------------------------------------------------------------------------------------
class ReadParaVehicle {
public:
    ReadParaVehicle() { UpdateMaxIterations(); }

    // read_para_vehicle.cpp:119-158, from the yaml file itself (the launch files only rosparam-load it)
    static ReadParaVehicle FromYaml(const std::string &path, bool show_read_info = false)
    {
        const FlatYaml y = FlatYaml::Load(path);
        ReadParaVehicle rv;
        rv.cpu_core_number_ = (int)y.Num("CPUcoreNumber", rv.cpu_core_number_);
        rv.state_dof_ = (int)y.Num("stateDOF", rv.state_dof_);
        rv.observation_dof_ = (int)y.Num("observationDOF", rv.observation_dof_);
        rv.t_start_ = y.Num("t_start", rv.t_start_);
        rv.dt_ = y.Num("dt", rv.dt_);
        rv.t_end_ = y.Num("t_end", rv.t_end_);
        rv.xi0_ = y.Num("xi0", rv.xi0_);
        rv.xidot0_ = y.Num("xidot0", rv.xidot0_);
        rv.disturbance_ = y.List("disturbance", {});
        rv.P0_ = y.Str("P0", rv.P0_);
        rv.nsimruns_ = (int)y.Num("Nsimruns", 10);            // read_para_vehicle.h:34 default
        rv.pnoise_ = y.List("pnoise", rv.pnoise_);
        rv.onoise_ = y.List("onoise", rv.onoise_);
        rv.dim_pn_ = (int)rv.pnoise_.size();
        rv.dim_on_ = (int)rv.onoise_.size();
        rv.optimization_choice_ = y.Str("optimizationChoice", rv.optimization_choice_);
        rv.cost_choice_ = y.Str("costChoice", rv.cost_choice_);
        if (rv.pnoise_.empty() || rv.onoise_.empty()) throw Error(KFBO_EINVAL, path + ": pnoise / onoise must hold at least one value");
        if (rv.cpu_core_number_ > 0 && rv.nsimruns_ % rv.cpu_core_number_ != 0)   // read_para_vehicle.cpp:141-142
            std::cerr << "Nsimruns is not a multiple of CPUcoreNumber: the remainder trials are dropped (trial.cpp:28)\n";
        rv.UpdateMaxIterations();
        if (show_read_info) rv.Print(std::cout);
        return rv;
    }

    // read_para_vehicle.cpp:145 / trial_node.cpp:113: truncating double -> int
    void UpdateMaxIterations() { max_iterations_ = (int)((t_end_ - t_start_) / dt_); }

    void Print(std::ostream &os) const
    {
        os << "CPUcoreNumber " << cpu_core_number_ << ", stateDOF " << state_dof_ << ", observationDOF " << observation_dof_
           << ", t_start " << t_start_ << ", dt " << dt_ << ", t_end " << t_end_ << ", max_iterations " << max_iterations_
           << ", Nsimruns " << nsimruns_ << ", pnoise[0] " << pnoise_[0] << ", onoise[0] " << onoise_[0]
           << ", optimizationChoice " << optimization_choice_ << ", costChoice " << cost_choice_ << "\n";
    }

    int cpu_core_number_ = 10;
    int state_dof_ = 2;
    int observation_dof_ = 1;
    double t_start_ = 0.0;
    double dt_ = 0.1;
    double t_end_ = 20.1;
    double xi0_ = 0.0;       // read but never used by the hot path: trial.cpp:31 hard-codes x0 = 0
    double xidot0_ = 0.0;
    std::vector<double> disturbance_;
    std::string P0_ = "identity";   // read but never used: trial.cpp:32 hard-codes P0 = I
    int nsimruns_ = 200;
    int max_iterations_ = 201;
    int dim_pn_ = 1;
    int dim_on_ = 1;
    std::vector<double> pnoise_{1.0};
    std::vector<double> onoise_{0.1};
    std::string optimization_choice_ = "all";
    std::string cost_choice_ = "JNEES";
};

inline int CostChoiceFromString(const std::string &s)
{
    if (s == "JNEES") return KFBO_JNEES;
    if (s == "CNEES") return KFBO_CNEES;
    if (s == "JNIS") return KFBO_JNIS;
    if (s == "CNIS") return KFBO_CNIS;
    throw Error(KFBO_EINVAL, "costChoice must be JNEES, CNEES, JNIS or CNIS, got '" + s + "'");
}

// (dt, t_end) the callback overwrites its parameters with after pass 0 (trial_node.cpp:105-108) and the
// variant it keeps commented out (trial_node.cpp:109-112; README.md:96 and BASELINE.json quote this one).
struct SecondSampleTime { double dt, t_end; };
constexpr SecondSampleTime kCodeSecondSampleTime{1.0, 201.0};
constexpr SecondSampleTime kReadmeSecondSampleTime{0.5, 105.5};

// ------------------------------------------------------------------------------------------
// Owning wrapper of a kfbo_handle.
// ------------------------------------------------------------------------------------------
class Evaluator {
public:
    // n_gpus > 1: ONE process drives n_gpus devices (a kfbo_group: each evaluation launches its shard on every
    // device and ends in one grouped ncclAllReduce of the cost sums) -- what a single-process caller like the ROS node
    // needs to use the whole box.  shard_mode: KFBO_SHARD_TRIALS (one candidate, many trials) or
    // KFBO_SHARD_CANDIDATES (candidate sweeps).
    explicit Evaluator(const kfbo_params &p, int n_gpus = 1, int shard_mode = KFBO_SHARD_TRIALS) : p_(p)
    {
        if (n_gpus > 1) {
            const int rc = kfbo_group_create(&p_, n_gpus, nullptr, shard_mode, &g_);
            if (rc) throw Error(rc, kfbo_group_last_error(nullptr));
            h_ = kfbo_group_handle(g_, 0);
        } else {
            const int rc = kfbo_create(&p_, &h_);
            if (rc) throw Error(rc, kfbo_last_error(nullptr));
        }
    }
    ~Evaluator()
    {
        if (g_) kfbo_group_destroy(g_);
        else kfbo_destroy(h_);
    }
    Evaluator(const Evaluator &) = delete;
    Evaluator &operator=(const Evaluator &) = delete;

    kfbo_result Eval(const std::vector<double> &pn, const std::vector<double> &on)
    {
        kfbo_result r;
        Check(g_ ? kfbo_group_eval(g_, pn.data(), (int32_t)pn.size(), on.data(), (int32_t)on.size(), &r)
                 : kfbo_eval(h_, pn.data(), (int32_t)pn.size(), on.data(), (int32_t)on.size(), &r));
        return r;
    }
    std::vector<kfbo_result> EvalBatch(const std::vector<double> &q_est, const std::vector<double> &r_est)
    {
        if (q_est.size() != r_est.size() || q_est.empty()) throw Error(KFBO_EINVAL, "EvalBatch: q_est and r_est must have equal, non-zero length");
        std::vector<kfbo_result> out(q_est.size());
        Check(g_ ? kfbo_group_eval_batch(g_, (int32_t)q_est.size(), q_est.data(), r_est.data(), out.data())
                 : kfbo_eval_batch(h_, (int32_t)q_est.size(), q_est.data(), r_est.data(), out.data()));
        return out;
    }
    // The same batch with the cost assembly on the device and a compact answer (kfbo_eval_batch_costs): what a cost-surface
    // sweep wants -- 16 bytes per candidate in, 8 (D + 1) + 4 out, no per-candidate host work.
    struct BatchCosts { std::vector<double> cost /*[C][D]*/, max_cost /*[C]*/; std::vector<int32_t> index_max /*[C]*/; };
    BatchCosts EvalBatchCosts(const std::vector<double> &q_est, const std::vector<double> &r_est)
    {
        if (q_est.size() != r_est.size() || q_est.empty()) throw Error(KFBO_EINVAL, "EvalBatchCosts: q_est and r_est must have equal, non-zero length");
        const size_t C = q_est.size();
        BatchCosts b{std::vector<double>(C * p_.n_sample_times), std::vector<double>(C), std::vector<int32_t>(C)};
        Check(g_ ? kfbo_group_eval_batch_costs(g_, (int32_t)C, q_est.data(), r_est.data(), b.cost.data(), b.max_cost.data(), b.index_max.data())
                 : kfbo_eval_batch_costs(h_, (int32_t)C, q_est.data(), r_est.data(), b.cost.data(), b.max_cost.data(), b.index_max.data()));
        return b;
    }
    // sums[n_candidates][D][4] of the last evaluation
    std::vector<double> LastSums(int n_candidates = 1) const
    {
        std::vector<double> s((size_t)n_candidates * p_.n_sample_times * 4);
        Check(g_ ? kfbo_group_last_sums(g_, s.data(), n_candidates) : kfbo_last_sums(h_, s.data(), n_candidates));
        return s;
    }
    // device time of the last evaluation's rollout kernel(s): the slowest device of a group
    double LastKernelMs(int *launches = nullptr) const
    {
        double ms = 0;
        int32_t n = 0;
        Check(g_ ? kfbo_group_last_kernel_ms(g_, &ms, &n) : kfbo_last_kernel_ms(h_, &ms, &n));
        if (launches) *launches = n;
        return ms;
    }
    // The Nsimruns == 1 record of trial.cpp:69-78: rows {time, e1, e2, lb1, ub1, lb2, ub2, nees, nis} of one trajectory.
    std::vector<double> TraceTrial(int k, double q_est, double r_est, uint64_t stream, uint32_t trial)
    {
        std::vector<double> t((size_t)p_.max_iterations[k] * KFBO_TRACE_COLUMNS);
        Check(kfbo_trace_trial(h_, k, q_est, r_est, stream, trial, t.data()));
        return t;
    }
    void SetOption(int option, long long value) { Check(g_ ? kfbo_group_set_option(g_, option, value) : kfbo_set_option(h_, option, value)); }
    void SetStreamBase(uint64_t base) { Check(g_ ? kfbo_group_set_stream_base(g_, base) : kfbo_set_stream_base(h_, base)); }
    void SetEvalCounter(uint64_t e) { Check(g_ ? kfbo_group_set_eval_counter(g_, e) : kfbo_set_eval_counter(h_, e)); }
    // one process per GPU instead (MPI / torchrun style): join the NCCL communicator of the ranks
    void CommInit(const void *nccl_id, int rank, int nranks, int shard_mode)
    {
        if (g_) throw Error(KFBO_ESTATE, "CommInit: this evaluator already drives a device group");
        Check(kfbo_comm_init(h_, nccl_id, rank, nranks, shard_mode));
    }
    // ... and, on one NVLink box, let the cost sums meet in the peers' memory instead of the NCCL call: every rank exports
    // the handle of its exchange buffer, the caller all-gathers the KFBO_PEER_HANDLE_BYTES of each (MPI_Allgather, ...),
    // every rank attaches to all of them (rank order).  kfbo.h: kfbo_peer_export / kfbo_peer_attach.
    std::array<unsigned char, KFBO_PEER_HANDLE_BYTES> PeerExport()
    {
        std::array<unsigned char, KFBO_PEER_HANDLE_BYTES> hb{};
        Check(kfbo_peer_export(h_, hb.data()));
        return hb;
    }
    void PeerAttach(const std::vector<std::array<unsigned char, KFBO_PEER_HANDLE_BYTES>> &all_ranks)
    {
        Check(kfbo_peer_attach(h_, all_ranks.data(), (int32_t)all_ranks.size()));
    }
    // how the sums of the last evaluation met: 0 one rank, 1 ncclAllReduce, 2 / 3 peer memory (behind / inside the rollout kernel)
    int LastReducePath() const { return kfbo_last_reduce_path(g_ ? kfbo_group_handle(g_, 0) : h_); }
    const kfbo_params &params() const { return p_; }
    kfbo_handle *handle() { return h_; }          // of a group: its first device
    int n_gpus() const { return g_ ? (int)kfbo_group_size(g_) : 1; }

private:
    void Check(int rc) const
    {
        if (rc) throw Error(rc, g_ && *kfbo_group_last_error(g_) ? kfbo_group_last_error(g_) : kfbo_last_error(h_));
    }
    kfbo_params p_;
    kfbo_handle *h_ = nullptr;
    kfbo_group *g_ = nullptr;
};

// kfbo_params for the truth side `rv_gt` and a list of (dt, max_iterations) passes.
inline kfbo_params MakeParams(const ReadParaVehicle &rv_gt, const std::vector<std::pair<double, int>> &sample_times,
                              const std::string &cost_choice, uint64_t seed, int max_candidates = 1, int device = -1)
{
    if (sample_times.empty() || sample_times.size() > KFBO_MAX_SAMPLE_TIMES) throw Error(KFBO_ERANGE, "1..4 sample times");
    kfbo_params p;
    kfbo_default_params(&p);
    p.n_sample_times = (int32_t)sample_times.size();
    for (size_t k = 0; k < sample_times.size(); ++k) {
        p.dt[k] = sample_times[k].first;
        p.max_iterations[k] = sample_times[k].second;
    }
    p.nsimruns = rv_gt.nsimruns_;
    p.cpu_core_number = rv_gt.cpu_core_number_;
    p.state_dof = rv_gt.state_dof_;
    p.observation_dof = rv_gt.observation_dof_;
    p.pnoise_truth = rv_gt.pnoise_.at(0);        // system_models.hpp:39
    p.onoise_truth = rv_gt.onoise_.at(0);        // system_models.hpp:143
    p.cost_choice = CostChoiceFromString(cost_choice);
    p.device = device;
    p.seed = seed;
    p.max_candidates = max_candidates;
    return p;
}

// The two passes of RunTrial::ProcessNoiseCallback as data (trial_node.cpp:52,105-114).
inline std::vector<std::pair<double, int>> SampleTimesFrom(const ReadParaVehicle &rv, const SecondSampleTime *second)
{
    std::vector<std::pair<double, int>> out;
    out.emplace_back(rv.dt_, (int)((rv.t_end_ - rv.t_start_) / rv.dt_));
    if (second) out.emplace_back(second->dt, (int)((second->t_end - rv.t_start_) / second->dt));
    return out;
}

constexpr uint64_t kDefaultSeed = 0x4B46424F20260101ull;

namespace detail {
inline std::atomic<uint64_t> &TrialStreamCounter()
{
    static std::atomic<uint64_t> c{KFBO_TRIAL_STREAM_BASE};   // the Trial adapters' own range of the 62-bit stream ids (kfbo.h)
    return c;
}

// Evaluator handles shared by the Trial objects of a process.  The reference's callers build CPUcoreNumber fresh
// Trial objects per sample time per callback (trial_node.cpp:55-63); creating a device handle for each would make
// every callback pay for allocations.  A Trial borrows a handle with its parameters for the duration of Run() and
// gives it back; handles are created on demand (as many as Trials run concurrently).  The pool is never destroyed:
// handles must not be torn down after the CUDA runtime at process exit.
class EvaluatorPool {
public:
    static EvaluatorPool &Get()
    {
        static EvaluatorPool *pool = new EvaluatorPool;
        return *pool;
    }
    std::unique_ptr<Evaluator> Borrow(const kfbo_params &p)
    {
        const std::string key = Key(p);
        {
            std::lock_guard<std::mutex> lock(mu_);
            std::vector<std::unique_ptr<Evaluator>> &free_list = free_[key];
            if (!free_list.empty()) {
                std::unique_ptr<Evaluator> ev = std::move(free_list.back());
                free_list.pop_back();
                return ev;
            }
        }
        return std::unique_ptr<Evaluator>(new Evaluator(p));
    }
    void Return(std::unique_ptr<Evaluator> ev)
    {
        const std::string key = Key(ev->params());
        std::lock_guard<std::mutex> lock(mu_);
        std::vector<std::unique_ptr<Evaluator>> &free_list = free_[key];
        if (free_list.size() < kMaxIdlePerKey) free_list.push_back(std::move(ev));
    }

private:
    static constexpr size_t kMaxIdlePerKey = 64;
    static std::string Key(const kfbo_params &p)
    {
        std::ostringstream k;
        k << std::hexfloat << p.n_sample_times;
        for (int i = 0; i < p.n_sample_times; ++i) k << '|' << p.dt[i] << ':' << p.max_iterations[i];
        k << '|' << p.nsimruns << '|' << p.cpu_core_number << '|' << p.state_dof << '|' << p.observation_dof << '|' << p.pnoise_truth
          << '|' << p.onoise_truth << '|' << p.cost_choice << '|' << p.device << '|' << p.seed << '|' << p.max_candidates;
        return k.str();
    }
    std::mutex mu_;
    std::map<std::string, std::vector<std::unique_ptr<Evaluator>>> free_;
};
}  // namespace detail

// ------------------------------------------------------------------------------------------
// Trial: include/trial.h:7-59.  One Trial is ONE thread's share of the Monte Carlo batch.  The
// callers keep their shape (trial_node.cpp:55-75, src/test_trial.cpp:41-51):
//     std::vector<Trial> t; for (cores) t.push_back(Trial(rv_gt, rv));
//     std::thread(&Trial::Run, &t[i]) ... join ... t[i].get_average_nees() / cores
// Run() is thread-safe across distinct Trial objects (each borrows a pooled handle and takes a fresh Philox stream).
// ------------------------------------------------------------------------------------------
class Trial {
public:
    Trial(ReadParaVehicle &simulator_parameters, ReadParaVehicle &estimator_parameters, int device = -1,
          uint64_t seed = kDefaultSeed)
        : param_vehicle_(simulator_parameters), estimator_parameters_(estimator_parameters), device_(device), seed_(seed)
    {
    }

    // Like the reference's, Run() is void and runs on a std::thread: a failure (no CUDA device, ...)
    // must not escape the thread, so it is recorded instead -- ok()/error() -- and the getters stay NaN.
    void Run() noexcept
    {
        try {
            RunOrThrow();
        } catch (const std::exception &e) {
            error_ = e.what();
        } catch (...) {
            error_ = "unknown failure";
        }
    }
    bool ok() const { return error_.empty(); }
    const std::string &error() const { return error_; }

    void RunOrThrow()
    {
        error_.clear();
        const ReadParaVehicle &sim = param_vehicle_;
        const int n = sim.nsimruns_ / sim.cpu_core_number_;                                   // trial.cpp:28
        const int T = sim.max_iterations_;                                                    // trial.cpp:45
        k_average_ = (double)((long long)T * sim.nsimruns_ / sim.cpu_core_number_);           // trial.cpp:19
        ReadParaVehicle one = sim;
        one.nsimruns_ = n;
        one.cpu_core_number_ = 1;
        // a pooled handle for these parameters; every Run draws its own noise (a fresh Philox stream)
        std::unique_ptr<Evaluator> ev = detail::EvaluatorPool::Get().Borrow(MakeParams(one, {{sim.dt_, T}}, "JNEES", seed_, 1, device_));
        ev->SetStreamBase(detail::TrialStreamCounter().fetch_add(1));
        ev->SetEvalCounter(0);
        ev->Eval(estimator_parameters_.pnoise_, estimator_parameters_.onoise_);
        const std::vector<double> s = ev->LastSums(1);
        detail::EvaluatorPool::Get().Return(std::move(ev));
        all_nees_ = s[0];
        all_nis_ = s[1];
        average_var_nees_ = s[2] * sim.cpu_core_number_ / sim.nsimruns_;                      // trial.cpp:110
        average_var_nis_ = s[3] * sim.cpu_core_number_ / sim.nsimruns_;                       // trial.cpp:111
        average_nees_ = all_nees_ / k_average_;                                               // trial.cpp:112
        average_nis_ = all_nis_ / k_average_;                                                 // trial.cpp:113
    }

    double get_average_nees() const { return average_nees_; }
    double get_average_nis() const { return average_nis_; }
    double get_average_var_nees() const { return average_var_nees_; }
    double get_average_var_nis() const { return average_var_nis_; }

    double det_p_ = 0;   // public members of the reference class (include/trial.h:35-36); only its plotting path used them
    double det_s_ = 0;

private:
    double average_nees_ = NAN, average_nis_ = NAN, average_var_nees_ = NAN, average_var_nis_ = NAN;
    double all_nees_ = 0, all_nis_ = 0, k_average_ = 0;
    ReadParaVehicle param_vehicle_, estimator_parameters_;
    int device_;
    uint64_t seed_;
    std::string error_;
};

// ------------------------------------------------------------------------------------------
// Messages on the two topics.
// ------------------------------------------------------------------------------------------
namespace msg {
#ifdef KFBO_WITH_ROS
using Float64 = std_msgs::Float64;
using Float64MultiArray = std_msgs::Float64MultiArray;
using MultiArrayDimension = std_msgs::MultiArrayDimension;
#else
struct MultiArrayDimension { std::string label; uint32_t size = 0; uint32_t stride = 0; };
struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; uint32_t data_offset = 0; };
struct Float64MultiArray { MultiArrayLayout layout; std::vector<double> data; };
struct Float64 { double data = 0.0; };
#endif

// The "noise" message exactly as Robot1D::evaluateSample builds it (robot1d_bayesopt.cpp:53-88).
inline Float64MultiArray MakeNoise(const std::vector<double> &pn, const std::vector<double> &on)
{
    Float64MultiArray m;
    m.layout.dim.push_back(MultiArrayDimension());
    m.layout.dim.push_back(MultiArrayDimension());
    m.layout.dim[0].size = (uint32_t)pn.size();
    m.layout.dim[0].label = "processNoiseDimension";
    m.layout.dim[1].size = (uint32_t)on.size();
    m.layout.dim[1].label = "measurmentNoiseDimension";
    m.data = pn;
    m.data.insert(m.data.end(), on.begin(), on.end());
    return m;
}
}  // namespace msg

// ------------------------------------------------------------------------------------------
// RunTrial: trial_node.cpp:13-142 without the ROS node handle.  The evaluator (device buffers,
// tables, stream) is built once here instead of per callback like the reference's vector<Trial>.
// ------------------------------------------------------------------------------------------
class RunTrial {
public:
    using Publisher = std::function<void(const msg::Float64 &)>;

    explicit RunTrial(const ReadParaVehicle &rv, const SecondSampleTime *second = &kCodeSecondSampleTime,
                      int n_init_samples = 20, Publisher pub_cost = nullptr, bool verbose = true, int device = -1,
                      uint64_t seed = kDefaultSeed, int max_candidates = 1, int n_gpus = 1)
        : rv_(rv), n_init_samples_(n_init_samples), pub_cost_(std::move(pub_cost)), verbose_(verbose),
          // single messages shard their trials over the GPUs; ProcessNoiseBatch sweeps shard the same way (every
          // device runs every candidate on its slice of the trials)
          ev_(new Evaluator(MakeParams(rv, SampleTimesFrom(rv, second), rv.cost_choice_, seed, max_candidates, device), n_gpus,
                            KFBO_SHARD_TRIALS))
    {
    }

    // One "noise" message in, one "cost" message out (and published, when a publisher is set).
    msg::Float64 ProcessNoiseCallback(const msg::Float64MultiArray &m)
    {
        if (m.layout.dim.empty() || m.layout.dim[0].size < 1 || m.layout.dim[0].size >= m.data.size())
            throw Error(KFBO_EINVAL, "noise message: layout.dim[0].size must split data into [pn..., on...]");
        if (verbose_) std::cout << "get the noise....................." << std::endl;
        const int npn = (int)m.layout.dim[0].size;                                   // trial_node.cpp:29
        const std::vector<double> pn(m.data.begin(), m.data.begin() + npn);          // :32
        const std::vector<double> on(m.data.begin() + npn, m.data.end());            // :33
        if (verbose_) {
            std::cout << "process noise ";
            for (double v : pn) std::cout << v << " ";
            std::cout << std::endl << "observation noise " << on[0] << std::endl;
        }
        last_ = ev_->Eval(pn, on);
        const int D = ev_->params().n_sample_times;
        if (verbose_) {
            const bool nees = rv_.cost_choice_ == "JNEES" || rv_.cost_choice_ == "CNEES";
            for (int k = 0; k < D; ++k) std::cout << (nees ? "cost JNEES " : "cost JNIS ") << last_.cost[k] << std::endl;   // :84,100
            for (int k = 0; k < D; ++k) std::cout << last_.cost[k] << (k + 1 < D ? "," : "\n");                               // :117
        }
        ++it_count_;                                                                 // :121-126
        if (it_count_ > n_init_samples_) {
            (last_.index_max == 1) ? (dt1_count_++) : (dt0_count_++);
            if (verbose_) {
                const float n = (float)(it_count_ - n_init_samples_);
                std::cout << "it count " << it_count_ - n_init_samples_ << "," << dt0_count_ << "," << dt1_count_ << std::endl;
                std::cout << "freq of index0 and index1 " << (float)dt0_count_ / n << "," << (float)dt1_count_ / n << std::endl;
            }
        }
        msg::Float64 cost;
        cost.data = last_.max_cost;                                                  // :127
        if (pub_cost_) pub_cost_(cost);                                              // :129
        if (verbose_) std::cout << "publish the cost.....................\n \n" << std::endl;
        return cost;
    }

    // A batch of "noise" messages in one launch (candidates x trials x sample times); costs in order.
    std::vector<msg::Float64> ProcessNoiseBatch(const std::vector<msg::Float64MultiArray> &msgs)
    {
        std::vector<double> q, r;
        for (const auto &m : msgs) {
            if (m.layout.dim.empty() || m.layout.dim[0].size < 1 || m.layout.dim[0].size >= m.data.size())
                throw Error(KFBO_EINVAL, "noise message: layout.dim[0].size must split data into [pn..., on...]");
            q.push_back(m.data[0]);                         // pnoise_[0] (system_models.hpp:39)
            r.push_back(m.data[m.layout.dim[0].size]);      // onoise_[0] (system_models.hpp:143)
        }
        last_batch_ = ev_->EvalBatch(q, r);
        std::vector<msg::Float64> out(msgs.size());
        for (size_t i = 0; i < msgs.size(); ++i) out[i].data = last_batch_[i].max_cost;
        it_count_ += (int)msgs.size();
        return out;
    }

    const kfbo_result &last_result() const { return last_; }
    const std::vector<kfbo_result> &last_batch() const { return last_batch_; }
    Evaluator &evaluator() { return *ev_; }
    int it_count() const { return it_count_; }
    int dt0_count() const { return dt0_count_; }
    int dt1_count() const { return dt1_count_; }

private:
    ReadParaVehicle rv_;
    int n_init_samples_;
    Publisher pub_cost_;
    bool verbose_;
    std::unique_ptr<Evaluator> ev_;
    kfbo_result last_{};
    std::vector<kfbo_result> last_batch_;
    int it_count_ = 0, dt0_count_ = 0, dt1_count_ = 0;   // trial_node.cpp:140
};

}  // namespace kfbo
#endif  // KFBO_TRIAL_HPP_
