// ref_driver.cpp -- C entry points around the REFERENCE'S OWN source files (trial.cpp, simulator.cpp,
// kalman_filter.cpp, read_para_vehicle.cpp and the headers under include/ of arpg/kf_bayesopt), compiled from
// where they lie under /root/reference into oracle/_ref/libkfbo_ref.so by oracle/Makefile.  Eigen3, ROS and
// matplotlibcpp are not installed here; oracle/ref_shim/ stands in for the few pieces of them these files use,
// and a counter replaces the wall clock the reference seeds its RNG from.  TEST INFRASTRUCTURE: used by tests/ to
// check the CPU oracle (oracle/kfbo_oracle.c) against the reference's own code on identical inputs and seeds.
//
// Besides the per-class entry points, the reference's NODE is here too: trial_node.cpp is compiled as it lies, its
// main() renamed by the Makefile (-Dmain=kfbo_ref_trial_node_main), on top of the topic / parameter-server
// stand-in of ref_shim/ros/ros.h.  kfbo_ref_node_run queues "noise" messages, runs that main() -- RunTrial's
// constructor, subscribe, spin -> RunTrial::ProcessNoiseCallback (trial_node.cpp:21-132) per message -- and
// returns what the node published on "cost".
#include <atomic>
#include <cstring>
#include <iostream>
#include <sstream>
#include <thread>

#include <std_msgs/Float64.h>
#include <std_msgs/Float64MultiArray.h>

#include "read_para_bayes.h"
#include "trial.h"   // the reference's include/trial.h

// The counter clock: atomic because Trial::Run reads the clock from CPUcoreNumber threads at once.
static std::atomic<long long> g_clock(0), g_calls(0);
extern "C" long long kfbo_ref_clock_next(void)
{
    g_calls.fetch_add(1, std::memory_order_relaxed);
    return g_clock.fetch_add(1, std::memory_order_relaxed);
}

int kfbo_ref_trial_node_main(int nargs, char *args[]);   // trial_node.cpp's main()

namespace {
struct QuietCout {           // the node narrates every evaluation on std::cout
    std::ostringstream sink;
    std::streambuf *old;
    explicit QuietCout(bool on) : old(on ? std::cout.rdbuf(sink.rdbuf()) : nullptr) {}
    ~QuietCout() { if (old) std::cout.rdbuf(old); }
};
}  // namespace

static ReadParaVehicle make_rv(double dt, double t_end, int nsimruns, int cores, double pnoise, double onoise)
{
    ReadParaVehicle rv;                       // the members read_para_vehicle.cpp:119-158 fills from the parameter server
    rv.cpu_core_number_ = cores;
    rv.state_dof_ = 2;
    rv.observation_dof_ = 1;
    rv.t_start_ = 0.0;
    rv.dt_ = dt;
    rv.t_end_ = t_end;
    rv.xi0_ = 0.0;
    rv.xidot0_ = 0.0;
    rv.P0_ = "identity";
    rv.nsimruns_ = nsimruns;
    rv.max_iterations_ = (rv.t_end_ - rv.t_start_) / rv.dt_;   // read_para_vehicle.cpp:145, verbatim
    rv.pnoise_ = {pnoise};
    rv.onoise_ = {onoise};
    rv.dim_pn_ = 1;
    rv.dim_on_ = 1;
    rv.optimization_choice_ = "all";
    rv.cost_choice_ = "JNEES";
    return rv;
}

extern "C" {

// ONE of the CPUcoreNumber Trial objects of trial_node.cpp:55-63: Trial(rv_gt, rv), Run(), the four getters.
// clock0: value of the counter clock at entry; *clock_calls: how many times the reference read the clock.
int kfbo_ref_trial_run(double dt, double t_end, int nsimruns, int cores, double q_truth, double r_truth, double q_est,
                       double r_est, long long clock0, double *out4, long long *clock_calls, int *max_iterations)
{
    ReadParaVehicle rv_gt = make_rv(dt, t_end, nsimruns, cores, q_truth, r_truth);
    ReadParaVehicle rv = make_rv(dt, t_end, nsimruns, cores, q_est, r_est);
    g_clock = clock0;
    g_calls = 0;
    Trial t(rv_gt, rv);
    t.Run();
    out4[0] = t.get_average_nees();
    out4[1] = t.get_average_nis();
    out4[2] = t.get_average_var_nees();
    out4[3] = t.get_average_var_nis();
    if (clock_calls) *clock_calls = g_calls.load();
    if (max_iterations) *max_iterations = rv_gt.max_iterations_;
    return 0;
}

// ProcessModel / ObservationModel as the estimator builds them (noise off): Fd, Qd (row-major 2x2), g*u[count]
// for count < n_u (through Predict on a zero state), H (1x2), R.
int kfbo_ref_models(double pnoise, double onoise, double dt, double *Fd, double *Qd, double *gu /*[n_u][2]*/, int n_u,
                    double *H, double *R)
{
    ReadParaVehicle rv = make_rv(dt, 20.1, 1, 1, pnoise, onoise);
    ProcessModel pm(rv, false);
    ObservationModel om(rv, false);
    const JacobianProcess f = pm.get_f();
    const ProcessNoiseCovariance q = pm.get_qd();
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) { Fd[2 * i + j] = f(i, j); Qd[2 * i + j] = q(i, j); }
    for (int c = 0; c < n_u; ++c) {
        const State x = pm.Predict(State::Zero(), c);
        gu[2 * c] = x(0, 0);
        gu[2 * c + 1] = x(1, 0);
    }
    H[0] = om.get_h()(0, 0);
    H[1] = om.get_h()(0, 1);
    R[0] = om.get_r()(0, 0);
    return 0;
}

// continuousToDiscrete<Matrix2d> (include/continuous2discrete.hpp:8-32) on arbitrary Fc, Qc (row-major).
int kfbo_ref_continuous_to_discrete(const double *Fc, const double *Qc, double dt, double *Fd, double *Qd)
{
    Eigen::Matrix2d fc, qc, fd, qd;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) { fc(i, j) = Fc[2 * i + j]; qc(i, j) = Qc[2 * i + j]; }
    continuousToDiscrete(fd, qd, fc, qc, dt);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) { Fd[2 * i + j] = fd(i, j); Qd[2 * i + j] = qd(i, j); }
    return 0;
}

// KalmanFilter (kalman_filter.cpp:6-38): Initialize(0, I), then Step(z[k], k) for k < T.
// out[k] = {xest0, xest1, P00, P01, P10, P11, nu, S}.
int kfbo_ref_kf_run(double pnoise_est, double onoise_est, double dt, int T, const double *z, double *out)
{
    ReadParaVehicle rv = make_rv(dt, 20.1, 1, 1, pnoise_est, onoise_est);
    KalmanFilter kf(rv);
    kf.Initialize(State::Zero(), StateCovariance::Identity());
    for (int k = 0; k < T; ++k) {
        Observation zk;
        zk(0, 0) = z[k];
        kf.Step(zk, k);
        double *o = out + 8 * k;
        o[0] = kf.get_xest()(0, 0); o[1] = kf.get_xest()(1, 0);
        o[2] = kf.get_pest()(0, 0); o[3] = kf.get_pest()(0, 1);
        o[4] = kf.get_pest()(1, 0); o[5] = kf.get_pest()(1, 1);
        o[6] = kf.get_residual()(0, 0);
        o[7] = kf.get_s()(0, 0);
    }
    return 0;
}

// ---- the parameter server of ref_shim/ros/ros.h (filled by the tests from the reference's own yaml files)
void kfbo_ref_param_clear(void) { ros::ShimMaster::get().params.clear(); }
void kfbo_ref_param_set_int(const char *name, long long v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kInt; p.i = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_ref_param_set_double(const char *name, double v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kDouble; p.d = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_ref_param_set_string(const char *name, const char *v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kString; p.s = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_ref_param_set_vector(const char *name, const double *v, int n)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kDoubleVector; p.v.assign(v, v + n);
    ros::ShimMaster::get().params[name] = p;
}

// ReadParaVehicle(nh, showReadInfo) (read_para_vehicle.cpp:5-158) on the current parameter server.
// nums[14] = {cpu_core_number_, state_dof_, observation_dof_, t_start_, dt_, t_end_, xi0_, xidot0_, nsimruns_,
//             max_iterations_, dim_pn_, dim_on_, pnoise_[0], onoise_[0]};  strs = "P0_\noptimization_choice_\ncost_choice_".
int kfbo_ref_read_para_vehicle(int show_read_info, double *nums, char *strs, int strs_cap)
{
    ros::NodeHandle nh;
    ReadParaVehicle rv(nh, show_read_info != 0);
    const double n[14] = {(double)rv.cpu_core_number_, (double)rv.state_dof_, (double)rv.observation_dof_, rv.t_start_, rv.dt_,
                          rv.t_end_, rv.xi0_, rv.xidot0_, (double)rv.nsimruns_, (double)rv.max_iterations_, (double)rv.dim_pn_,
                          (double)rv.dim_on_, rv.pnoise_.empty() ? 0.0 : rv.pnoise_[0], rv.onoise_.empty() ? 0.0 : rv.onoise_[0]};
    std::memcpy(nums, n, sizeof n);
    const std::string s = rv.P0_ + "\n" + rv.optimization_choice_ + "\n" + rv.cost_choice_;
    if ((int)s.size() + 1 > strs_cap) return -1;
    std::memcpy(strs, s.c_str(), s.size() + 1);
    return 0;
}

// ReadParaBayes(nh) (read_para_bayes.cpp): the members the evaluator side reads (trial_node.cpp:122) and the bounds.
// out[7] = {n_init_samples_, n_iterations_, lower_bound_[0], lower_bound_[1], upper_bound_[0], upper_bound_[1], noise_}
int kfbo_ref_read_para_bayes(double *out)
{
    ros::NodeHandle nh;
    ReadParaBayes rb(nh);
    out[0] = rb.n_init_samples_; out[1] = rb.n_iterations_;
    out[2] = rb.lower_bound_.size() > 0 ? rb.lower_bound_[0] : 0.0; out[3] = rb.lower_bound_.size() > 1 ? rb.lower_bound_[1] : 0.0;
    out[4] = rb.upper_bound_.size() > 0 ? rb.upper_bound_[0] : 0.0; out[5] = rb.upper_bound_.size() > 1 ? rb.upper_bound_[1] : 0.0;
    out[6] = rb.noise_;
    return 0;
}

// THE NODE: n_msgs "noise" messages (message m: dim[0].size = npn[m], dim[1].size = non[m], data = the next
// npn[m] + non[m] doubles of `data`) -> trial_node.cpp's main() -> the "cost" messages it published, in order.
// The vehicle / bayes parameters come from the parameter server (kfbo_ref_param_*).  Returns the number of costs.
int kfbo_ref_node_run(int n_msgs, const int *npn, const int *non, const double *data, long long clock0, int quiet,
                      double *costs, int costs_cap, long long *clock_calls)
{
    ros::ShimMaster &m = ros::ShimMaster::get();
    m.inbox["noise"].clear();
    m.outbox["cost"].clear();
    for (int i = 0; i < n_msgs; ++i) {
        std::shared_ptr<std_msgs::Float64MultiArray> msg = std::make_shared<std_msgs::Float64MultiArray>();
        msg->layout.dim.resize(2);                      // robot1d_bayesopt.cpp:53-63
        msg->layout.dim[0].size = (uint32_t)npn[i];
        msg->layout.dim[1].size = (uint32_t)non[i];
        msg->data.assign(data, data + npn[i] + non[i]);
        data += npn[i] + non[i];
        m.inbox["noise"].push_back(msg);
    }
    g_clock = clock0;
    g_calls = 0;
    {
        QuietCout q(quiet != 0);
        char arg0[] = "robot1d_kf";
        char *args[] = {arg0, nullptr};
        kfbo_ref_trial_node_main(1, args);
    }
    int n = 0;
    for (const std::shared_ptr<void> &p : m.outbox["cost"]) {
        if (n < costs_cap) costs[n] = std::static_pointer_cast<std_msgs::Float64>(p)->data;
        ++n;
    }
    m.outbox["cost"].clear();
    if (clock_calls) *clock_calls = g_calls.load();
    return n;
}

// One sample time of the callback (trial_node.cpp:54-75; src/test_trial.cpp:41-72 is the same code): CPUcoreNumber
// Trial objects, one std::thread per Trial::Run, join, the four getters averaged over the threads.  This is the
// reference's threaded CPU path at an arbitrary (dt, t_end, Nsimruns, CPUcoreNumber) -- the callback itself
// hard-codes its second sample time -- and what bench.py --impl reference times.
int kfbo_ref_eval_threads(double dt, double t_end, int nsimruns, int cores, double q_truth, double r_truth, double q_est,
                          double r_est, long long clock0, double *out4, int *max_iterations)
{
    ReadParaVehicle rv_gt = make_rv(dt, t_end, nsimruns, cores, q_truth, r_truth);
    ReadParaVehicle rv = make_rv(dt, t_end, nsimruns, cores, q_est, r_est);
    g_clock = clock0;
    g_calls = 0;
    std::vector<std::thread> th;
    std::vector<Trial> t;
    for (int i = 0; i < rv_gt.cpu_core_number_; i++) t.push_back(Trial(rv_gt, rv));
    for (int i = 0; i < rv_gt.cpu_core_number_; i++) th.push_back(std::thread(&Trial::Run, &t[i]));
    for (int i = 0; i < rv_gt.cpu_core_number_; i++) th[i].join();
    out4[0] = out4[1] = out4[2] = out4[3] = 0.0;
    for (int i = 0; i < rv_gt.cpu_core_number_; i++) {
        out4[0] += t[i].get_average_nees() / rv_gt.cpu_core_number_;
        out4[1] += t[i].get_average_nis() / rv_gt.cpu_core_number_;
        out4[2] += t[i].get_average_var_nees() / rv_gt.cpu_core_number_;
        out4[3] += t[i].get_average_var_nis() / rv_gt.cpu_core_number_;
    }
    if (max_iterations) *max_iterations = rv_gt.max_iterations_;
    return 0;
}

}  // extern "C"
