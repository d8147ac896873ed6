// ref_driver.cpp -- C entry points around the REFERENCE'S OWN source files (trial.cpp, simulator.cpp,
// kalman_filter.cpp, read_para_vehicle.cpp and the headers under include/ of arpg/kf_bayesopt), compiled from
// where they lie under /root/reference into oracle/_ref/libkfbo_ref.so by oracle/Makefile.  Eigen3, ROS and
// matplotlibcpp are not installed here; oracle/ref_shim/ stands in for the few pieces of them these files use,
// and a counter replaces the wall clock the reference seeds its RNG from.  TEST INFRASTRUCTURE: used by tests/ to
// check the CPU oracle (oracle/kfbo_oracle.c) against the reference's own code on identical inputs and seeds.
#include <cstring>

#include "trial.h"   // the reference's include/trial.h

static long long g_clock = 0, g_calls = 0;
extern "C" long long kfbo_ref_clock_next(void)
{
    ++g_calls;
    return g_clock++;
}

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
    if (clock_calls) *clock_calls = g_calls;
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

}  // extern "C"
