// bayesopt/bayesopt.hpp -- STAND-IN for the external bayesopt library (zhaozhongch/bayesopt, a fork of rmcantin/bayesopt, linked
// by the reference as /usr/local/lib/libbayesopt.a -- CMakeLists.txt:158; not installed here, no network) so that the reference's
// UNMODIFIED optimiser node robot1d_bayesopt.cpp compiles and runs.  It restates the INTERFACE that file uses --
// bayesopt::Parameters, initialize_parameters_to_default(), ContinuousModel(dim, par), setBoundingBox, optimize, the virtual
// evaluateSample / checkReachability, the public mParameters, the plotting getters of the fork -- and NOT the library's
// algorithms: optimize() runs this repository's small GP / expected-improvement optimiser (kf_bayesopt_b200/host/gp_ei.hpp)
// with the same budget (n_init_samples Latin-hypercube points, then n_iterations sequential evaluations).  Optimiser parity is
// therefore unpinned; what the run pins is everything AROUND the optimiser, in the reference's own code:
// the "noise" message per optimizationChoice, the wait for "cost", the parameter files.  Test infrastructure.
#pragma once
#include <string>
#include <vector>

#include "../specialtypes.hpp"
#include "gp_ei.hpp"        // kf_bayesopt_b200/host

namespace bayesopt {

struct Parameters {
    size_t n_iterations = 190;
    size_t n_inner_iterations = 500;
    size_t n_init_samples = 10;
    size_t n_iter_relearn = 50;
    int verbose_level = 1;
    int random_seed = -1;
    std::string surr_name = "sGaussianProcess";
    double noise = 1e-6;
    int force_jump = 20;
};

class ContinuousModel {
public:
    ContinuousModel(size_t dim, Parameters par) : mParameters(par), dim_(dim), lower_(dim, 0.0), upper_(dim, 1.0) {}
    virtual ~ContinuousModel() {}
    virtual double evaluateSample(const vectord &query) = 0;
    virtual bool checkReachability(const vectord & /*query*/) { return true; }

    void setBoundingBox(const vectord &lower, const vectord &upper)
    {
        lower_.assign(lower.begin(), lower.end());
        upper_.assign(upper.begin(), upper.end());
    }

    void optimize(vectord &result)
    {
        kfbo::bo::ReadParaBayes rb;
        rb.n_init_samples_ = (int)mParameters.n_init_samples;
        rb.n_iterations_ = (int)mParameters.n_iterations;
        rb.random_seed_ = mParameters.random_seed;
        rb.verbose_level_ = mParameters.verbose_level;
        rb.noise_ = mParameters.noise;
        rb.lower_bound_ = lower_;
        rb.upper_bound_ = upper_;
        kfbo::bo::GpEi opt(rb, mParameters.random_seed >= 0 ? (uint64_t)mParameters.random_seed : 0x4B46424Full);
        const std::vector<double> best = opt.Optimize([&](const std::vector<double> &x) {
            vectord q(x.size());
            for (size_t i = 0; i < x.size(); ++i) q(i) = x[i];
            const double y = evaluateSample(q);
            sample_x_.insert(sample_x_.end(), x.begin(), x.end());
            sample_y_.push_back(y);
            return y;
        });
        for (size_t i = 0; i < best.size() && i < result.size(); ++i) result(i) = best[i];
    }

    // The fork's getters for the surrogate / acquisition plot (robot1d_bayesopt.cpp:126-218 publishes them on "sgAc" for
    // scripts/plot_surrogate_acquisition.py).  The stand-in has no plot grid: the curves are empty, the samples are real.
    std::vector<double> get1Dx() { return {}; }
    std::vector<double> get1Dmean() { return {}; }
    std::vector<double> get1DStdLowerBound() { return {}; }
    std::vector<double> get1DStdUpperBound() { return {}; }
    std::vector<double> get1DAcquisitionFunctionValue() { return {}; }
    std::vector<double> get2Dx() { return {}; }
    std::vector<std::vector<double>> get2Dmean() { return {}; }
    std::vector<std::vector<double>> get2DStdLowerBound() { return {}; }
    std::vector<std::vector<double>> get2DStdUpperBound() { return {}; }
    std::vector<std::vector<double>> get2DAcquisitionFunctionValue() { return {}; }
    std::vector<double> getSampleX() { return sample_x_; }
    std::vector<double> getSampleY() { return sample_y_; }

    Parameters mParameters;        // "mParameter is a public member of BayesoptBase class" (robot1d_bayesopt.cpp:161)

private:
    size_t dim_;
    std::vector<double> lower_, upper_, sample_x_, sample_y_;
};

}  // namespace bayesopt

inline bayesopt::Parameters initialize_parameters_to_default() { return bayesopt::Parameters(); }
