// bo_loop.cpp -- BASELINE.json configs[4]: the full Bayesian-optimisation loop driving the GPU cost
// evaluator over the "noise" / "cost" topics, end to end, without ROS or the bayesopt library (neither
// is installed here).  Two nodes on an in-process topic bus:
//   * the evaluator node: kfbo::RunTrial subscribed to "noise", publishing "cost"  (trial_node.cpp:144-157)
//   * the optimiser node: Robot1D::evaluateSample's contract (robot1d_bayesopt.cpp:47-113) -- build the
//     Float64MultiArray for the configured optimizationChoice, publish it, spin until a cost arrives --
//     driven by the GP/EI stand-in of gp_ei.hpp with the settings of config/bayesParams.yaml.
// --reference-sleeps adds the reference's fixed usleep(300 ms) + usleep(200 ms) per iteration
// (robot1d_bayesopt.cpp:91,96).  Prints one JSON line with wall time and the time spent inside the
// evaluator callbacks.
//
//   kfbo_bo_loop [--vehicle vehicleParams.yaml] [--bayes bayesParams.yaml] [--iterations N]
//                [--init N] [--nsimruns N] [--second code|readme|none] [--reference-sleeps]
//                [--device D] [--gpus N] [--seed S] [--quiet]     (--gpus N: this one process drives N GPUs)
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iomanip>

#include "gp_ei.hpp"
#include "kfbo_trial.hpp"
#include "topic_bus.hpp"

namespace {

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// The optimiser-side node: robot1d_bayesopt.cpp:28-118 without the bayesopt base class.
class Robot1D {
public:
    Robot1D(kfbo::TopicBus &bus, const kfbo::ReadParaVehicle &rv, bool reference_sleeps) : bus_(bus), rv_(rv), sleeps_(reference_sleeps)
    {
        bus_.Subscribe<kfbo::msg::Float64>("cost", 10, [this](const kfbo::msg::Float64 &m) { cost_ = m.data; });   // :39,115-118
    }
    double evaluateSample(const std::vector<double> &xin)
    {
        cost_ = -1;                                                            // :51
        ++it_counter_;
        std::vector<double> pn, on;
        if (rv_.optimization_choice_ == "processNoise") { pn = xin; on = rv_.onoise_; }                 // :68-72
        else if (rv_.optimization_choice_ == "observationNoise") { pn = rv_.pnoise_; on = xin; }       // :73-77
        else if (rv_.optimization_choice_ == "all") {                                                   // :78-80
            pn.assign(xin.begin(), xin.begin() + rv_.pnoise_.size());
            on.assign(xin.begin() + rv_.pnoise_.size(), xin.end());
        } else throw kfbo::Error(KFBO_EINVAL, "Please choose optimization method!!!");                  // :82
        if (sleeps_) usleep(300000);                                           // :91
        bus_.Publish("noise", kfbo::msg::MakeNoise(pn, on));                   // :92
        if (sleeps_) usleep(200000);                                           // :96
        while (cost_ == -1) bus_.SpinOnce();                                   // :107-109
        if (cost_ < min_cost_) min_cost_ = cost_;
        return cost_;
    }
    int it_counter() const { return it_counter_; }
    double min_cost() const { return min_cost_; }
private:
    kfbo::TopicBus &bus_;
    kfbo::ReadParaVehicle rv_;
    bool sleeps_;
    double cost_ = -1, min_cost_ = 1e300;
    int it_counter_ = 0;
};

}  // namespace

int main(int nargs, char *args[])
{
    std::string vehicle, bayes, second = "code";
    int iterations = -1, init = -1, nsimruns = -1, device = -1, gpus = 1;
    uint64_t seed = kfbo::kDefaultSeed;
    bool sleeps = false, verbose = true;
    for (int i = 1; i < nargs; ++i) {
        auto next = [&](const char *name) -> const char * {
            if (i + 1 >= nargs) { std::cerr << name << " needs a value\n"; std::exit(64); }
            return args[++i];
        };
        if (!std::strcmp(args[i], "--vehicle")) vehicle = next("--vehicle");
        else if (!std::strcmp(args[i], "--bayes")) bayes = next("--bayes");
        else if (!std::strcmp(args[i], "--iterations")) iterations = std::atoi(next("--iterations"));
        else if (!std::strcmp(args[i], "--init")) init = std::atoi(next("--init"));
        else if (!std::strcmp(args[i], "--nsimruns")) nsimruns = std::atoi(next("--nsimruns"));
        else if (!std::strcmp(args[i], "--second")) second = next("--second");
        else if (!std::strcmp(args[i], "--reference-sleeps")) sleeps = true;
        else if (!std::strcmp(args[i], "--device")) device = std::atoi(next("--device"));
        else if (!std::strcmp(args[i], "--gpus")) gpus = std::atoi(next("--gpus"));
        else if (!std::strcmp(args[i], "--seed")) seed = std::strtoull(next("--seed"), nullptr, 0);
        else if (!std::strcmp(args[i], "--quiet")) verbose = false;
        else { std::cerr << "unknown argument " << args[i] << "\n"; return 64; }
    }
    try {
        kfbo::ReadParaVehicle rv = vehicle.empty() ? kfbo::ReadParaVehicle() : kfbo::ReadParaVehicle::FromYaml(vehicle);
        kfbo::bo::ReadParaBayes rb = bayes.empty() ? kfbo::bo::ReadParaBayes() : kfbo::bo::ReadParaBayes::FromYaml(bayes);
        if (iterations >= 0) rb.n_iterations_ = iterations;
        if (init >= 0) rb.n_init_samples_ = init;
        if (nsimruns > 0) rb.n_init_samples_ = rb.n_init_samples_, rv.nsimruns_ = nsimruns;
        // the optimisation dimension follows optimizationChoice like the reference's bounds do (bayesParams.yaml:18-19)
        if (rv.optimization_choice_ == "processNoise") { rb.lower_bound_.resize(rv.pnoise_.size()); rb.upper_bound_.resize(rv.pnoise_.size()); }
        else if (rv.optimization_choice_ == "observationNoise") {
            rb.lower_bound_.erase(rb.lower_bound_.begin(), rb.lower_bound_.end() - (long)rv.onoise_.size());
            rb.upper_bound_.erase(rb.upper_bound_.begin(), rb.upper_bound_.end() - (long)rv.onoise_.size());
        }
        const kfbo::SecondSampleTime *sec = second == "code" ? &kfbo::kCodeSecondSampleTime
                                          : second == "readme" ? &kfbo::kReadmeSecondSampleTime : nullptr;

        kfbo::TopicBus bus;
        double eval_s = 0.0;
        int evals = 0;
        kfbo::RunTrial rT(rv, sec, rb.n_init_samples_, [&](const kfbo::msg::Float64 &c) { bus.Publish("cost", c); }, false, device, seed, 1, gpus);
        bus.Subscribe<kfbo::msg::Float64MultiArray>("noise", 10, [&](const kfbo::msg::Float64MultiArray &m) {
            const double t0 = now_s();
            rT.ProcessNoiseCallback(m);
            eval_s += now_s() - t0;
            ++evals;
        });
        Robot1D sc(bus, rv, sleeps);
        kfbo::bo::GpEi opt(rb, rb.random_seed_ >= 0 ? (uint64_t)rb.random_seed_ : seed);

        const double t0 = now_s();
        const std::vector<double> result = opt.Optimize([&](const std::vector<double> &x) {
            const double c = sc.evaluateSample(x);
            if (verbose) {
                std::cout << "iteration " << sc.it_counter() << " x";
                for (double v : x) std::cout << " " << v;
                std::cout << " cost " << c << " (min " << sc.min_cost() << ")" << std::endl;
            }
            return c;
        });
        const double wall = now_s() - t0;

        std::cout << "final optimization result ";                              // robot1d_bayesopt.cpp:266-269
        for (double r : result) std::cout << r << ",";
        std::cout << std::endl;
        std::cout << std::setprecision(10) << "{\"workload\": \"bo_loop\", \"iterations\": " << sc.it_counter()
                  << ", \"n_init_samples\": " << rb.n_init_samples_ << ", \"n_iterations\": " << rb.n_iterations_
                  << ", \"lower_bound\": [" << rb.lower_bound_[0] << (rb.lower_bound_.size() > 1 ? ", " + std::to_string(rb.lower_bound_[1]) : "")
                  << "], \"upper_bound\": [" << rb.upper_bound_[0] << (rb.upper_bound_.size() > 1 ? ", " + std::to_string(rb.upper_bound_[1]) : "")
                  << "], \"optimization_choice\": \"" << rv.optimization_choice_ << "\", \"nsimruns\": " << rv.nsimruns_
                  << ", \"sample_times\": " << rT.evaluator().params().n_sample_times << ", \"reference_sleeps\": " << (sleeps ? "true" : "false")
                  << ", \"wall_s\": " << wall << ", \"evaluator_s\": " << eval_s << ", \"evaluator_ms_per_iteration\": " << eval_s / std::max(evals, 1) * 1e3
                  << ", \"optimiser_s\": " << wall - eval_s - (sleeps ? 0.5 * sc.it_counter() : 0.0) << ", \"best_cost\": " << opt.best_value() << ", \"best_x\": [";
        for (size_t i = 0; i < result.size(); ++i) std::cout << (i ? ", " : "") << result[i];
        std::cout << "], \"dt0_count\": " << rT.dt0_count() << ", \"dt1_count\": " << rT.dt1_count() << "}" << std::endl;
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
