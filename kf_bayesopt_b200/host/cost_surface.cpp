// cost_surface.cpp -- the cost over a grid of candidate noise settings in ONE batched launch
// (BASELINE.json configs[3]), written as CSV for plotting: the evaluator-side counterpart of what
// scripts/plot_surrogate_acquisition.py shows of the optimiser's surrogate.
//
//   kfbo_cost_surface [--vehicle vehicleParams.yaml] [--bayes bayesParams.yaml] [--grid 64] [--nsimruns N]
//                     [--second code|readme|none] [--linear] [--out surface.csv] [--device D] [--gpus N] [--seed S]
// Columns: pnoise, onoise, cost per sample time..., max_cost, index_max.  Bounds from bayesParams.yaml:18-19.
#include <cstdlib>
#include <cstring>

#include "gp_ei.hpp"
#include "kfbo_trial.hpp"

int main(int nargs, char *args[])
{
    std::string vehicle, bayes, second = "code", out_path = "surface.csv";
    int grid = 64, nsimruns = -1, device = -1, gpus = 1;
    bool linear = false;
    uint64_t seed = kfbo::kDefaultSeed;
    for (int i = 1; i < nargs; ++i) {
        auto next = [&](const char *name) -> const char * {
            if (i + 1 >= nargs) { std::cerr << name << " needs a value\n"; std::exit(64); }
            return args[++i];
        };
        if (!std::strcmp(args[i], "--vehicle")) vehicle = next("--vehicle");
        else if (!std::strcmp(args[i], "--bayes")) bayes = next("--bayes");
        else if (!std::strcmp(args[i], "--grid")) grid = std::atoi(next("--grid"));
        else if (!std::strcmp(args[i], "--nsimruns")) nsimruns = std::atoi(next("--nsimruns"));
        else if (!std::strcmp(args[i], "--second")) second = next("--second");
        else if (!std::strcmp(args[i], "--linear")) linear = true;
        else if (!std::strcmp(args[i], "--out")) out_path = next("--out");
        else if (!std::strcmp(args[i], "--device")) device = std::atoi(next("--device"));
        else if (!std::strcmp(args[i], "--gpus")) gpus = std::atoi(next("--gpus"));
        else if (!std::strcmp(args[i], "--seed")) seed = std::strtoull(next("--seed"), nullptr, 0);
        else { std::cerr << "unknown argument " << args[i] << "\n"; return 64; }
    }
    try {
        if (grid < 1 || grid > 1024) throw kfbo::Error(KFBO_EINVAL, "--grid must be in [1, 1024]");
        kfbo::ReadParaVehicle rv = vehicle.empty() ? kfbo::ReadParaVehicle() : kfbo::ReadParaVehicle::FromYaml(vehicle);
        if (nsimruns > 0) rv.nsimruns_ = nsimruns;
        const kfbo::bo::ReadParaBayes rb = bayes.empty() ? kfbo::bo::ReadParaBayes() : kfbo::bo::ReadParaBayes::FromYaml(bayes);
        if (rb.lower_bound_.size() < 2) throw kfbo::Error(KFBO_EINVAL, "the surface needs bounds for (pnoise, onoise)");
        const kfbo::SecondSampleTime *sec = second == "code" ? &kfbo::kCodeSecondSampleTime
                                          : second == "readme" ? &kfbo::kReadmeSecondSampleTime : nullptr;
        auto axis = [&](int d, int i) {
            const double lo = rb.lower_bound_[d], hi = rb.upper_bound_[d], t = grid > 1 ? (double)i / (grid - 1) : 0.5;
            return linear ? lo + t * (hi - lo) : lo * std::pow(hi / lo, t);
        };
        std::vector<double> q, r;
        for (int i = 0; i < grid; ++i)
            for (int j = 0; j < grid; ++j) { q.push_back(axis(0, i)); r.push_back(axis(1, j)); }
        kfbo::Evaluator ev(kfbo::MakeParams(rv, kfbo::SampleTimesFrom(rv, sec), rv.cost_choice_, seed, grid * grid, device), gpus, KFBO_SHARD_CANDIDATES);
        const kfbo::Evaluator::BatchCosts res = ev.EvalBatchCosts(q, r);     // costs assembled on the device
        const int D = ev.params().n_sample_times;
        std::ofstream f(out_path);
        if (!f) throw kfbo::Error(KFBO_EINVAL, "cannot write " + out_path);
        f.precision(17);
        f << "pnoise,onoise";
        for (int k = 0; k < D; ++k) f << ",cost_dt" << ev.params().dt[k];
        f << ",max_cost,index_max\n";
        size_t best = 0;
        for (size_t c = 0; c < q.size(); ++c) {
            f << q[c] << "," << r[c];
            for (int k = 0; k < D; ++k) f << "," << res.cost[c * D + k];
            f << "," << res.max_cost[c] << "," << res.index_max[c] << "\n";
            if (res.max_cost[c] < res.max_cost[best]) best = c;
        }
        int launches = 0;
        const double ms = ev.LastKernelMs(&launches);
        std::cout << grid * grid << " candidates x " << kfbo_effective_trials(&ev.params()) << " trials x " << D << " sample times: "
                  << ms << " ms in " << launches << " launch(es); minimum " << rv.cost_choice_ << " " << res.max_cost[best]
                  << " at pnoise " << q[best] << ", onoise " << r[best] << "; written to " << out_path << std::endl;
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
