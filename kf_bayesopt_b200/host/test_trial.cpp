// test_trial.cpp -- the reference's manual smoke executable (src/test_trial.cpp:17-95,
// launch/robot1d_test_trail.launch:4-5) on the B200 evaluator: one single-sample-time cost
// evaluation with a fixed estimator (pnoise_est, onoise_est), printed the same way.
//
//   kfbo_test_trial [--vehicle config/vehicleParams.yaml] [--pnoise_est 1.0] [--onoise_est 0.1]
//                   [--nsimruns N] [--device D] [--gpus N] [--seed S] [--fused] [--trace out.csv]   (--gpus applies to --fused)
//
// Default path = the reference's caller shape: CPUcoreNumber Trial objects, one std::thread each
// (src/test_trial.cpp:41-51), every Trial::Run() rolling its share out on the GPU.  --fused runs the
// whole batch as ONE launch through the evaluator (what RunTrial does per sample time).
// --trace FILE: the reference's Nsimruns == 1 path (trial.cpp:69-78,116-129 plots the estimation error inside
// its 95 % band with matplotlibcpp): the same per-step record of trajectory 0, written as CSV instead.
#include <cstdlib>
#include <cstring>
#include <thread>

#include "kfbo_trial.hpp"

int main(int nargs, char *args[])
{
    std::string vehicle;
    double pnoise_est = 1.0, onoise_est = 0.1;   // src/test_trial.cpp:23
    int device = -1, nsimruns = -1, gpus = 1;
    uint64_t seed = kfbo::kDefaultSeed;
    bool fused = false;
    std::string trace_path;
    for (int i = 1; i < nargs; ++i) {
        auto next = [&](const char *name) -> const char * {
            if (i + 1 >= nargs) { std::cerr << name << " needs a value\n"; std::exit(64); }
            return args[++i];
        };
        if (!std::strcmp(args[i], "--vehicle")) vehicle = next("--vehicle");
        else if (!std::strcmp(args[i], "--pnoise_est")) pnoise_est = std::atof(next("--pnoise_est"));
        else if (!std::strcmp(args[i], "--onoise_est")) onoise_est = std::atof(next("--onoise_est"));
        else if (!std::strcmp(args[i], "--nsimruns")) nsimruns = std::atoi(next("--nsimruns"));
        else if (!std::strcmp(args[i], "--device")) device = std::atoi(next("--device"));
        else if (!std::strcmp(args[i], "--gpus")) gpus = std::atoi(next("--gpus"));
        else if (!std::strcmp(args[i], "--seed")) seed = std::strtoull(next("--seed"), nullptr, 0);
        else if (!std::strcmp(args[i], "--fused")) fused = true;
        else if (!std::strcmp(args[i], "--trace")) trace_path = next("--trace");
        else { std::cerr << "unknown argument " << args[i] << "\n"; return 64; }
    }
    try {
        kfbo::ReadParaVehicle rv_gt = vehicle.empty() ? kfbo::ReadParaVehicle() : kfbo::ReadParaVehicle::FromYaml(vehicle);
        if (nsimruns > 0) rv_gt.nsimruns_ = nsimruns;
        kfbo::ReadParaVehicle rv = rv_gt;
        rv.pnoise_[0] = pnoise_est;   // src/test_trial.cpp:29-30
        rv.onoise_[0] = onoise_est;
        std::cout << "estimator process noise " << rv.pnoise_[0] << std::endl;
        std::cout << "estimator observation noise " << rv.onoise_[0] << std::endl;
        std::cout << "simulator process noise " << rv_gt.pnoise_[0] << std::endl;
        std::cout << "simulator observation noise " << rv_gt.onoise_[0] << std::endl;

        double average_nees = 0.0, average_var_nees = 0.0, average_nis = 0.0, average_var_nis = 0.0;
        if (!fused) {
            std::vector<std::thread> th;
            std::vector<kfbo::Trial> t;
            for (int i = 0; i < rv_gt.cpu_core_number_; i++) t.push_back(kfbo::Trial(rv_gt, rv, device, seed));
            for (int i = 0; i < rv_gt.cpu_core_number_; i++) th.push_back(std::thread(&kfbo::Trial::Run, &t[i]));
            for (int i = 0; i < rv_gt.cpu_core_number_; i++) th[i].join();
            for (int i = 0; i < rv_gt.cpu_core_number_; i++)
                if (!t[i].ok()) throw std::runtime_error("Trial " + std::to_string(i) + ": " + t[i].error());
            for (int i = 0; i < rv_gt.cpu_core_number_; i++) {   // src/test_trial.cpp:70-72,82-84
                average_nees += t[i].get_average_nees() / rv_gt.cpu_core_number_;
                average_var_nees += t[i].get_average_var_nees() / rv_gt.cpu_core_number_;
                average_nis += t[i].get_average_nis() / rv_gt.cpu_core_number_;
                average_var_nis += t[i].get_average_var_nis() / rv_gt.cpu_core_number_;
            }
        } else {
            kfbo::Evaluator ev(kfbo::MakeParams(rv_gt, kfbo::SampleTimesFrom(rv_gt, nullptr), "JNEES", seed, 1, device), gpus);
            const kfbo_result r = ev.Eval(rv.pnoise_, rv.onoise_);
            average_nees = r.average_nees[0]; average_var_nees = r.average_var_nees[0];
            average_nis = r.average_nis[0]; average_var_nis = r.average_var_nis[0];
            int launches = 0;
            const double ms = ev.LastKernelMs(&launches);
            std::cout << "rollout kernel " << ms << " ms, " << launches << " launch(es)" << std::endl;
        }
        if (!trace_path.empty()) {
            kfbo::Evaluator ev(kfbo::MakeParams(rv_gt, kfbo::SampleTimesFrom(rv_gt, nullptr), "JNEES", seed, 1, device));
            const std::vector<double> tr = ev.TraceTrial(0, rv.pnoise_[0], rv.onoise_[0], 0u, 0u);
            std::ofstream f(trace_path);
            f.precision(17);
            f << "time,e1,e2,lb1,ub1,lb2,ub2,nees,nis\n";
            for (size_t s = 0; s < tr.size() / KFBO_TRACE_COLUMNS; ++s)
                for (int c = 0; c < KFBO_TRACE_COLUMNS; ++c) f << tr[s * KFBO_TRACE_COLUMNS + c] << (c + 1 < KFBO_TRACE_COLUMNS ? "," : "\n");
            std::cout << "trace of trajectory 0 written to " << trace_path << " (" << tr.size() / KFBO_TRACE_COLUMNS << " steps)" << std::endl;
        }
        if (rv.cost_choice_ == "JNEES" || rv.cost_choice_ == "CNEES") {
            const double J_NEES = std::abs(std::log(average_nees / rv.state_dof_));
            std::cout << "cost JNEES " << J_NEES << std::endl;
            std::cout << "variance NEES " << average_var_nees << std::endl;
        } else {
            const double J_NIS = std::abs(std::log(average_nis / rv.observation_dof_));
            const double tmp_NIS_var = std::abs(std::log(0.5 * average_var_nis / rv.observation_dof_));
            std::cout << "average NIS " << average_nis << " and its cost JNIS " << J_NIS << std::endl;
            std::cout << "average variance" << average_var_nis << ", and its log value " << tmp_NIS_var << std::endl;
        }
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
