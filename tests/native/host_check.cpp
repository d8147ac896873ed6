// host_check.cpp -- CPU-only checks of the C++ host layer (no GPU calls): the flat-yaml reader on
// files shaped like the reference's config/*.yaml, ReadParaVehicle's derived fields, the "noise"
// message builder, the in-process topic bus, and the GP/EI optimiser on analytic functions.
// Prints "ok <name>" per check; exits non-zero on the first failure.
#include <cstdio>
#include <fstream>

#include "gp_ei.hpp"
#include "kfbo_trial.hpp"
#include "topic_bus.hpp"

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

int main(int argc, char **argv)
{
    CHECK(argc == 3);
    const std::string vehicle = argv[1], bayes = argv[2];
    {   // ReadParaVehicle from the yaml (read_para_vehicle.cpp:119-158)
        kfbo::ReadParaVehicle rv = kfbo::ReadParaVehicle::FromYaml(vehicle);
        CHECK(rv.cpu_core_number_ == 10 && rv.state_dof_ == 2 && rv.observation_dof_ == 1);
        CHECK(rv.dt_ == 0.1 && rv.t_end_ == 20.1 && rv.t_start_ == 0.0);
        CHECK(rv.max_iterations_ == 201);   // truncating (20.1-0)/0.1 = 201.00000000000003
        CHECK(rv.nsimruns_ == 200 && rv.pnoise_.size() == 1 && rv.pnoise_[0] == 1.0 && rv.onoise_[0] == 0.1);
        CHECK(rv.optimization_choice_ == "all" && rv.cost_choice_ == "JNEES");
        auto st = kfbo::SampleTimesFrom(rv, &kfbo::kCodeSecondSampleTime);
        CHECK(st.size() == 2 && st[0].second == 201 && st[1].first == 1.0 && st[1].second == 201);
        st = kfbo::SampleTimesFrom(rv, &kfbo::kReadmeSecondSampleTime);
        CHECK(st[1].first == 0.5 && st[1].second == 211);
        const kfbo_params p = kfbo::MakeParams(rv, st, rv.cost_choice_, 7, 3, -1);
        CHECK(p.n_sample_times == 2 && p.max_iterations[1] == 211 && p.nsimruns == 200 && p.cpu_core_number == 10);
        CHECK(p.pnoise_truth == 1.0 && p.onoise_truth == 0.1 && p.cost_choice == KFBO_JNEES && p.max_candidates == 3);
        CHECK(kfbo_effective_trials(&p) == 200);
        std::printf("ok read_para_vehicle\n");
    }
    {   // ReadParaBayes
        kfbo::bo::ReadParaBayes rb = kfbo::bo::ReadParaBayes::FromYaml(bayes);
        CHECK(rb.n_init_samples_ == 20 && rb.n_iterations_ == 100 && rb.random_seed_ == -1);
        CHECK(rb.lower_bound_.size() == 2 && rb.lower_bound_[0] == 0.1 && rb.lower_bound_[1] == 0.01);
        CHECK(rb.upper_bound_[0] == 5.0 && rb.upper_bound_[1] == 1.0 && rb.noise_ == 1e-8);
        CHECK(rb.n_iter_relearn_ == 10 && rb.n_inner_evaluation_ == 500);
        std::printf("ok read_para_bayes\n");
    }
    {   // the "noise" message (robot1d_bayesopt.cpp:53-88)
        const auto m = kfbo::msg::MakeNoise({2.5}, {0.3});
        CHECK(m.layout.dim.size() == 2 && m.layout.dim[0].size == 1 && m.layout.dim[1].size == 1);
        CHECK(m.layout.dim[0].label == "processNoiseDimension" && m.layout.dim[1].label == "measurmentNoiseDimension");
        CHECK(m.data.size() == 2 && m.data[0] == 2.5 && m.data[1] == 0.3);
        std::printf("ok noise_message\n");
    }
    {   // topic bus: queue depth, ordering, callbacks that publish
        kfbo::TopicBus bus;
        std::vector<double> got;
        int replies = 0;
        bus.Subscribe<kfbo::msg::Float64>("a", 2, [&](const kfbo::msg::Float64 &m) {
            got.push_back(m.data);
            kfbo::msg::Float64 r; r.data = -m.data;
            bus.Publish("b", r);
        });
        bus.Subscribe<kfbo::msg::Float64>("b", 10, [&](const kfbo::msg::Float64 &) { ++replies; });
        for (int i = 1; i <= 3; ++i) { kfbo::msg::Float64 m; m.data = i; bus.Publish("a", m); }
        CHECK(bus.SpinOnce() == 2);                       // queue depth 2: the oldest message was dropped
        CHECK(got.size() == 2 && got[0] == 2 && got[1] == 3 && replies == 0);
        CHECK(bus.SpinOnce() == 2 && replies == 2);       // the replies run on the next spin
        CHECK(bus.SpinOnce() == 0);
        std::printf("ok topic_bus\n");
    }
    {   // GP/EI: a smooth 2-D bowl in the reference's search box, and a 1-D |log| cost shaped like J_NEES(q)
        kfbo::bo::ReadParaBayes rb;
        rb.n_init_samples_ = 10; rb.n_iterations_ = 40; rb.n_inner_evaluation_ = 200;
        kfbo::bo::GpEi opt(rb, 1234);
        const auto f = [](const std::vector<double> &x) { return (x[0] - 1.0) * (x[0] - 1.0) + 30.0 * (x[1] - 0.1) * (x[1] - 0.1); };
        const auto best = opt.Optimize(f);
        CHECK(opt.evaluations() == 50);
        CHECK(opt.best_value() < 0.02);
        CHECK(std::abs(best[0] - 1.0) < 0.25 && std::abs(best[1] - 0.1) < 0.05);
        double mu, sd;
        opt.Predict(best, &mu, &sd);
        CHECK(std::abs(mu - opt.best_value()) < 0.05 && sd < 0.1);
        kfbo::bo::ReadParaBayes rb1 = rb;
        rb1.lower_bound_ = {0.1}; rb1.upper_bound_ = {5.0};
        kfbo::bo::GpEi opt1(rb1, 99);
        const auto b1 = opt1.Optimize([](const std::vector<double> &x) { return std::abs(std::log(x[0] / 1.0)); });
        CHECK(std::abs(b1[0] - 1.0) < 0.15);
        std::printf("ok gp_ei\n");
    }
    return 0;
}
