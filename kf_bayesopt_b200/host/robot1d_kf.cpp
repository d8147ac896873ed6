// robot1d_kf.cpp -- the cost-evaluation node (trial_node.cpp:144-157) on the B200 evaluator.
//
// With -DKFBO_WITH_ROS (roscpp + std_msgs available) this is the ROS node "robot_1d_kf": it
// subscribes "noise" (queue 10), publishes "cost" (queue 10), parameters from the yaml path in
// ~vehicle_yaml -- the bayesopt node and both yaml files stay unchanged.
//
// Without ROS (this container) the same RunTrial is driven over stdin/stdout, one message per line:
//     noise  <npn> <non> <pn...> <on...>         ->  cost <max_cost> <index_max> <pic_max...>
//     batch  <C> <q_0> <r_0> ... <q_C-1> <r_C-1> ->  C lines "cost ..."
//   kfbo_robot1d_kf [--vehicle config/vehicleParams.yaml] [--second code|readme|none] [--n_init_samples 20]
//                   [--max_candidates C] [--device D] [--gpus N] [--seed S] [--quiet]   (--gpus N: the node drives N GPUs)
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <limits>
#include <sstream>

#include "kfbo_trial.hpp"

#ifdef KFBO_WITH_ROS
#include <ros/ros.h>
#endif

int main(int nargs, char *args[])
{
    std::string vehicle, second = "code";
    int device = -1, n_init = 20, max_cand = 1, gpus = 1;
    uint64_t seed = kfbo::kDefaultSeed;
    bool verbose = true;
#ifdef KFBO_WITH_ROS
    ros::init(nargs, args, "robot_1d_kf");   // trial_node.cpp:146
    ros::NodeHandle n;
    ros::param::get("~vehicle_yaml", vehicle);
#endif
    for (int i = 1; i < nargs; ++i) {
        auto next = [&](const char *name) -> const char * {
            if (i + 1 >= nargs) { std::cerr << name << " needs a value\n"; std::exit(64); }
            return args[++i];
        };
        if (!std::strcmp(args[i], "--vehicle")) vehicle = next("--vehicle");
        else if (!std::strcmp(args[i], "--second")) second = next("--second");
        else if (!std::strcmp(args[i], "--n_init_samples")) n_init = std::atoi(next("--n_init_samples"));
        else if (!std::strcmp(args[i], "--max_candidates")) max_cand = std::atoi(next("--max_candidates"));
        else if (!std::strcmp(args[i], "--device")) device = std::atoi(next("--device"));
        else if (!std::strcmp(args[i], "--gpus")) gpus = std::atoi(next("--gpus"));
        else if (!std::strcmp(args[i], "--seed")) seed = std::strtoull(next("--seed"), nullptr, 0);
        else if (!std::strcmp(args[i], "--quiet")) verbose = false;
        else { std::cerr << "unknown argument " << args[i] << "\n"; return 64; }
    }
    try {
        const kfbo::ReadParaVehicle rv = vehicle.empty() ? kfbo::ReadParaVehicle() : kfbo::ReadParaVehicle::FromYaml(vehicle);
        const kfbo::SecondSampleTime *sec = second == "code" ? &kfbo::kCodeSecondSampleTime
                                          : second == "readme" ? &kfbo::kReadmeSecondSampleTime : nullptr;
#ifdef KFBO_WITH_ROS
        ros::Publisher pub_cost = n.advertise<std_msgs::Float64>("cost", 10);                     // trial_node.cpp:17
        kfbo::RunTrial rT(rv, sec, n_init, [&](const std_msgs::Float64 &c) { pub_cost.publish(c); }, verbose, device, seed, max_cand, gpus);
        boost::function<void(const std_msgs::Float64MultiArray::ConstPtr &)> cb =
            [&](const std_msgs::Float64MultiArray::ConstPtr &m) {
                // a malformed message or a failed evaluation must not escape ros::spin() and kill the node: log it and
                // answer with a NaN cost, which ends the optimiser's wait loop (robot1d_bayesopt.cpp:107)
                try {
                    rT.ProcessNoiseCallback(*m);
                } catch (const std::exception &e) {
                    ROS_ERROR("robot_1d_kf: noise message dropped: %s", e.what());
                    std_msgs::Float64 c;
                    c.data = std::numeric_limits<double>::quiet_NaN();
                    pub_cost.publish(c);
                }
            };
        ros::Subscriber sub = n.subscribe<std_msgs::Float64MultiArray>("noise", 10, cb);           // trial_node.cpp:150
        (void)sub;              // kept alive until spin() returns
        ros::spin();
#else
        kfbo::RunTrial rT(rv, sec, n_init, nullptr, verbose, device, seed, max_cand, gpus);
        std::cout << std::setprecision(17);
        std::string line;
        while (std::getline(std::cin, line)) {
            std::stringstream ss(line);
            std::string kind;
            if (!(ss >> kind) || kind[0] == '#') continue;
            if (kind == "noise") {
                size_t npn = 0, non = 0;
                ss >> npn >> non;
                std::vector<double> pn(npn), on(non);
                for (double &v : pn) ss >> v;
                for (double &v : on) ss >> v;
                if (!ss || npn < 1 || non < 1) { std::cout << "error malformed noise line" << std::endl; continue; }
                rT.ProcessNoiseCallback(kfbo::msg::MakeNoise(pn, on));
                const kfbo_result &r = rT.last_result();
                std::cout << "cost " << r.max_cost << " " << r.index_max;
                for (int k = 0; k < rT.evaluator().params().n_sample_times; ++k) std::cout << " " << r.cost[k];
                std::cout << std::endl;
            } else if (kind == "batch") {
                size_t C = 0;
                ss >> C;
                std::vector<kfbo::msg::Float64MultiArray> msgs;
                for (size_t c = 0; c < C; ++c) {
                    double q, r;
                    ss >> q >> r;
                    msgs.push_back(kfbo::msg::MakeNoise({q}, {r}));
                }
                if (!ss || C < 1) { std::cout << "error malformed batch line" << std::endl; continue; }
                rT.ProcessNoiseBatch(msgs);
                for (const kfbo_result &r : rT.last_batch()) {
                    std::cout << "cost " << r.max_cost << " " << r.index_max;
                    for (int k = 0; k < rT.evaluator().params().n_sample_times; ++k) std::cout << " " << r.cost[k];
                    std::cout << std::endl;
                }
            } else {
                std::cout << "error unknown message kind '" << kind << "'" << std::endl;
            }
        }
#endif
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
