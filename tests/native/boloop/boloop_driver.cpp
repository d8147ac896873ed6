// boloop_driver.cpp -- BASELINE.json configs[4] with BOTH of the reference's nodes unmodified, in one process:
//   thread A: /root/reference/trial_node.cpp           (the evaluator node "robot_1d_kf") compiled against include/compat, i.e.
//             its Trial::Run calls are kernel launches on the B200 (as in tests/native/dropin)
//             -- or (evaluator = 1) THIS repository's ROS node, kf_bayesopt_b200/host/robot1d_kf.cpp under -DKFBO_WITH_ROS: one
//             fused launch per message; the deployment BASELINE.json configs[4] describes (bayesopt node unchanged, GPU evaluator)
//   thread B: /root/reference/robot1d_bayesopt.cpp     (the optimiser node "robot1d_bayes": Robot1D::evaluateSample builds the
//             "noise" message per optimizationChoice, publishes it, waits for "cost") compiled against the stand-in of the
//             absent bayesopt library (shim/bayesopt/bayesopt.hpp: the GP / EI optimiser of kf_bayesopt_b200/host/gp_ei.hpp)
// talking over the roscpp stand-in in bus mode (oracle/ref_shim/ros/ros.h).  Only main() is renamed in both, and the optimiser
// node's usleep() calls (0.5 s per iteration, robot1d_bayesopt.cpp:91,96) go through kfbo_boloop_usleep, which scales them.
// TEST INFRASTRUCTURE (tests/test_boloop.py).
#include <chrono>
#include <cstring>
#include <iostream>
#include <mutex>
#include <sstream>
#include <streambuf>
#include <string>
#include <thread>
#include <vector>

#include <ros/ros.h>
#include <std_msgs/Float64.h>
#include <std_msgs/Float64MultiArray.h>

int kfbo_boloop_trial_node_main(int nargs, char *args[]);   // trial_node.cpp's main()
int kfbo_boloop_robot1d_kf_main(int nargs, char *args[]);   // kf_bayesopt_b200/host/robot1d_kf.cpp's main(), -DKFBO_WITH_ROS
int kfbo_boloop_bayesopt_main(int nargs, char *args[]);     // robot1d_bayesopt.cpp's main()

namespace {
double g_sleep_scale = 0.0;
double g_slept_s = 0.0;

// std::cout of BOTH node threads lands here while the loop runs: the two nodes print at the same moment (the evaluator its
// "publish the cost", the optimiser its "get the cost"), so the buffer serialises its writers.
class LockedSink : public std::streambuf {
public:
    std::string text()
    {
        std::lock_guard<std::mutex> lk(mu_);
        return text_;
    }
protected:
    std::streamsize xsputn(const char *s, std::streamsize n) override
    {
        std::lock_guard<std::mutex> lk(mu_);
        text_.append(s, (size_t)n);
        return n;
    }
    int_type overflow(int_type c) override
    {
        if (c != traits_type::eof()) {
            std::lock_guard<std::mutex> lk(mu_);
            text_.push_back((char)c);
        }
        return c;
    }
private:
    std::mutex mu_;
    std::string text_;
};
}  // namespace

extern "C" {

// what robot1d_bayesopt.cpp's usleep() calls become (Makefile: -D'usleep(x)=kfbo_boloop_usleep(x)')
int kfbo_boloop_usleep(useconds_t us)
{
    const double s = 1e-6 * (double)us * g_sleep_scale;
    g_slept_s += s;
    if (s > 0) std::this_thread::sleep_for(std::chrono::duration<double>(s));
    return 0;
}

void kfbo_boloop_param_clear(void) { ros::ShimMaster::get().params.clear(); }
void kfbo_boloop_param_set_int(const char *name, long long v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kInt; p.i = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_boloop_param_set_double(const char *name, double v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kDouble; p.d = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_boloop_param_set_string(const char *name, const char *v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kString; p.s = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_boloop_param_set_vector(const char *name, const double *v, int n)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kDoubleVector; p.v.assign(v, v + n);
    ros::ShimMaster::get().params[name] = p;
}

// Runs the two nodes until the optimiser node's main() returns.  noise[i][0..3] = (layout.dim[0].size, layout.dim[1].size,
// data[0], data[1]) of the i-th "noise" message, costs[i] the i-th "cost" message; log: the END of what both nodes printed.
// Returns the number of "noise" messages; *n_costs the number of "cost" messages; *wall_s, *slept_s the loop's wall time and
// the time spent in the optimiser node's (scaled) sleeps.
// evaluator: 0 = the reference's trial_node.cpp on include/compat, 1 = this repository's robot1d_kf.cpp (its command line after
// the program name in evaluator_args, space separated; its yaml comes from the ~vehicle_yaml parameter).
int kfbo_boloop_run(int evaluator_kind, const char *evaluator_args, double sleep_scale, double *noise, double *costs, int cap, int *n_costs,
                    double *wall_s, double *slept_s, char *log, int log_cap)
{
    ros::ShimMaster &m = ros::ShimMaster::get();
    {
        std::lock_guard<std::mutex> lk(m.mu);
        m.inbox.clear(); m.outbox.clear(); m.callbacks.clear();
    }
    m.set_bus(true);
    g_sleep_scale = sleep_scale;
    g_slept_s = 0.0;
    LockedSink sink;
    std::streambuf *old = std::cout.rdbuf(&sink);
    char b0[] = "robot1d_bayesopt";
    char *args_b[] = {b0, nullptr};
    std::vector<std::string> words{"robot1d_kf"};
    for (const char *p = evaluator_args ? evaluator_args : ""; *p;) {
        while (*p == ' ') ++p;
        const char *e = p;
        while (*e && *e != ' ') ++e;
        if (e > p) words.emplace_back(p, e);
        p = e;
    }
    std::vector<char *> args_a;
    for (std::string &w : words) args_a.push_back(&w[0]);
    args_a.push_back(nullptr);
    std::thread evaluator([&] {
        try {
            if (evaluator_kind == 1) kfbo_boloop_robot1d_kf_main((int)words.size(), args_a.data());
            else kfbo_boloop_trial_node_main(1, args_a.data());
        } catch (const std::exception &e) { std::cerr << "evaluator node: " << e.what() << std::endl; }
    });
    for (int i = 0; i < 30000 && m.subscribers("noise") == 0; ++i) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    const auto t0 = std::chrono::steady_clock::now();
    try { kfbo_boloop_bayesopt_main(1, args_b); } catch (const std::exception &e) { std::cerr << "optimiser node: " << e.what() << std::endl; }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ros::shutdown();                       // ends the evaluator node's spin()
    evaluator.join();
    std::cout.rdbuf(old);
    if (wall_s) *wall_s = wall;
    if (slept_s) *slept_s = g_slept_s;
    const std::string printed = sink.text();
    if (log && log_cap > 0) {
        const size_t n = printed.size() < (size_t)log_cap - 1 ? printed.size() : (size_t)log_cap - 1;
        std::memcpy(log, printed.data() + (printed.size() - n), n);
        log[n] = 0;
    }
    int n = 0, nc = 0;
    {
        std::lock_guard<std::mutex> lk(m.mu);
        for (const std::shared_ptr<void> &p : m.outbox["noise"]) {
            if (n < cap) {
                const std_msgs::Float64MultiArray &msg = *std::static_pointer_cast<std_msgs::Float64MultiArray>(p);
                noise[4 * n + 0] = msg.layout.dim.size() > 0 ? msg.layout.dim[0].size : -1;
                noise[4 * n + 1] = msg.layout.dim.size() > 1 ? msg.layout.dim[1].size : -1;
                noise[4 * n + 2] = msg.data.size() > 0 ? msg.data[0] : 0.0;
                noise[4 * n + 3] = msg.data.size() > 1 ? msg.data[1] : 0.0;
            }
            ++n;
        }
        for (const std::shared_ptr<void> &p : m.outbox["cost"]) {
            if (nc < cap) costs[nc] = std::static_pointer_cast<std_msgs::Float64>(p)->data;
            ++nc;
        }
        m.inbox.clear(); m.outbox.clear(); m.callbacks.clear();
    }
    m.set_bus(false);
    if (n_costs) *n_costs = nc;
    return n;
}

}  // extern "C"
