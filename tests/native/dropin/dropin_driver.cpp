// dropin_driver.cpp -- C entry points around THE REFERENCE'S UNMODIFIED NODE SOURCES compiled against the B200
// evaluator: /root/reference/trial_node.cpp and /root/reference/src/test_trial.cpp are built where they lie with
// include/compat ahead of the reference's include/ (so "trial.h" / "read_para_vehicle.h" are this repository's
// drop-ins over the C ABI) and oracle/ref_shim/ros standing in for roscpp.  Only main() is renamed (Makefile).
// TEST INFRASTRUCTURE: tests/test_gpu_dropin.py queues "noise" messages, runs the node and reads its "cost" messages.
#include <iostream>
#include <sstream>

#include <ros/ros.h>
#include <std_msgs/Float64.h>
#include <std_msgs/Float64MultiArray.h>

int kfbo_dropin_trial_node_main(int nargs, char *args[]);   // trial_node.cpp's main()
int kfbo_dropin_test_trial_main(int nargs, char *args[]);   // src/test_trial.cpp's main()

namespace {
struct CaptureCout {
    std::ostringstream sink;
    std::streambuf *old;
    CaptureCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CaptureCout() { std::cout.rdbuf(old); }
};
int copy_out(const std::string &s, char *buf, int cap)
{
    if (buf && cap > 0) {
        const size_t n = s.size() < (size_t)cap - 1 ? s.size() : (size_t)cap - 1;
        s.copy(buf, n);
        buf[n] = 0;
    }
    return (int)s.size();
}
}  // namespace

extern "C" {

void kfbo_dropin_param_clear(void) { ros::ShimMaster::get().params.clear(); }
void kfbo_dropin_param_set_int(const char *name, long long v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kInt; p.i = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_dropin_param_set_double(const char *name, double v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kDouble; p.d = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_dropin_param_set_string(const char *name, const char *v)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kString; p.s = v;
    ros::ShimMaster::get().params[name] = p;
}
void kfbo_dropin_param_set_vector(const char *name, const double *v, int n)
{
    ros::ShimParam p; p.kind = ros::ShimParam::kDoubleVector; p.v.assign(v, v + n);
    ros::ShimMaster::get().params[name] = p;
}

// n_msgs "noise" messages (message m: dim[0].size = npn[m], dim[1].size = non[m], data = the next npn[m] + non[m]
// doubles) -> the reference's trial_node.cpp main() -> the "cost" messages it published.  log: what the node printed.
int kfbo_dropin_node_run(int n_msgs, const int *npn, const int *non, const double *data, double *costs, int costs_cap,
                         char *log, int log_cap)
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
    std::string printed;
    {
        CaptureCout cap;
        char arg0[] = "robot1d_kf";
        char *args[] = {arg0, nullptr};
        kfbo_dropin_trial_node_main(1, args);
        printed = cap.sink.str();
    }
    copy_out(printed, log, log_cap);
    int n = 0;
    for (const std::shared_ptr<void> &p : m.outbox["cost"]) {
        if (n < costs_cap) costs[n] = std::static_pointer_cast<std_msgs::Float64>(p)->data;
        ++n;
    }
    m.outbox["cost"].clear();
    return n;
}

// src/test_trial.cpp's main() (runTrial::run: one sample time, fixed pnoise_est / onoise_est parameters); it
// publishes nothing but "state_his" and prints its cost, so the text it printed is the result.
int kfbo_dropin_test_trial_run(char *log, int log_cap)
{
    ros::ShimMaster::get().outbox["state_his"].clear();
    std::string printed;
    {
        CaptureCout cap;
        char arg0[] = "test_trial";
        char *args[] = {arg0, nullptr};
        kfbo_dropin_test_trial_main(1, args);
        printed = cap.sink.str();
    }
    ros::ShimMaster::get().outbox["state_his"].clear();
    return copy_out(printed, log, log_cap);
}

}  // extern "C"
