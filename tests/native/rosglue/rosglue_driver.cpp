// rosglue_driver.cpp -- C entry point around this repository's ROS node (kf_bayesopt_b200/host/robot1d_kf.cpp compiled with
// -DKFBO_WITH_ROS, main() renamed) on the roscpp stand-in of oracle/ref_shim: set the private parameter, queue "noise"
// messages, run the node's main() -- ros::spin() delivers the queue and returns --, hand back the "cost" messages it published.
// TEST INFRASTRUCTURE (tests/test_rosglue.py).
#include <cstring>
#include <string>
#include <vector>

#include <ros/ros.h>
#include <std_msgs/Float64.h>
#include <std_msgs/Float64MultiArray.h>

int kfbo_rosglue_robot1d_kf_main(int nargs, char *args[]);   // robot1d_kf.cpp's main()

extern "C" {

// message m: layout.dim[0].size = npn[m], layout.dim[1].size = non[m], data = the next ndata[m] doubles (ndata[m] != npn[m] + non[m]
// makes a malformed message).  argv: the node's command line after its name, space separated ("--seed 7 --quiet").
// Returns the number of "cost" messages published; the first costs_cap of them in costs.
int kfbo_rosglue_run(const char *vehicle_yaml, const char *argv_line, int n_msgs, const int *npn, const int *non, const int *ndata,
                     const double *data, double *costs, int costs_cap, int *exit_code)
{
    ros::ShimMaster &m = ros::ShimMaster::get();
    m.params.clear();
    m.inbox["noise"].clear();
    m.outbox["cost"].clear();
    m.callbacks.clear();
    if (vehicle_yaml && *vehicle_yaml) {
        ros::ShimParam p; p.kind = ros::ShimParam::kString; p.s = vehicle_yaml;
        m.params["~vehicle_yaml"] = p;
    }
    for (int i = 0; i < n_msgs; ++i) {
        std::shared_ptr<std_msgs::Float64MultiArray> msg = std::make_shared<std_msgs::Float64MultiArray>();
        msg->layout.dim.resize(2);                      // robot1d_bayesopt.cpp:53-63
        msg->layout.dim[0].size = (uint32_t)npn[i];
        msg->layout.dim[1].size = (uint32_t)non[i];
        msg->data.assign(data, data + ndata[i]);
        data += ndata[i];
        m.inbox["noise"].push_back(msg);
    }
    std::vector<std::string> words{"robot1d_kf"};
    for (const char *p = argv_line ? argv_line : ""; *p;) {
        while (*p == ' ') ++p;
        const char *e = p;
        while (*e && *e != ' ') ++e;
        if (e > p) words.emplace_back(p, e);
        p = e;
    }
    std::vector<char *> args;
    for (std::string &w : words) args.push_back(&w[0]);
    args.push_back(nullptr);
    const int rc = kfbo_rosglue_robot1d_kf_main((int)words.size(), args.data());
    if (exit_code) *exit_code = rc;
    int n = 0;
    for (const std::shared_ptr<void> &p : m.outbox["cost"]) {
        if (n < costs_cap) costs[n] = std::static_pointer_cast<std_msgs::Float64>(p)->data;
        ++n;
    }
    m.outbox["cost"].clear();
    return n;
}

}  // extern "C"
