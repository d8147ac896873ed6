// ros/ros.h -- stand-in so that the reference's read_para_vehicle.h / .cpp compile without ROS (test
// infrastructure, see Eigen/Dense in this directory).  The parameter server is empty: getParam finds nothing;
// oracle/ref_driver.cpp fills the ReadParaVehicle members directly.
#pragma once
#include <string>
namespace ros {
class NodeHandle {
public:
    template <typename T> bool getParam(const std::string &, T &) const { return false; }
};
inline bool ok() { return true; }
inline void shutdown() {}
}  // namespace ros
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_FATAL(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
