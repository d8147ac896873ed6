// std_msgs/Float64.h -- the message shape of the "cost" topic (std_msgs/Float64: float64 data), so that the
// reference's trial_node.cpp compiles without ROS.  Test infrastructure (see ../ros/ros.h).
#pragma once
#include <memory>
namespace std_msgs {
struct Float64 {
    double data = 0.0;
    typedef std::shared_ptr<Float64> Ptr;
    typedef std::shared_ptr<const Float64> ConstPtr;
};
}  // namespace std_msgs
