// std_msgs/Float64MultiArray.h -- the message shape of the "noise" topic (std_msgs/Float64MultiArray:
// MultiArrayLayout layout {MultiArrayDimension[] dim {string label, uint32 size, uint32 stride}, uint32
// data_offset}, float64[] data), so that the reference's trial_node.cpp compiles without ROS.
// Test infrastructure (see ../ros/ros.h).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace std_msgs {
struct MultiArrayDimension {
    std::string label;
    uint32_t size = 0;
    uint32_t stride = 0;
};
struct MultiArrayLayout {
    std::vector<MultiArrayDimension> dim;
    uint32_t data_offset = 0;
};
struct Float64MultiArray {
    MultiArrayLayout layout;
    std::vector<double> data;
    typedef std::shared_ptr<Float64MultiArray> Ptr;
    typedef std::shared_ptr<const Float64MultiArray> ConstPtr;
};
}  // namespace std_msgs
