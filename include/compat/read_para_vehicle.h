// compat/read_para_vehicle.h -- drop-in for the reference's include/read_para_vehicle.h (+ read_para_vehicle.cpp),
// so that the reference's OWN callers -- trial_node.cpp, src/test_trial.cpp -- compile UNMODIFIED against the B200
// evaluator:  put include/compat ahead of the reference's include/ on the include path and link libkfbo.so.
//
//   class ReadParaVehicle           include/read_para_vehicle.h:9-49: same public members (inherited from
//                                   kfbo::ReadParaVehicle), same two constructors
//   ReadParaVehicle(nh, show)       read_para_vehicle.cpp:5-158: the same parameter names read from the ROS parameter
//                                   server; a missing parameter leaves the member at its default (the reference
//                                   leaves it uninitialised); max_iterations_ = (t_end_ - t_start_)/dt_ truncated (:145)
//
// <ros/ros.h> is whatever the build provides: roscpp for a user of the reference, the stand-in of
// oracle/ref_shim in this repository's tests (tests/native/dropin).  Only NodeHandle::getParam is used.
#ifndef KFBO_COMPAT_READ_PARA_VEHICLE_H_
#define KFBO_COMPAT_READ_PARA_VEHICLE_H_

#include <ros/ros.h>

#include "../kfbo_trial.hpp"

class ReadParaVehicle : public kfbo::ReadParaVehicle {
public:
    ReadParaVehicle() {}
    ReadParaVehicle(ros::NodeHandle nh, bool showReadInfo)
    {
        nh.getParam("CPUcoreNumber", cpu_core_number_);
        nh.getParam("stateDOF", state_dof_);
        nh.getParam("observationDOF", observation_dof_);
        nh.getParam("dt", dt_);
        nh.getParam("t_start", t_start_);
        nh.getParam("t_end", t_end_);
        nh.getParam("xi0", xi0_);
        nh.getParam("xidot0", xidot0_);
        nh.getParam("P0", P0_);
        nh.getParam("disturbance", disturbance_);
        nsimruns_ = 10;                                          // read_para_vehicle.h:34
        nh.getParam("Nsimruns", nsimruns_);
        if (cpu_core_number_ > 0 && nsimruns_ % cpu_core_number_ != 0)
            ROS_WARN("the nsimruns should be an integral multiple of cpu core number, or the final result may be slightly off");
        UpdateMaxIterations();                                   // read_para_vehicle.cpp:145
        if (!nh.getParam("pnoise", pnoise_)) ROS_FATAL("process noise must be set");
        if (!nh.getParam("onoise", onoise_)) ROS_FATAL("observation noise must be set");
        dim_pn_ = (int)pnoise_.size();
        dim_on_ = (int)onoise_.size();
        nh.getParam("optimizationChoice", optimization_choice_);
        nh.getParam("costChoice", cost_choice_);
        if (showReadInfo) Print(std::cout);
    }
    double tmpf = 17.32;                                         // read_para_vehicle.h:48
};

#endif  // KFBO_COMPAT_READ_PARA_VEHICLE_H_
