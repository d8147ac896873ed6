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
        nsimruns_ = 10;                                          // the reference's in-class default (read_para_vehicle.h:34)
        // parameter name -> member, by type; a name the server does not hold leaves its member untouched
        const struct { const char *name; int *member; } ints[] = {
            {"CPUcoreNumber", &cpu_core_number_}, {"stateDOF", &state_dof_}, {"observationDOF", &observation_dof_},
            {"Nsimruns", &nsimruns_}};
        const struct { const char *name; double *member; } reals[] = {
            {"dt", &dt_}, {"t_start", &t_start_}, {"t_end", &t_end_}, {"xi0", &xi0_}, {"xidot0", &xidot0_}};
        const struct { const char *name; std::string *member; } strings[] = {
            {"P0", &P0_}, {"optimizationChoice", &optimization_choice_}, {"costChoice", &cost_choice_}};
        for (const auto &e : ints) nh.getParam(e.name, *e.member);
        for (const auto &e : reals) nh.getParam(e.name, *e.member);
        for (const auto &e : strings) nh.getParam(e.name, *e.member);
        nh.getParam("disturbance", disturbance_);
        const bool have_noise = nh.getParam("pnoise", pnoise_) & nh.getParam("onoise", onoise_);
        if (!have_noise) ROS_FATAL("process noise and observation noise must be set");          // read_para_vehicle.cpp:95,102
        dim_pn_ = (int)pnoise_.size();
        dim_on_ = (int)onoise_.size();
        if (cpu_core_number_ > 0 && nsimruns_ % cpu_core_number_ != 0)                            // :141-142
            ROS_WARN("Nsimruns is not a multiple of CPUcoreNumber: the remainder trials are dropped (trial.cpp:28)");
        UpdateMaxIterations();                                   // (t_end - t_start)/dt, truncated (:145)
        if (showReadInfo) Print(std::cout);
    }
    double tmpf = 17.32;                                         // read_para_vehicle.h:48
};

#endif  // KFBO_COMPAT_READ_PARA_VEHICLE_H_
