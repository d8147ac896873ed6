// compat/trial.h -- drop-in for the reference's include/trial.h (+ trial.cpp, simulator.cpp, kalman_filter.cpp and
// the model headers behind it): the global class `Trial` with the reference's constructor, Run() and four getters
// (include/trial.h:7-59), evaluated on the B200 through the C ABI of kfbo.h.
//
// With include/compat ahead of the reference's include/ on the include path, the reference's own callers compile
// UNMODIFIED and keep their shape (trial_node.cpp:55-75, src/test_trial.cpp:41-51):
//     std::vector<Trial> t;  t.push_back(Trial(rv_gt, rv));  std::thread(&Trial::Run, &t[i]);  t[i].get_average_nees()
// Each Trial::Run is one thread's share of the Monte Carlo batch (nsimruns_/cpu_core_number_ trials, trial.cpp:28)
// as one kernel launch with its own Philox stream; evaluator handles are pooled across Trial objects
// (kfbo::detail::EvaluatorPool), so the CPUcoreNumber x 2 Trials a callback creates cost launches, not allocations.
// tests/native/dropin builds the reference's trial_node.cpp this way and tests/test_dropin.py runs it on the GPU.
//
// kfbo::RunTrial (kfbo_trial.hpp) is the faster way in -- one launch for all threads' shares and both sample
// times -- when the caller can change.
#ifndef KFBO_COMPAT_TRIAL_H_
#define KFBO_COMPAT_TRIAL_H_

#include "read_para_vehicle.h"

// include/system_models.hpp:11-12, which the reference's trial.h pulls in and src/test_trial.cpp:57 uses
#ifndef N
#define N 2
#endif
#ifndef Me
#define Me 1
#endif

class Trial : public kfbo::Trial {
public:
    Trial(ReadParaVehicle &simulator_parameters, ReadParaVehicle &estimator_parameters)
        : kfbo::Trial(simulator_parameters, estimator_parameters)
    {
    }
};

#endif  // KFBO_COMPAT_TRIAL_H_
