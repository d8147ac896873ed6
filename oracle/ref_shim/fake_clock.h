// fake_clock.h -- force-included (-include) before every reference source: the reference seeds its RNG from
// std::chrono::system_clock::now() (system_models.hpp:74,147, simulator.cpp:9), which makes it unrepeatable.
// Here `system_clock` names a counter the driver sets, so a run of the reference's own code is reproducible
// and the oracle can replay the same seeds.  Test infrastructure only.
#pragma once
#include <chrono>
extern "C" long long kfbo_ref_clock_next(void);
namespace std {
namespace chrono {
struct kfbo_ref_counter_clock {
    typedef std::chrono::nanoseconds duration;
    typedef duration::rep rep;
    typedef duration::period period;
    typedef std::chrono::time_point<kfbo_ref_counter_clock, duration> time_point;
    static time_point now() { return time_point(duration(kfbo_ref_clock_next())); }
};
}  // namespace chrono
}  // namespace std
#define system_clock kfbo_ref_counter_clock
