// matplotlibcpp.h -- stand-in for the plotting header trial.cpp includes (its Nsimruns == 1 path, trial.cpp:116-129):
// every call is a no-op.  Test infrastructure, see Eigen/Dense in this directory.
#pragma once
#include <string>
namespace matplotlibcpp {
template <typename... A> bool plot(const A &...) { return true; }
inline void subplot(long, long, long) {}
inline void ylabel(const std::string &) {}
inline void xlabel(const std::string &) {}
inline void title(const std::string &) {}
inline void show() {}
}  // namespace matplotlibcpp
