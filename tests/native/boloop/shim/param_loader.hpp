// param_loader.hpp -- stand-in for bayesopt's utils/param_loader.hpp: robot1d_bayesopt.cpp:235-240 loads a bayesopt parameter
// file when the node is given one on its command line.  The tests never do; loading is refused.
#pragma once
#include "bayesopt/bayesopt.hpp"
namespace bayesopt { namespace utils {
struct ParamLoader {
    static bool load(const char * /*file*/, Parameters & /*par*/) { return false; }
};
}}  // namespace bayesopt::utils
