// boost/math/constants/constants.hpp -- stand-in: robot1d_bayesopt.cpp includes it and uses nothing of it.
#pragma once
