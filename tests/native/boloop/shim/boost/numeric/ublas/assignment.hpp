// boost/numeric/ublas/assignment.hpp -- stand-in: robot1d_bayesopt.cpp includes it and uses nothing of it beyond the vector type.
#pragma once
#include "vector.hpp"
