// specialtypes.hpp -- stand-in for bayesopt's header of that name: the vector type of the optimiser's interface.
#pragma once
#include <boost/numeric/ublas/vector.hpp>
typedef boost::numeric::ublas::vector<double> vectord;
