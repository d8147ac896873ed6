// boost/numeric/ublas/vector.hpp -- stand-in (Boost is not installed): what the reference's robot1d_bayesopt.cpp uses of
// ublas::vector<double> -- size constructor, operator()(i), size(), begin() / end().  Test infrastructure (tests/native/boloop/Makefile).
#pragma once
#include <cstddef>
#include <vector>
namespace boost { namespace numeric { namespace ublas {
template <typename T>
class vector {
public:
    typedef typename std::vector<T>::iterator iterator;
    typedef typename std::vector<T>::const_iterator const_iterator;
    vector() {}
    explicit vector(std::size_t n) : v_(n) {}
    vector(std::size_t n, const T &x) : v_(n, x) {}
    T &operator()(std::size_t i) { return v_[i]; }
    const T &operator()(std::size_t i) const { return v_[i]; }
    T &operator[](std::size_t i) { return v_[i]; }
    const T &operator[](std::size_t i) const { return v_[i]; }
    std::size_t size() const { return v_.size(); }
    iterator begin() { return v_.begin(); }
    iterator end() { return v_.end(); }
    const_iterator begin() const { return v_.begin(); }
    const_iterator end() const { return v_.end(); }
private:
    std::vector<T> v_;
};
}}}  // namespace boost::numeric::ublas
