#pragma once
#include <random>
namespace boost {
template <typename RealT = double> class bernoulli_distribution {
  RealT p_;
 public:
  typedef bool result_type;
  explicit bernoulli_distribution(RealT p = RealT(0.5)) : p_(p) {}
  template <class Engine> bool operator()(Engine& eng) {
    std::bernoulli_distribution d(static_cast<double>(p_));
    return d(eng);
  }
};
}
