#pragma once
namespace boost {
template <class EnginePtr, class Dist> class variate_generator;
template <class Engine, class Dist> class variate_generator<Engine*, Dist> {
  Engine* e_; Dist d_;
 public:
  variate_generator(Engine* e, Dist d) : e_(e), d_(d) {}
  typename Dist::result_type operator()() { return d_(*e_); }
};
}
