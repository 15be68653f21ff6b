// std-based stand-in for boost::thread_specific_ptr (oracle build only)
#pragma once
#include <memory>
namespace boost {
template <typename T> class thread_specific_ptr {
  static std::unique_ptr<T>& slot() { static thread_local std::unique_ptr<T> p; return p; }
 public:
  T* get() const { return slot().get(); }
  void reset(T* p = nullptr) { slot().reset(p); }
  T& operator*() const { return *slot(); }
  T* operator->() const { return slot().get(); }
};
}
