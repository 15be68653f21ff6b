// std-based stand-in for <boost/shared_ptr.hpp> (oracle build only; dropout is the sole user and runs at rate 0)
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; }
