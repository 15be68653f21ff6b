#pragma once
#include <random>
namespace boost { template <typename T = int> using uniform_int = std::uniform_int_distribution<T>; }
