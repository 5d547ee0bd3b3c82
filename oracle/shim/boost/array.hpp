// oracle/shim: boost::array -> std::array (used at forcing_component.cpp:402).
#pragma once
#include "config.hpp"
#include <array>
namespace boost {
template <class T, std::size_t N> using array = std::array<T, N>;
}
