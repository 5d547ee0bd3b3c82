/* include/compat/h_exception.hpp -- stands in for inst/include/h_exception.hpp:26-102: the global
 * h_exception (defined in include/hector_b200_core.hpp) and the H_THROW / H_ASSERT macros. */
#ifndef HECTOR_B200_COMPAT_H_EXCEPTION_HPP
#define HECTOR_B200_COMPAT_H_EXCEPTION_HPP
#include "core.hpp"
#define H_THROW(s) throw h_exception(s, __func__, __FILE__, __LINE__);
#define H_ASSERT(x, s)                                                         \
  if (!(x))                                                                    \
    H_THROW("Assertion failed: " + std::string(s));
#endif
