/* include/compat/unitval.hpp -- stands in for inst/include/unitval.hpp; the types live in
 * include/hector_b200_core.hpp (namespace Hector = hector_b200). */
#ifndef HECTOR_B200_COMPAT_UNITVAL_HPP
#define HECTOR_B200_COMPAT_UNITVAL_HPP
#include "core.hpp"
#endif
