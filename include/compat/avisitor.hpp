/* include/compat/avisitor.hpp -- stands in for inst/include/avisitor.hpp; the types live in
 * include/hector_b200_core.hpp (namespace Hector = hector_b200). */
#ifndef HECTOR_B200_COMPAT_AVISITOR_HPP
#define HECTOR_B200_COMPAT_AVISITOR_HPP
#include "core.hpp"
#endif
