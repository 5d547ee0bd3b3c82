/* include/compat/ini_to_core_reader.hpp -- stands in for inst/include/ini_to_core_reader.hpp; the types live in
 * include/hector_b200_core.hpp (namespace Hector = hector_b200). */
#ifndef HECTOR_B200_COMPAT_INI_TO_CORE_READER_HPP
#define HECTOR_B200_COMPAT_INI_TO_CORE_READER_HPP
#include "core.hpp"
#endif
