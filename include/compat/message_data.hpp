/* include/compat/message_data.hpp -- stands in for inst/include/message_data.hpp; the types live in
 * include/hector_b200_core.hpp (namespace Hector = hector_b200). */
#ifndef HECTOR_B200_COMPAT_MESSAGE_DATA_HPP
#define HECTOR_B200_COMPAT_MESSAGE_DATA_HPP
#include "core.hpp"
#endif
