/* include/compat/logger.hpp -- stands in for inst/include/logger.hpp:35-118: Hector::Logger
 * (defined in include/hector_b200_core.hpp) and the H_LOG macro. */
#ifndef HECTOR_B200_COMPAT_LOGGER_HPP
#define HECTOR_B200_COMPAT_LOGGER_HPP
#include "core.hpp"
#define H_LOG(log, level)                                                      \
  if (log.shouldWrite(level))                                                  \
  log.write(level, __func__)
#endif
