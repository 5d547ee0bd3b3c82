/* include/compat/h_util.hpp -- stands in for inst/include/h_util.hpp:23-37. */
#ifndef HECTOR_B200_COMPAT_H_UTIL_HPP
#define HECTOR_B200_COMPAT_H_UTIL_HPP
#include <sys/stat.h>

#include <string>

#include "h_exception.hpp"

#define MODEL_NAME "hector"
#define MODEL_VERSION "3.5.0 (hector_b200)"
#define OUTPUT_DIRECTORY "output/"

namespace hector_b200 {
/* h_util.cpp: make sure `dir` exists, creating it if need be; throws when it cannot */
inline void ensure_dir_exists(const std::string &dir) {
  struct stat st;
  if (stat(dir.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) return;
  if (mkdir(dir.c_str(), 0755) != 0 && !(stat(dir.c_str(), &st) == 0 && S_ISDIR(st.st_mode)))
    H_THROW("Directory " + dir + " does not exist and could not be created.")
}
} // namespace hector_b200
#endif
