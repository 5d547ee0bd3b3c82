/* include/compat/h_reader.hpp -- stands in for inst/include/h_reader.hpp:30-42: the CLI opens the
 * ini file once with an h_reader to find out early whether it parses (src/main.cpp:49). */
#ifndef HECTOR_B200_COMPAT_H_READER_HPP
#define HECTOR_B200_COMPAT_H_READER_HPP
#include <string>

#include "h_exception.hpp"

namespace hector_b200 {
enum readertype_t { INI_style, table_style };
class h_reader {
  std::string filename_;

 public:
  h_reader(std::string fname, readertype_t /*style*/, bool doparse = true) : filename_(fname) {
    if (doparse) parse();
  }
  virtual ~h_reader() {}
  void parse() { /* the engine's own ini / csv reader; no device needed */
    int32_t s = 0, e = 0;
    if (hx_ini_read(filename_.c_str(), &s, &e, nullptr, 0) != HX_OK)
      H_THROW(std::string("Parse error in file ") + filename_ + ": " + hx_last_error(nullptr))
  }
  double get_number(std::string /*section*/, std::string name, double defaultvalue) {
    double v = defaultvalue;
    return hx_ini_scalar(filename_.c_str(), name.c_str(), &v) == HX_OK ? v : defaultvalue;
  }
  std::string get_string(std::string /*section*/, std::string name, std::string defaultvalue) {
    char buf[512];
    return hx_ini_string(filename_.c_str(), name.c_str(), buf, (int32_t)sizeof buf) == HX_OK
               ? std::string(buf) : defaultvalue;
  }
};
} // namespace hector_b200
#endif
