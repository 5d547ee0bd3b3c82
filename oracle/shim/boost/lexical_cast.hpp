// oracle/shim: boost::lexical_cast for the handful of conversions the reference uses
// (string -> double in the readers; arithmetic -> string in visitors/messages).
#pragma once
#include "config.hpp"
#include <cstdlib>
#include <cerrno>
#include <sstream>
#include <stdexcept>
#include <type_traits>
#include <typeinfo>
namespace boost {
class bad_lexical_cast : public std::bad_cast {
public:
  const char *what() const noexcept override {
    return "bad lexical cast: source type value could not be interpreted as target";
  }
};
namespace shim_detail {
template <class Target, class Source, class Enable = void> struct caster;
// string -> floating point: whole string must be consumed, no leading blanks
template <class Target>
struct caster<Target, std::string,
              typename std::enable_if<std::is_floating_point<Target>::value>::type> {
  static Target run(const std::string &s) {
    if (s.empty() || std::isspace(static_cast<unsigned char>(s[0])))
      throw bad_lexical_cast();
    char *end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    if (end != s.c_str() + s.size())
      throw bad_lexical_cast();
    return static_cast<Target>(v);
  }
};
template <class Target>
struct caster<Target, std::string,
              typename std::enable_if<std::is_integral<Target>::value>::type> {
  static Target run(const std::string &s) {
    if (s.empty() || std::isspace(static_cast<unsigned char>(s[0])))
      throw bad_lexical_cast();
    char *end = nullptr;
    long long v = std::strtoll(s.c_str(), &end, 10);
    if (end != s.c_str() + s.size())
      throw bad_lexical_cast();
    return static_cast<Target>(v);
  }
};
// arithmetic -> string
template <class Source>
struct caster<std::string, Source,
              typename std::enable_if<std::is_arithmetic<Source>::value>::type> {
  static std::string run(const Source &v) {
    std::ostringstream os;
    if (std::is_floating_point<Source>::value)
      os.precision(17);
    os << v;
    return os.str();
  }
};
template <> struct caster<std::string, std::string, void> {
  static std::string run(const std::string &v) { return v; }
};
} // namespace shim_detail
template <class Target, class Source> Target lexical_cast(const Source &s) {
  return shim_detail::caster<Target, Source>::run(s);
}
template <class Target> Target lexical_cast(const char *s) {
  return shim_detail::caster<Target, std::string>::run(std::string(s));
}
} // namespace boost
