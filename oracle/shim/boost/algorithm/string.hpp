// oracle/shim: the four boost string algorithms the reference readers use.
#pragma once
#include "../config.hpp"
#include <cctype>
namespace boost {
struct shim_is_any_of {
  std::string set;
  bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline shim_is_any_of is_any_of(const std::string &s) { return shim_is_any_of{s}; }
template <class Seq, class Pred>
Seq &split(Seq &out, const std::string &in, Pred pred) {
  out.clear();
  std::string cur;
  for (char c : in) {
    if (pred(c)) {
      out.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  out.push_back(cur);
  return out;
}
inline void trim(std::string &s) {
  std::size_t b = 0, e = s.size();
  while (b < e && std::isspace(static_cast<unsigned char>(s[b]))) ++b;
  while (e > b && std::isspace(static_cast<unsigned char>(s[e - 1]))) --e;
  s = s.substr(b, e - b);
}
inline bool starts_with(const std::string &s, const std::string &prefix) {
  return s.size() >= prefix.size() && s.compare(0, prefix.size(), prefix) == 0;
}
} // namespace boost
