/* tests/cpp/rcpp_stub/Rcpp.h -- TEST INFRASTRUCTURE.  Just enough of the Rcpp API for the
 * reference's R glue (src/rcpp_hector.cpp, compiled unmodified by tests/test_compat_cpu.py against
 * include/compat) to build and be driven from C++ without R: Environment, String, NumericVector,
 * StringVector, DataFrame::create / Named, Function, stop, new_env.  Not an R binding. */
#ifndef RCPP_STUB_H
#define RCPP_STUB_H
#include <cmath>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace Rcpp {

struct exception : std::runtime_error {
  explicit exception(const std::string &m) : std::runtime_error(m) {}
};
[[noreturn]] inline void stop(const std::string &msg) { throw exception(msg); }

class String {
  std::string s_;

 public:
  String() {}
  String(const char *s) : s_(s) {}
  String(const std::string &s) : s_(s) {}
  const char *get_cstring() const { return s_.c_str(); }
  operator std::string() const { return s_; }
};

/* one slot of an environment: a number, a flag or a string */
struct Value {
  double num = 0.0;
  std::string str;
  Value() {}
  Value(int v) : num(v) {}
  Value(double v) : num(v) {}
  Value(bool v) : num(v ? 1 : 0) {}
  Value(const String &v) : str(v) {}
  Value(const std::string &v) : str(v) {}
  Value(const char *v) : str(v) {}
  operator int() const { return (int)num; }
  operator double() const { return num; }
  operator bool() const { return num != 0.0; }
  bool operator!() const { return num == 0.0; }
  operator std::string() const { return str; }
};

class Environment {
  std::shared_ptr<std::map<std::string, Value>> m_;

 public:
  Environment() : m_(std::make_shared<std::map<std::string, Value>>()) {}
  Value &operator[](const std::string &k) { return (*m_)[k]; }
  const Value &operator[](const std::string &k) const { return (*m_)[k]; }
};
inline Environment new_env() { return Environment(); }

class NumericVector {
  std::vector<double> v_;

 public:
  NumericVector() {}
  explicit NumericVector(int n) : v_(n, 0.0) {}
  NumericVector(std::initializer_list<double> l) : v_(l) {}
  int size() const { return (int)v_.size(); }
  double &operator[](int i) { return v_[i]; }
  double operator[](int i) const { return v_[i]; }
  static bool is_na(double x) { return std::isnan(x); }
  static double get_na() { return std::numeric_limits<double>::quiet_NaN(); }
  const std::vector<double> &data() const { return v_; }
};
class StringVector {
  std::vector<std::string> v_;

 public:
  explicit StringVector(int n = 0) : v_(n) {}
  std::string &operator[](int i) { return v_[i]; }
  const std::vector<std::string> &data() const { return v_; }
};

/* DataFrame::create(Named("year") = date, ...): keeps the numeric and string columns by name */
struct Column {
  std::string name;
  std::vector<double> num;
  std::vector<std::string> str;
};
struct Named {
  std::string name;
  explicit Named(const std::string &n) : name(n) {}
  Column operator=(const NumericVector &v) const { Column c; c.name = name; c.num = v.data(); return c; }
  Column operator=(const StringVector &v) const { Column c; c.name = name; c.str = v.data(); return c; }
  Column operator=(const String &v) const { Column c; c.name = name; c.str.push_back(v); return c; }
  Column operator=(bool v) const { Column c; c.name = name; c.num.push_back(v); return c; }
};
class DataFrame {
 public:
  std::vector<Column> columns;
  template <class... C>
  static DataFrame create(const C &...cols) {
    DataFrame d;
    (d.columns.push_back(cols), ...);
    return d;
  }
  const Column &col(const std::string &n) const {
    for (const Column &c : columns)
      if (c.name == n) return c;
    throw exception("no such column: " + n);
  }
};

class Function { /* Function f("message"); f(msg); */
  std::string name_;

 public:
  explicit Function(const std::string &n) : name_(n) {}
  void operator()(const std::string &msg) const { std::clog << name_ << ": " << msg << std::endl; }
};

} // namespace Rcpp
#endif
