// oracle/shim: boost::math::lognormal + cdf (simpleNbox-runtime.cpp:154, 1028).
// cdf(x) = erfc(-(ln x - mu)/(sigma*sqrt2))/2, 0 at x == 0.  std::erfc stands in for
// boost::math::erfc (knowing deviation, ~1 ulp; permafrost_c still matches the golden file).
#pragma once
#include "../../config.hpp"
#include <stdexcept>
namespace boost { namespace math {
class lognormal {
public:
  lognormal(double location = 0, double scale = 1) : m_location(location), m_scale(scale) {}
  double location() const { return m_location; }
  double scale() const { return m_scale; }
private:
  double m_location, m_scale;
};
typedef lognormal lognormal_distribution;
inline double cdf(const lognormal &dist, const double &x) {
  if (!(x >= 0)) throw std::domain_error("lognormal cdf: x < 0");
  if (x == 0) return 0;
  const double root_two = 1.414213562373095048801688724209698078569671875376948073176679737990732478462;
  double diff = (std::log(x) - dist.location()) / (dist.scale() * root_two);
  return std::erfc(-diff) / 2;
}
}} // namespace boost::math
