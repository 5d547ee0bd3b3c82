// oracle/shim: boost::math::tools::polynomial<T> -- only construction from a coefficient
// array (ascending powers) and Horner evaluation are used (ocean_csys.cpp:104-118).
#pragma once
#include "../../config.hpp"
namespace boost { namespace math { namespace tools {
template <class T> class polynomial {
public:
  polynomial() {}
  template <class U> polynomial(const U *data, unsigned order) : m_data(data, data + order + 1) {}
  // evaluate_polynomial(poly, z, count): sum = poly[count-1]; for i = count-2..0: sum *= z; sum += poly[i]
  T evaluate(T z) const {
    if (m_data.empty()) return T(0);
    T sum = m_data[m_data.size() - 1];
    for (int i = static_cast<int>(m_data.size()) - 2; i >= 0; --i) {
      sum *= z;
      sum += m_data[i];
    }
    return sum;
  }
private:
  std::vector<T> m_data;
};
}}} // namespace boost::math::tools
