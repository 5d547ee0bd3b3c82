// oracle/shim: boost::math::tools::newton_raphson_iterate (Boost >= 1.7x algorithm), the
// bracketed / damped Newton iteration used at ocean_csys.cpp:152-153.  f(x) returns the
// pair (f, f').  Restated from the published Boost algorithm; validated end-to-end by the
// golden-file check (HL_pH in tests/testthat/compdata/hector_comp.csv).
#pragma once
#include "../../config.hpp"
#include <cfloat>
#include <cstdint>
#include <stdexcept>
#include <utility>
namespace boost { namespace math { namespace tools {
inline std::uint64_t shim_newton_iterations = 0;   // instrumentation (not part of Boost)
inline std::uint64_t shim_newton_calls = 0;
template <class F, class T>
T newton_raphson_iterate(F f, T guess, T min, T max, int digits) {
  using std::fabs;
  if (min > max) throw std::domain_error("newton_raphson_iterate: min > max");
  T f0(0), f1, last_f0(0);
  T result = guess;
  T factor = static_cast<T>(std::ldexp(1.0, 1 - digits));
  T delta = DBL_MAX, delta1 = DBL_MAX, delta2 = DBL_MAX;
  T max_range_f = 0, min_range_f = 0;
  ++shim_newton_calls;
  do {
    last_f0 = f0;
    delta2 = delta1;
    delta1 = delta;
    std::pair<T, T> fv = f(result);
    f0 = fv.first;
    f1 = fv.second;
    ++shim_newton_iterations;
    if (0 == f0) break;
    if (f1 == 0) {
      // handle_zero_derivative: bisect towards the side the function sign suggests
      if (last_f0 == 0) {
        guess = (result == min) ? max : min;
        last_f0 = f(guess).first;
        delta = guess - result;
      }
      if ((last_f0 < 0 ? -1 : 1) * (f0 < 0 ? -1 : 1) < 0) {
        delta = (delta < 0) ? (result - min) / 2 : (result - max) / 2;
      } else {
        delta = (delta < 0) ? (result - max) / 2 : (result - min) / 2;
      }
    } else {
      delta = f0 / f1;
    }
    if (fabs(delta * 2) > fabs(delta2)) {
      // last two steps have not converged: damped / bisection step
      T shift = (delta > 0) ? (result - min) / 2 : (result - max) / 2;
      if ((result != 0) && (fabs(shift) > fabs(result))) {
        delta = (delta > 0 ? 1 : (delta < 0 ? -1 : 0)) * fabs(result) * 1.1f;
      } else {
        delta = shift;
      }
      delta1 = 3 * delta;
      delta2 = 3 * delta;
    }
    guess = result;
    result -= delta;
    if (result <= min) {
      delta = 0.5F * (guess - min);
      result = guess - delta;
      if ((result == min) || (result == max)) break;
    } else if (result >= max) {
      delta = 0.5F * (guess - max);
      result = guess - delta;
      if ((result == min) || (result == max)) break;
    }
    if (delta > 0) {
      max = guess;
      max_range_f = f0;
    } else {
      min = guess;
      min_range_f = f0;
    }
    if (max_range_f * min_range_f > 0)
      throw std::domain_error("newton_raphson_iterate: no root bracketed");
  } while (fabs(result * factor) < fabs(delta));
  return result;
}
}}} // namespace boost::math::tools
