// oracle/shim: boost::math::tools::brent_find_minima (oceanbox.cpp:437-438).  Restated from
// the published Boost algorithm (Brent 1973 with Boost's tolerance handling: bits clamped
// to digits/2 = 26, golden ratio as a float literal).
#pragma once
#include "../../config.hpp"
#include <cstdint>
#include <utility>
namespace boost { namespace math { namespace tools {
template <class F, class T>
std::pair<T, T> brent_find_minima(F f, T min, T max, int bits) {
  using std::fabs;
  bits = (std::min)(53 / 2, bits);
  T tolerance = static_cast<T>(std::ldexp(1.0, 1 - bits));
  T x, w, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2;
  static const T golden = 0.3819660f;
  x = w = v = max;
  fw = fv = fx = f(x);
  delta2 = delta = 0;
  std::uintmax_t count = UINTMAX_MAX;
  do {
    mid = (min + max) / 2;
    fract1 = tolerance * fabs(x) + tolerance / 4;
    fract2 = 2 * fract1;
    if (fabs(x - mid) <= (fract2 - (max - min) / 2)) break;
    if (fabs(delta2) > fract1) {
      T r = (x - w) * (fx - fv);
      T q = (x - v) * (fx - fw);
      T p = (x - v) * q - (x - w) * r;
      q = 2 * (q - r);
      if (q > 0) p = -p;
      q = fabs(q);
      T td = delta2;
      delta2 = delta;
      if ((fabs(p) >= fabs(q * td / 2)) || (p <= q * (min - x)) || (p >= q * (max - x))) {
        delta2 = (x >= mid) ? min - x : max - x;
        delta = golden * delta2;
      } else {
        delta = p / q;
        u = x + delta;
        if (((u - min) < fract2) || ((max - u) < fract2))
          delta = (mid - x) < 0 ? (T)-fabs(fract1) : (T)fabs(fract1);
      }
    } else {
      delta2 = (x >= mid) ? min - x : max - x;
      delta = golden * delta2;
    }
    u = (fabs(delta) >= fract1) ? T(x + delta)
                                : (delta > 0 ? T(x + fabs(fract1)) : T(x - fabs(fract1)));
    fu = f(u);
    if (fu <= fx) {
      if (u >= x) min = x; else max = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) min = u; else max = u;
      if ((fu <= fw) || (w == x)) {
        v = w; w = u;
        fv = fw; fw = fu;
      } else if ((fu <= fv) || (v == x) || (v == w)) {
        v = u;
        fv = fu;
      }
    }
  } while (--count);
  return std::make_pair(x, fx);
}
}}} // namespace boost::math::tools
