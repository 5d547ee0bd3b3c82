// oracle/shim: the slice of boost::numeric::odeint used at carbon-cycle-solver.cpp:257-261:
//   integrate_adaptive(make_controlled<runge_kutta_dopri5<std::vector<double>>>(eps_abs, eps_rel),
//                      system, x, t0, t1, dt, observer)
// Restated from the published algorithm: FSAL Dormand-Prince 5(4), default_error_checker
// (a_x = a_dxdt = 1), default_step_adjuster, less_with_sign end test.
#pragma once
#include "../config.hpp"
#include <cfloat>
#include <cstdint>
#include <stdexcept>
namespace boost { namespace numeric { namespace odeint {
inline std::uint64_t shim_rhs_evals = 0;       // instrumentation (not part of Boost)
inline std::uint64_t shim_steps_accepted = 0;
inline std::uint64_t shim_steps_rejected = 0;
inline std::uint64_t shim_integrate_calls = 0;

template <class State> struct runge_kutta_dopri5 { typedef State state_type; };

enum controlled_step_result { success, fail };

template <class Stepper> class controlled_runge_kutta;

template <> class controlled_runge_kutta<runge_kutta_dopri5<std::vector<double>>> {
public:
  typedef std::vector<double> state_type;
  controlled_runge_kutta(double eps_abs, double eps_rel)
      : m_eps_abs(eps_abs), m_eps_rel(eps_rel), m_first_call(true) {}

  template <class System>
  controlled_step_result try_step(System &sys, state_type &x, double &t, double &dt) {
    const std::size_t n = x.size();
    if (m_first_call) {
      m_dxdt.resize(n);
      sys(x, m_dxdt, t); ++shim_rhs_evals;
      m_first_call = false;
    }
    m_xnew.resize(n); m_dxdtnew.resize(n); m_xerr.resize(n);
    do_step(sys, x, m_dxdt, t, m_xnew, m_dxdtnew, dt, m_xerr);
    // default_error_checker::error -> max_i |xerr_i| / (eps_abs + eps_rel*(|x_i| + |dt|*|dxdt_i|))
    double max_rel_err = 0.0;
    const double a_dxdt = 1.0 * std::fabs(dt);
    for (std::size_t i = 0; i < n; ++i) {
      double e = std::fabs(m_xerr[i]) /
                 (m_eps_abs + m_eps_rel * (1.0 * std::fabs(x[i]) + a_dxdt * std::fabs(m_dxdt[i])));
      max_rel_err = std::max(max_rel_err, e);
    }
    if (max_rel_err > 1.0) {
      // decrease_step: error_order = 4
      dt *= std::max(9.0 / 10.0 * std::pow(max_rel_err, -1.0 / (4.0 - 1.0)), 1.0 / 5.0);
      ++shim_steps_rejected;
      return fail;
    }
    t += dt;
    if (max_rel_err < 0.5) {
      // increase_step: stepper_order = 5
      double error = std::max(std::pow(5.0, -5.0), max_rel_err);
      dt *= 9.0 / 10.0 * std::pow(error, -1.0 / 5.0);
    }
    x = m_xnew;
    m_dxdt = m_dxdtnew;
    ++shim_steps_accepted;
    return success;
  }

private:
  template <class System>
  void do_step(System &sys, const state_type &in, const state_type &dxdt_in, double t,
               state_type &out, state_type &dxdt_out, double dt, state_type &xerr) {
    const std::size_t n = in.size();
    const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
    const double b21 = 1.0 / 5.0;
    const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
    const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
    const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0,
                 b54 = -212.0 / 729.0;
    const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0,
                 b64 = 49.0 / 176.0, b65 = -5103.0 / 18656.0;
    const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0,
                 c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
    const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0,
                 dc4 = c4 - 393.0 / 640.0, dc5 = c5 - (-92097.0 / 339200.0),
                 dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;
    m_tmp.resize(n); m_k2.resize(n); m_k3.resize(n); m_k4.resize(n); m_k5.resize(n); m_k6.resize(n);
    double f1, f2, f3, f4, f5, f6;
    f1 = dt * b21;
    for (std::size_t i = 0; i < n; ++i) m_tmp[i] = 1.0 * in[i] + f1 * dxdt_in[i];
    sys(m_tmp, m_k2, t + dt * a2); ++shim_rhs_evals;
    f1 = dt * b31; f2 = dt * b32;
    for (std::size_t i = 0; i < n; ++i) m_tmp[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * m_k2[i];
    sys(m_tmp, m_k3, t + dt * a3); ++shim_rhs_evals;
    f1 = dt * b41; f2 = dt * b42; f3 = dt * b43;
    for (std::size_t i = 0; i < n; ++i)
      m_tmp[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * m_k2[i] + f3 * m_k3[i];
    sys(m_tmp, m_k4, t + dt * a4); ++shim_rhs_evals;
    f1 = dt * b51; f2 = dt * b52; f3 = dt * b53; f4 = dt * b54;
    for (std::size_t i = 0; i < n; ++i)
      m_tmp[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * m_k2[i] + f3 * m_k3[i] + f4 * m_k4[i];
    sys(m_tmp, m_k5, t + dt * a5); ++shim_rhs_evals;
    f1 = dt * b61; f2 = dt * b62; f3 = dt * b63; f4 = dt * b64; f5 = dt * b65;
    for (std::size_t i = 0; i < n; ++i)
      m_tmp[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * m_k2[i] + f3 * m_k3[i] + f4 * m_k4[i] +
                 f5 * m_k5[i];
    sys(m_tmp, m_k6, t + dt); ++shim_rhs_evals;
    f1 = dt * c1; f2 = dt * c3; f3 = dt * c4; f4 = dt * c5; f5 = dt * c6;
    for (std::size_t i = 0; i < n; ++i)
      out[i] = 1.0 * in[i] + f1 * dxdt_in[i] + f2 * m_k3[i] + f3 * m_k4[i] + f4 * m_k5[i] +
               f5 * m_k6[i];
    sys(out, dxdt_out, t + dt); ++shim_rhs_evals;
    f1 = dt * dc1; f2 = dt * dc3; f3 = dt * dc4; f4 = dt * dc5; f5 = dt * dc6; f6 = dt * dc7;
    for (std::size_t i = 0; i < n; ++i)
      xerr[i] = f1 * dxdt_in[i] + f2 * m_k3[i] + f3 * m_k4[i] + f4 * m_k5[i] + f5 * m_k6[i] +
                f6 * dxdt_out[i];
  }
  double m_eps_abs, m_eps_rel;
  bool m_first_call;
  state_type m_dxdt, m_xnew, m_dxdtnew, m_xerr, m_tmp, m_k2, m_k3, m_k4, m_k5, m_k6;
};

template <class Stepper>
controlled_runge_kutta<Stepper> make_controlled(double eps_abs, double eps_rel) {
  return controlled_runge_kutta<Stepper>(eps_abs, eps_rel);
}

// integrate_adaptive for a controlled stepper (stepper taken by value => fresh FSAL state)
template <class Stepper, class System, class State, class Observer>
std::size_t integrate_adaptive(Stepper st, System sys, State &x, double t, double t_end, double dt,
                               Observer obs) {
  ++shim_integrate_calls;
  std::size_t count = 0;
  const double eps = DBL_EPSILON;
  while (t_end - t > eps) {               // less_with_sign(t, t_end, dt), dt > 0
    obs(x, t);
    if ((t + dt) - t_end > eps) dt = t_end - t;
    controlled_step_result res;
    int fails = 0;
    do {
      res = st.try_step(sys, x, t, dt);
      if (res == fail && ++fails >= 500)
        throw std::overflow_error("Max number of iterations exceeded (500). A new step size was not found.");
    } while (res == fail);
    ++count;
  }
  obs(x, t);
  return count;
}
}}} // namespace boost::numeric::odeint
