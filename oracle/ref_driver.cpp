/* oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C ABI over the UNMODIFIED reference C++ (compiled from /root/reference/src by
 * oracle/Makefile against oracle/shim).  It drives the reference exactly the way its own
 * front ends do: Core::mkcore -> init -> INIToCoreReader::parse -> setData/sendMessage ->
 * prepareToRun -> run -> sendMessage(M_GETDATA)  (cf. src/main.cpp:41-112,
 * src/rcpp_hector.cpp:31-365).  Used (a) to pin the C restatement (oracle/hector_oracle.c)
 * and the CUDA engine, (b) as the CPU baseline ("kind": "reference") in bench.py.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

/* The tracked pools (fluxpool source maps) are private members read only by the friend class
 * CSVFluxPoolVisitor, whose CSV carries six significant digits.  To pin the tracking arithmetic
 * at full precision this TEST driver -- and only this translation unit; the reference objects
 * are compiled unmodified -- reads them directly.  Access specifiers do not change layout. */
#define private public
#define protected public
#include "hector.hpp"
#include "ocean_component.hpp"
#include "simpleNbox.hpp"
#undef private
#undef protected
#include "csv_outputstream_visitor.hpp"
#include "imodel_component.hpp"
#include "ini_to_core_reader.hpp"
#include "logger.hpp"
#include "message_data.hpp"
#include "unitval.hpp"
#include <boost/math/tools/roots.hpp>
#include <boost/numeric/odeint.hpp>

using namespace Hector;

static thread_local std::string g_err;

static int fail(const std::string &m) {
  g_err = m;
  return -1;
}

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

/* create core, init components, parse ini. Returns handle >= 0 or -1. */
int ref_open(const char *ini_path) {
  try {
    int idx = Core::mkcore(false, Logger::SEVERE, false);
    Core *core = Core::getcore(idx);
    core->init();
    INIToCoreReader reader(core);
    reader.parse(ini_path);
    return idx;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

/* like an extra ini line  [component] var=value  or var[date]=value (date < 0 => undated);
 * only valid before ref_prepare */
int ref_setdata(int h, const char *component, const char *var, double date, const char *value) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    message_data d{std::string(value)};
    d.date = date < 0 ? Core::undefinedIndex() : date;
    core->setData(component, var, d);
    return 0;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

/* R-style setvar: sendMessage(M_SETDATA, var, (date, unitval(value, unit))) */
int ref_setvar(int h, const char *var, double date, double value, const char *unit) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    unitval v(value, unitval::parseUnitsName(unit));
    message_data d(date < 0 ? Core::undefinedIndex() : date, v);
    core->sendMessage(M_SETDATA, var, d);
    return 0;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

int ref_prepare(int h) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    core->prepareToRun();
    return 0;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

int ref_run(int h, double to_date) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    core->run(to_date);
    return 0;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

int ref_reset(int h, double date) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    core->reset(date);
    return 0;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

/* GETDATA through the core's capability routing; date < 0 => undated */
int ref_fetch(int h, const char *var, double date, double *out) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    unitval v = core->sendMessage(M_GETDATA, var,
                                  message_data(date < 0 ? Core::undefinedIndex() : date));
    *out = v.value(v.units());
    return 0;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

int ref_fetch_series(int h, const char *var, double date0, int n, double *out) {
  for (int i = 0; i < n; ++i) {
    int rc = ref_fetch(h, var, date0 + i, out + i);
    if (rc) return rc;
  }
  return 0;
}

/* GETDATA sent straight to a named component (for un-registered data such as the ocean's
 * "ocean_timesteps", ocean_component.cpp:508) */
int ref_fetch_component(int h, const char *component, const char *var, double date, double *out) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    IModelComponent *c = core->getComponentByName(component);
    unitval v = c->sendMessage(M_GETDATA, var,
                               message_data(date < 0 ? Core::undefinedIndex() : date));
    *out = v.value(v.units());
    return 0;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

/* Core::getTrackingData() (core.cpp:199-209): the CSVFluxPoolVisitor's buffered rows
 * "year,component,pool_name,pool_value,pool_units,source_name,source_fraction" for every year
 * >= trackingDate.  Copies at most cap-1 bytes; returns the full length (call again with a
 * larger buffer if it is >= cap), or -1. */
long ref_tracking_data(int h, char *buf, long cap) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    const std::string s = core->getTrackingData();
    if (buf && cap > 0) {
      const long n = (long)s.size() < cap - 1 ? (long)s.size() : cap - 1;
      memcpy(buf, s.data(), (size_t)n);
      buf[n] = 0;
    }
    return (long)s.size();
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

/* Full-precision view of one tracked pool at the core's current date.  pool / source names:
 * atmos_co2 earth_c veg_c detritus_c soil_c permafrost_c thawedp_c HL LL intermediate deep
 * (+ source "untracked").  frac[i] / present[i] follow the order of `sources` (n of them);
 * a source missing from the pool's map reports present 0, frac 0.  Returns 1 if the pool is
 * tracking, 0 if not, -1 on error. */
int ref_tracking_pool(int h, const char *pool, int n, const char **sources, double *value,
                      double *frac, int *present) {
  try {
    Core *core = Core::getcore(h);
    if (!core) return fail("bad handle");
    SimpleNbox *nb = dynamic_cast<SimpleNbox *>(core->getComponentByName(SIMPLENBOX_COMPONENT_NAME));
    OceanComponent *oc = dynamic_cast<OceanComponent *>(core->getComponentByName(OCEAN_COMPONENT_NAME));
    if (!nb || !oc) return fail("components not found");
    const std::string p(pool);
    fluxpool x;
    if (p == "atmos_co2") x = nb->atmos_c;
    else if (p == "earth_c") x = nb->earth_c;
    else if (p == "veg_c") x = nb->veg_c.at(SNBOX_DEFAULT_BIOME);
    else if (p == "detritus_c") x = nb->detritus_c.at(SNBOX_DEFAULT_BIOME);
    else if (p == "soil_c") x = nb->soil_c.at(SNBOX_DEFAULT_BIOME);
    else if (p == "permafrost_c") x = nb->permafrost_c.at(SNBOX_DEFAULT_BIOME);
    else if (p == "thawedp_c") x = nb->thawed_permafrost_c.at(SNBOX_DEFAULT_BIOME);
    else if (p == "HL") x = oc->surfaceHL.get_carbon();
    else if (p == "LL") x = oc->surfaceLL.get_carbon();
    else if (p == "intermediate") x = oc->inter.get_carbon();
    else if (p == "deep") x = oc->deep.get_carbon();
    else return fail("unknown pool " + p);
    *value = x.value(U_PGC);
    const std::unordered_map<std::string, double> m = x.get_tracking_map();
    for (int i = 0; i < n; ++i) {
      auto it = m.find(sources[i]);
      present[i] = it != m.end();
      frac[i] = present[i] ? it->second : 0.0;
    }
    return x.tracking ? 1 : 0;
  } catch (h_exception &e) {
    return fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    return fail(std::string("exception: ") + e.what());
  }
}

double ref_start_date(int h) { Core *c = Core::getcore(h); return c ? c->getStartDate() : -1; }
double ref_end_date(int h) { Core *c = Core::getcore(h); return c ? c->getEndDate() : -1; }
double ref_current_date(int h) { Core *c = Core::getcore(h); return c ? c->getCurrentDate() : -1; }

void ref_close(int h) {
  try {
    Core *core = Core::getcore(h);
    if (core) {
      core->shutDown();
      Core::delcore(h);
    }
  } catch (...) {
  }
}

/* shim instrumentation: [rhs_evals, steps_accepted, steps_rejected, integrate_calls,
 * newton_iterations, newton_calls]; reset != 0 zeroes them afterwards */
void ref_counters(uint64_t out[6], int reset) {
  using namespace boost::numeric::odeint;
  using namespace boost::math::tools;
  out[0] = shim_rhs_evals; out[1] = shim_steps_accepted; out[2] = shim_steps_rejected;
  out[3] = shim_integrate_calls; out[4] = shim_newton_iterations; out[5] = shim_newton_calls;
  if (reset) {
    shim_rhs_evals = shim_steps_accepted = shim_steps_rejected = shim_integrate_calls = 0;
    shim_newton_iterations = shim_newton_calls = 0;
  }
}

/* One whole member: open ini, apply n overrides (component/var/value as strings are packed
 * by the caller), prepare (spin-up), run year by year to `to_date` and record `nvars`
 * variables per year plus the ocean sub-step count.  out is [nvars+1][nyears] with year
 * index 0 = start date + 1.  Returns 0, or -1 with ref_last_error() set (out beyond the
 * failing year is left untouched).  This is the loop bench.py times as the CPU baseline. */
int ref_run_member(const char *ini_path, int n_over, const char **comps, const char **vars_over,
                   const double *values, int nvars, const char **vars, double to_date,
                   double *out, int nyears_cap, double *run_seconds) {
  int h = ref_open(ini_path);
  if (h < 0) return -1;
  int rc = 0;
  char buf[64];
  for (int i = 0; i < n_over && !rc; ++i) {
    snprintf(buf, sizeof buf, "%.17g", values[i]);
    rc = ref_setdata(h, comps[i], vars_over[i], -1, buf);
  }
  if (!rc) rc = ref_prepare(h);
  if (!rc) {
    Core *core = Core::getcore(h);
    double y0 = core->getStartDate();
    double y1 = to_date < 0 ? core->getEndDate() : to_date;
    int ny = (int)(y1 - y0);
    if (ny > nyears_cap) ny = nyears_cap;
    struct timespec t0, t1;
    double secs = 0;
    if (out == nullptr || nvars == 0) {
      clock_gettime(CLOCK_MONOTONIC, &t0);
      rc = ref_run(h, y0 + ny);
      clock_gettime(CLOCK_MONOTONIC, &t1);
      secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    } else {
      for (int i = 0; i < ny && !rc; ++i) {
        clock_gettime(CLOCK_MONOTONIC, &t0);
        rc = ref_run(h, y0 + i + 1);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        secs += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        for (int v = 0; v < nvars && !rc; ++v)
          rc = ref_fetch(h, vars[v], y0 + i + 1, out + (size_t)v * nyears_cap + i);
        if (!rc) rc = ref_fetch_component(h, "ocean", "ocean_timesteps", -1,
                                          out + (size_t)nvars * nyears_cap + i);
      }
    }
    if (run_seconds) *run_seconds = secs;
  }
  std::string keep = g_err;
  ref_close(h);
  g_err = keep;
  return rc;
}

/* The reference's outputstream_<run>.csv (src/main.cpp:91-106: a CSVOutputStreamVisitor added
 * to the core before prepareToRun) for a run from the ini's start date to `to_date`, written to
 * `path`.  The text the compat visitor of include/compat is held to. */
int ref_outputstream(const char *ini_path, double to_date, const char *path) {
  int h = ref_open(ini_path);
  if (h < 0) return -1;
  int rc = 0;
  try {
    Core *core = Core::getcore(h);
    std::ofstream f(path);
    CSVOutputStreamVisitor visitor(f);
    core->addVisitor(&visitor);
    core->prepareToRun();
    core->run(to_date < 0 ? core->getEndDate() : to_date);
  } catch (h_exception &e) {
    rc = fail(std::string("h_exception: ") + e.what());
  } catch (std::exception &e) {
    rc = fail(std::string("exception: ") + e.what());
  }
  std::string keep = g_err;
  ref_close(h);
  g_err = keep;
  return rc;
}

} // extern "C"
