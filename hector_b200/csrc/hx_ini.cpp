/* hx_ini.cpp -- Hector input formats for the engine: an INI file with `name=value`,
 * `name[year]=value` and `name=csv:table.csv` entries becomes dense per-year scenario series
 * and scalar parameters ("input/wire formats").
 *
 * Behaviour follows the reference readers:
 *   - inih line rules: `;`/`#` comment lines, inline `;` comments only after whitespace,
 *     sections, name=value                                          (src/ini.c:60-127)
 *   - `csv:` paths resolve relative to the ini file if not found as given
 *                                                   (src/ini_to_core_reader.cpp:134-167)
 *   - csv tables: `;` comment lines, header row names the columns, optional UNITS row, blank
 *     cells skipped, CRLF tolerated                          (src/csv_table_reader.cpp:115-196)
 *   - a series evaluated at an integer year = exact key, else the single value if there is
 *     only one, else linear interpolation with flat extrapolation
 *                        (inst/include/tseries.hpp:317-334, src/h_interpolator.cpp:109-125);
 *     ffi/daccs/luc series do not extrapolate (src/simpleNbox.cpp:49-56) and must cover the run.
 * Biomes and spinup_chem=1 are reported as unsupported rather than ignored.
 */
#include <sys/stat.h>

#include <algorithm>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hector_b200.h"
#include "hx_ini.h"
#include "hx_names.h"

namespace hx {

static std::string trim(const std::string &s) {
  size_t b = 0, e = s.size();
  while (b < e && std::isspace((unsigned char)s[b])) ++b;
  while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
  return s.substr(b, e - b);
}

static bool file_exists(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

static std::string dirname_of(const std::string &p) {
  size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

/* inih: cut at c, or at a ';' that follows whitespace */
static size_t find_char_or_comment(const std::string &s, size_t from, char c) {
  bool was_ws = false;
  size_t i = from;
  for (; i < s.size(); ++i) {
    if (s[i] == c) break;
    if (was_ws && s[i] == ';') break;
    was_ws = std::isspace((unsigned char)s[i]) != 0;
  }
  return i;
}

static bool parse_number(const std::string &v, double &out) {
  /* "value" or "value,units" (unitval::parse_unitval) */
  std::string t = trim(v.substr(0, v.find(',')));
  if (t.empty()) return false;
  char *end = nullptr;
  out = std::strtod(t.c_str(), &end);
  return end == t.c_str() + t.size();
}

typedef std::map<double, double> Series;

/* csv_table_reader.cpp:115-196 */
static bool read_csv_column(const std::string &path, const std::string &var, Series &out,
                            std::string &err) {
  std::ifstream f(path.c_str());
  if (!f) {
    err = "cannot open csv table " + path;
    return false;
  }
  std::string line;
  int col = -1;
  bool have_header = false;
  while (std::getline(f, line)) {
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    if (line.empty() || line[0] == ';') continue;
    std::vector<std::string> cells;
    {
      std::string cur;
      for (char ch : line) {
        if (ch == ',') { cells.push_back(cur); cur.clear(); }
        else cur.push_back(ch);
      }
      cells.push_back(cur);
    }
    if (!have_header) {
      for (size_t c = 1; c < cells.size() && col < 0; ++c)
        if (trim(cells[c]) == var) col = (int)c;
      if (col < 0) {
        err = "Could not find a column for " + var + " in " + path;
        return false;
      }
      have_header = true;
      continue;
    }
    if ((int)cells.size() <= col) continue;
    std::string c0 = trim(cells[0]), cv = trim(cells[col]);
    if (c0 == "UNITS" || cv.empty()) continue;
    double date, val;
    if (!parse_number(c0, date)) {
      err = "Could not convert index to double in " + path + ": " + c0;
      return false;
    }
    if (!parse_number(cv, val)) {
      err = "bad value in " + path + " column " + var + ": " + cv;
      return false;
    }
    out[date] = val;
  }
  if (!have_header) {
    err = "empty csv table " + path;
    return false;
  }
  return true;
}

/* tseries::get + h_interpolator::f_linear */
static bool series_at(const Series &s, double t, bool extrapolate, double &out) {
  if (s.empty()) return false;
  if (s.size() == 1) { out = s.begin()->second; return true; }
  Series::const_iterator it = s.find(t);
  if (it != s.end()) { out = it->second; return true; }
  if (t < s.begin()->first) {
    if (!extrapolate) return false;
    out = s.begin()->second;
    return true;
  }
  if (t >= s.rbegin()->first) {
    if (!extrapolate) return false;
    out = s.rbegin()->second;
    return true;
  }
  Series::const_iterator hi = s.upper_bound(t), lo = hi;
  --lo;
  out = lo->second + (t - lo->first) * (hi->second - lo->second) / (hi->first - lo->first);
  return true;
}

static int raw_index(const std::string &name) {
  for (int i = 0; i < RAW_HALO0; ++i)
    if (name == kRawNames[i]) return i;
  for (int g = 0; g < HX_NHALO; ++g)
    if (name == std::string(kHaloNames[g]) + "_emissions") return RAW_HALO0 + g;
  return -1;
}

static const char *const kConstraintNames[CN_HALO0] = {"CO2_constrain", "NBP_constrain",
                                                       "CH4_constrain", "N2O_constrain",
                                                       "RF_tot_constrain", "tas_constrain"};
static int constraint_index(const std::string &name) {
  for (int i = 0; i < CN_HALO0; ++i)
    if (name == kConstraintNames[i]) return i;
  for (int g = 0; g < HX_NHALO; ++g)
    if (name == std::string(kHaloNames[g]) + "_constrain") return CN_HALO0 + g;
  return -1;
}
static std::string constraint_name(int i) {
  return i < CN_HALO0 ? std::string(kConstraintNames[i])
                      : std::string(kHaloNames[i - CN_HALO0]) + "_constrain";
}

static bool is_engine_param(const std::string &name) {
  for (int i = 0; i < PI_COUNT; ++i)
    if (name == kParams[i].name) return true;
  return name == "baseyear" || name == "max_spinup" || name == "UC_N2O" || name == "TN2O0";
}

bool read_ini(const std::string &path, IniInputs &out) {
  std::ifstream f(path.c_str());
  if (!f) {
    out.error = "cannot open ini file " + path;
    return false;
  }
  out.start_year = 0; out.end_year = 0; out.do_spinup = true; out.tracking_date = 9999;
  std::map<int, Series> series, cons;
  const std::string dir = dirname_of(path);
  std::string line, section;
  int lineno = 0;
  while (std::getline(f, line)) {
    ++lineno;
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    std::string s = trim(line);
    if (s.empty() || s[0] == ';' || s[0] == '#') continue;
    if (s[0] == '[') {
      size_t e = find_char_or_comment(s, 1, ']');
      if (e >= s.size() || s[e] != ']') {
        out.error = path + ":" + std::to_string(lineno) + ": no ']' on section line";
        return false;
      }
      section = s.substr(1, e - 1);
      continue;
    }
    size_t eq = find_char_or_comment(s, 0, '=');
    if (eq >= s.size() || s[eq] != '=') {
      out.error = path + ":" + std::to_string(lineno) + ": no '=' on name=value line";
      return false;
    }
    std::string name = trim(s.substr(0, eq));
    std::string rest = s.substr(eq + 1);
    size_t b = 0;
    while (b < rest.size() && std::isspace((unsigned char)rest[b])) ++b;
    size_t ce = find_char_or_comment(rest, b, '\0');
    std::string value = trim(rest.substr(b, ce - b));

    /* name[date] */
    double date = NAN;
    size_t lb = name.find('[');
    if (lb != std::string::npos) {
      size_t rb = name.find(']', lb);
      if (rb == std::string::npos || !parse_number(name.substr(lb + 1, rb - lb - 1), date)) {
        out.error = path + ":" + std::to_string(lineno) + ": bad time-series index in " + name;
        return false;
      }
      name = name.substr(0, lb);
    }
    if (name.find('.') != std::string::npos && section == "simpleNbox") {
      /* <biome>.<name>: a biome-specific pool or parameter (simpleNbox.cpp:201-236) */
      const size_t dot = name.find('.');
      const std::string biome = name.substr(0, dot), var = name.substr(dot + 1);
      bool known = false;
      for (int f = 0; f < BP_COUNT; ++f) known = known || var == kBiomeParams[f].name;
      double v;
      if (!known) {
        out.error = "Unknown variable name while parsing simpleNbox: " + name;
        return false;
      }
      if (!std::isnan(date) || !parse_number(value, v)) {
        out.error = "[simpleNbox] " + name + ": expected a scalar, got: " + value;
        return false;
      }
      if (biome == "global") {
        name = var; /* the default biome spelled out */
      } else {
        if (std::find(out.biomes.begin(), out.biomes.end(), biome) == out.biomes.end())
          out.biomes.push_back(biome);
        out.biome_scalars.push_back(std::make_pair(name, v));
        continue;
      }
    }

    /* ---- [core] ---- */
    if (section == "core") {
      double v = 0;
      if (name == "run_name") { out.run_name = value; continue; }
      if (!parse_number(value, v)) {
        out.error = "[core] " + name + ": not a number: " + value;
        return false;
      }
      if (name == "startDate") out.start_year = (int)v;
      else if (name == "endDate") out.end_year = (int)v;
      else if (name == "do_spinup") out.do_spinup = v != 0;
      else if (name == "max_spinup") out.scalars["max_spinup"] = v;
      else if (name == "trackingDate") out.tracking_date = v;
      else {
        out.error = "Unknown variable name while parsing core: " + name;
        return false;
      }
      continue;
    }
    if (name == "enabled") {
      double v = 1;
      parse_number(value, v);
      if (v == 0) {
        out.error = "[" + section + "] enabled=0: disabling components is not supported";
        out.unsupported = true;
        return false;
      }
      continue;
    }
    if (name == "spinup_chem") {
      double v = 0;
      parse_number(value, v);
      if (v > 0) {
        out.error = "spinup_chem=1 is not supported by the ensemble engine";
        out.unsupported = true;
        return false;
      }
      continue;
    }
    if (name == "atmos_co2") continue; /* overwritten from C0 in prepareToRun (simpleNbox-runtime.cpp:172) */

    /* ---- halocarbon sections: tau, rho_X, delta_X, H0, molarMass, X_emissions ---- */
    std::string pname = name;
    const std::string hsuffix = "_halocarbon";
    if (section.size() > hsuffix.size() &&
        section.compare(section.size() - hsuffix.size(), hsuffix.size(), hsuffix) == 0) {
      const std::string gas = section.substr(0, section.size() - hsuffix.size());
      if (name == "tau" || name == "H0" || name == "molarMass") pname = gas + "." + name;
      else if (name == "rho_" + gas) pname = gas + ".rho";
      else if (name == "delta_" + gas) pname = gas + ".delta";
    }

    /* ---- series ---- */
    const int ri = raw_index(pname);
    if (ri >= 0) {
      Series &sr = series[ri];
      if (value.compare(0, 4, "csv:") == 0) {
        std::string csv = value.substr(4);
        if (!file_exists(csv)) csv = dir + "/" + csv;
        Series tmp;
        if (!read_csv_column(csv, pname, tmp, out.error)) return false;
        sr = tmp; /* a later section naming the same series (e.g. [ozone] NOX_emissions) re-reads it */
      } else {
        double v;
        if (!parse_number(value, v)) {
          out.error = "[" + section + "] " + name + ": not a number: " + value;
          return false;
        }
        sr[std::isnan(date) ? 0.0 : date] = v;
      }
      continue;
    }

    /* ---- user constraints: dated entries or a csv column ---- */
    const int cni = constraint_index(pname);
    if (cni >= 0) {
      Series tmp;
      if (value.compare(0, 4, "csv:") == 0) {
        std::string csv = value.substr(4);
        if (!file_exists(csv)) csv = dir + "/" + csv;
        if (!read_csv_column(csv, pname, tmp, out.error)) return false;
      } else {
        double v;
        if (std::isnan(date) || !parse_number(value, v)) {
          out.error = "[" + section + "] " + name + ": a constraint needs a date and a number";
          return false;
        }
        tmp[date] = v;
      }
      for (Series::const_iterator it = tmp.begin(); it != tmp.end(); ++it) {
        if (it->first != std::floor(it->first)) {
          out.error = "[" + section + "] " + name + ": constraint dates must be whole years";
          out.unsupported = true;
          return false;
        }
        cons[cni][it->first] = it->second;
      }
      continue;
    }

    /* ---- scalars ---- */
    double v;
    if (is_engine_param(pname) || pname.find('.') != std::string::npos) {
      if (!std::isnan(date) || !parse_number(value, v)) {
        out.error = "[" + section + "] " + name + ": expected a scalar, got: " + value;
        return false;
      }
      out.scalars[pname] = v;
      continue;
    }
    if (name.find("constrain") != std::string::npos) {
      out.error = "[" + section + "] " + name + ": not supported by the ensemble engine";
      out.unsupported = true;
      return false;
    }
    out.error = "Unknown variable name while parsing " + section + ": " + name;
    return false;
  }
  if (!out.biomes.empty()) {
    /* simpleNbox-runtime.cpp:66-69 */
    for (int f = 0; f < BP_COUNT; ++f)
      if (out.scalars.count(kBiomeParams[f].name)) {
        out.error = "Cannot have both global and biome-specific data! (" +
                    std::string(kBiomeParams[f].name) + ")";
        return false;
      }
    if (out.biomes.size() < 2 || out.biomes.size() > HX_MAX_BIOMES) {
      out.error = "the ensemble engine runs the global biome or 2 .. " +
                  std::to_string(HX_MAX_BIOMES) + " named biomes";
      out.unsupported = true;
      return false;
    }
  }
  if (out.end_year <= out.start_year) {
    out.error = "[core] startDate/endDate missing or inconsistent";
    return false;
  }
  /* dense per-year tables */
  const int nrow = out.end_year - out.start_year + 1;
  out.constraints.clear();
  for (std::map<int, Series>::const_iterator it = cons.begin(); it != cons.end(); ++it) {
    std::vector<double> dense(nrow, NAN);
    for (Series::const_iterator e = it->second.begin(); e != it->second.end(); ++e) {
      const int r = (int)e->first - out.start_year;
      if (r >= 0 && r < nrow) dense[r] = e->second;
    }
    out.constraints[it->first] = dense;
  }
  out.series.assign(RAW_COUNT, std::vector<double>());
  for (int i = 0; i < RAW_COUNT; ++i) {
    std::string nm = i < RAW_HALO0 ? kRawNames[i] : std::string(kHaloNames[i - RAW_HALO0]) + "_emissions";
    std::map<int, Series>::const_iterator it = series.find(i);
    if (it == series.end() || it->second.empty()) {
      if (i == RAW_ALBEDO) { /* "If no albedo data, assume constant" simpleNbox-runtime.cpp:165-169 */
        out.series[i].assign(nrow, -0.2);
        continue;
      }
      out.error = "input series missing from ini: " + nm;
      return false;
    }
    const bool extrap = !(i == RAW_FFI || i == RAW_DACCS || i == RAW_LUC_E || i == RAW_LUC_U);
    out.series[i].resize(nrow);
    for (int r = 0; r < nrow; ++r) {
      double v;
      if (!series_at(it->second, (double)(out.start_year + r), extrap, v)) {
        if (!extrap && r == nrow - 1) { /* emissions of the last year are never read (t = y-1) */
          out.series[i][r] = 0.0;
          continue;
        }
        out.error = "Interpolation requested but not allowed (" + nm + ") date: " +
                    std::to_string(out.start_year + r);
        return false;
      }
      out.series[i][r] = v;
    }
  }
  return true;
}

} // namespace hx

extern "C" int hx_create_from_ini(const char *const *ini_paths, int32_t n_inis, int32_t n_members,
                                  int32_t device, uint32_t flags, hx_handle *out) {
  if (!ini_paths || n_inis <= 0 || !out) return HX_ERR_ARG;
  *out = nullptr;
  std::vector<hx::IniInputs> in(n_inis);
  for (int s = 0; s < n_inis; ++s) {
    if (!hx::read_ini(ini_paths[s], in[s])) {
      hx_set_create_error((std::string(ini_paths[s]) + ": " + in[s].error).c_str());
      return in[s].unsupported ? HX_ERR_UNSUPPORTED : HX_ERR_ARG;
    }
    if (in[s].start_year != in[0].start_year || in[s].end_year != in[0].end_year) {
      hx_set_create_error("all ini files of one engine must share startDate/endDate");
      return HX_ERR_ARG;
    }
  }
  hx_config cfg;
  cfg.n_members = n_members; cfg.n_scenarios = n_inis;
  cfg.start_year = in[0].start_year; cfg.end_year = in[0].end_year;
  cfg.device = device;
  cfg.flags = flags | (in[0].do_spinup ? 0u : HX_FLAG_NO_SPINUP);
  hx_handle h = nullptr;
  int rc = hx_create(&cfg, &h);
  if (rc) return rc;
  auto bail = [&](int code) {
    hx_set_create_error(hx_last_error(h));
    hx_destroy(h);
    return code;
  };
  if (in[0].tracking_date < 9999) {
    rc = hx_set_tracking(h, (int32_t)in[0].tracking_date, 1);
    if (rc) return bail(rc);
  }
  if (!in[0].biomes.empty()) {
    std::vector<const char *> names;
    for (const std::string &b : in[0].biomes) names.push_back(b.c_str());
    rc = hx_set_biomes(h, (int32_t)names.size(), names.data());
    if (rc) return bail(rc);
    for (const std::pair<std::string, double> &kv : in[0].biome_scalars) {
      rc = hx_set_param_scalar(h, kv.first.c_str(), kv.second);
      if (rc) return bail(rc);
    }
  }
  for (std::map<std::string, double>::const_iterator it = in[0].scalars.begin();
       it != in[0].scalars.end(); ++it) {
    rc = hx_set_param_scalar(h, it->first.c_str(), it->second);
    if (rc) return bail(rc);
  }
  const int nrow = cfg.end_year - cfg.start_year + 1;
  for (int s = 0; s < n_inis; ++s)
    for (int i = 0; i < RAW_COUNT; ++i) {
      std::string nm = i < RAW_HALO0 ? hx::kRawNames[i]
                                     : std::string(hx::kHaloNames[i - RAW_HALO0]) + "_emissions";
      rc = hx_set_scenario_series(h, s, nm.c_str(), cfg.start_year, nrow, in[s].series[i].data());
      if (rc) return bail(rc);
    }
  for (int s = 0; s < n_inis; ++s)
    for (std::map<int, std::vector<double> >::const_iterator it = in[s].constraints.begin();
         it != in[s].constraints.end(); ++it) {
      rc = hx_set_scenario_series(h, s, hx::constraint_name(it->first).c_str(), cfg.start_year,
                                  nrow, it->second.data());
      if (rc) return bail(rc);
    }
  *out = h;
  return HX_OK;
}

/* host-only view of the reader (no device needed): dates, dense table [nrow][RAW_COUNT] in
 * RAW_* order, and scalar lookups */
extern "C" int hx_ini_read(const char *ini_path, int32_t *start_year, int32_t *end_year,
                           double *table, int32_t table_rows) {
  if (!ini_path) return HX_ERR_ARG;
  hx::IniInputs in;
  if (!hx::read_ini(ini_path, in)) {
    hx_set_create_error((std::string(ini_path) + ": " + in.error).c_str());
    return in.unsupported ? HX_ERR_UNSUPPORTED : HX_ERR_ARG;
  }
  if (start_year) *start_year = in.start_year;
  if (end_year) *end_year = in.end_year;
  const int nrow = in.end_year - in.start_year + 1;
  if (table) {
    if (table_rows < nrow) return HX_ERR_ARG;
    for (int r = 0; r < nrow; ++r)
      for (int i = 0; i < RAW_COUNT; ++i) table[(size_t)r * RAW_COUNT + i] = in.series[i][r];
  }
  return HX_OK;
}

extern "C" int hx_ini_scalar(const char *ini_path, const char *name, double *out) {
  if (!ini_path || !name || !out) return HX_ERR_ARG;
  hx::IniInputs in;
  if (!hx::read_ini(ini_path, in)) {
    hx_set_create_error((std::string(ini_path) + ": " + in.error).c_str());
    return in.unsupported ? HX_ERR_UNSUPPORTED : HX_ERR_ARG;
  }
  if (!strcmp(name, "trackingDate")) { *out = in.tracking_date; return HX_OK; }
  if (!strcmp(name, "do_spinup")) { *out = in.do_spinup ? 1.0 : 0.0; return HX_OK; }
  std::map<std::string, double>::const_iterator it = in.scalars.find(name);
  if (it == in.scalars.end()) return HX_ERR_ARG;
  *out = it->second;
  return HX_OK;
}

extern "C" int hx_ini_string(const char *ini_path, const char *name, char *buf, int32_t cap) {
  if (!ini_path || !name || !buf || cap < 1) return HX_ERR_ARG;
  hx::IniInputs in;
  if (!hx::read_ini(ini_path, in)) {
    hx_set_create_error((std::string(ini_path) + ": " + in.error).c_str());
    return in.unsupported ? HX_ERR_UNSUPPORTED : HX_ERR_ARG;
  }
  if (strcmp(name, "run_name")) return HX_ERR_ARG; /* [core] run_name is the one string input */
  snprintf(buf, (size_t)cap, "%s", in.run_name.c_str());
  return HX_OK;
}
