/* include/hector_b200_core.hpp -- C++ host facade over the C ABI (hector_b200.h).
 *
 * Header only; link with -lhector_b200.  It re-states, for the accelerated path, the surface that
 * the reference's wrappers are written against (paths relative to JGCRI/hector v3.5.0):
 *
 *   hector_b200::Core            Hector::Core              inst/include/core.hpp:37-114
 *     Core(loglvl, toscreen, tofile), init(), setData(component, var, message_data),
 *     prepareToRun(), run(runtodate), reset(resetdate), shutDown(),
 *     sendMessage(msg, datum[, message_data]) -> unitval, getStartDate/EndDate/CurrentDate,
 *     statics mkcore / getcore / delcore (src/core.cpp:817-857)
 *   hector_b200::INIToCoreReader Hector::INIToCoreReader   inst/include/ini_to_core_reader.hpp
 *   hector_b200::message_data    Hector::message_data      inst/include/message_data.hpp:29-93
 *   hector_b200::unitval         Hector::unitval           inst/include/unitval.hpp:132-190
 *   hector_b200::h_exception     Hector::h_exception       inst/include/h_exception.hpp:26-102
 *   M_GETDATA / M_SETDATA        component_data.hpp:409-412
 *
 * plus the batch entry the reference does not have:
 *
 *   hector_b200::EnsembleCore    M members behind the same messages; every per-member call takes
 *                                a member index, every batch call a vector / caller buffer.
 *
 * `Core` is an EnsembleCore with one member, so code written against Hector::Core (the R glue
 * src/rcpp_hector.cpp:31-365, src/main.cpp:41-112, misc/main-api.cpp:68-149) compiles against it
 * with `namespace Hector = hector_b200;` (define HECTOR_B200_AS_HECTOR before including).
 * Errors surface as h_exception exactly where the reference throws: unknown variables and unit
 * mismatches in sendMessage / setData, dates outside the run in getData, and model failures
 * ("Flux and pool values may not be negative", "Mass not conserved", ...) from run().
 */
#ifndef HECTOR_B200_CORE_HPP
#define HECTOR_B200_CORE_HPP

#include <cmath>
#include <cstring>
#include <exception>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "hector_b200.h"

namespace hector_b200 {

#define HXB_M_GETDATA "getData"
#define HXB_M_SETDATA "setData"
#ifndef M_GETDATA
#define M_GETDATA HXB_M_GETDATA
#define M_SETDATA HXB_M_SETDATA
#endif

/* ---- h_exception (h_exception.hpp:26-102): message + where it was raised ---- */
class h_exception : public std::exception {
  std::string msg_, func_, file_;
  int line_;
  mutable std::string full_;

 public:
  h_exception(const std::string &msg, const std::string &func, const std::string &file, int line)
      : msg_(msg), func_(func), file_(file), line_(line) {}
  const char *what() const noexcept override { return msg_.c_str(); }
  const std::string &func() const { return func_; }
  const std::string &file() const { return file_; }
  int line() const { return line_; }
  friend std::ostream &operator<<(std::ostream &os, const h_exception &e) {
    return os << "msg:\t" << e.msg_ << "\nfunc:\t" << e.func_ << "\nfile:\t" << e.file_
              << "\nline:\t" << e.line_ << "\n";
  }
};
#define HXB_THROW(m) throw ::hector_b200::h_exception((m), __func__, __FILE__, __LINE__)

/* ---- units (unitval.hpp:68-130): the ones variables of this path carry ---- */
enum unit_types {
  U_UNITLESS, U_PPMV_CO2, U_PPBV_CH4, U_PPBV_N2O, U_DU_O3, U_TG_PPBV, U_DEGC, U_CM2_S, U_PGC,
  U_PGC_YR, U_W_M2, U_W_M2_TG, U_W_M2_GG, U_M3_S, U_PH, U_UATM, U_YRS, U_PPTV, U_UNDEFINED
};

class unitval {
  double val_;
  unit_types units_;

 public:
  unitval() : val_(0.0), units_(U_UNDEFINED) {}
  unitval(double v, unit_types u) : val_(v), units_(u) {}
  static std::string unitsName(unit_types u) {
    static const char *const names[] = {"(unitless)", "ppmv CO2", "ppbv CH4", "ppbv N2O", "DU O3",
                                        "Tg/ppbv", "degC", "cm2/s", "Pg C", "Pg C/yr", "W/m2",
                                        "W/m2/Tg", "W/m2/Gg", "m3/s", "pH", "uatm", "Years",
                                        "pptv", "(undefined)"};
    return names[(int)u];
  }
  static unit_types parseUnitsName(const std::string &s) {
    for (int u = 0; u <= (int)U_UNDEFINED; ++u)
      if (unitsName((unit_types)u) == s) return (unit_types)u;
    HXB_THROW("Couldn't parse unknown units: " + s);
  }
  /* value(u) insists on the unit, like the reference (unitval.hpp:205-210) */
  double value(unit_types u) const {
    if (u != units_)
      HXB_THROW("variable is not in the expected units: " + unitsName(units_) + " vs " + unitsName(u));
    return val_;
  }
  unit_types units() const { return units_; }
  std::string unitsName() const { return unitsName(units_); }
  void expecting_unit(unit_types u) {
    if (units_ == U_UNDEFINED) units_ = u;
    else if (units_ != u)
      HXB_THROW("Units: " + unitsName() + " do not match expected: " + unitsName(u));
  }
  operator double() const { return val_; }
};

/* ---- message_data (message_data.hpp:29-93) ---- */
struct message_data {
  message_data() : date(undefined()), isVal(false) {}
  message_data(double d) : date(d), isVal(false) {}
  message_data(const std::string &value) : date(undefined()), value_str(value), isVal(false) {}
  message_data(const unitval &value) : date(undefined()), value_unitval(value), isVal(true) {}
  message_data(double d, const unitval &value) : date(d), value_unitval(value), isVal(true) {}
  static double undefined() { return -1.0; } /* Core::undefinedIndex(), core.cpp:806-808 */
  unitval getUnitval(unit_types expected, bool strict = false) const {
    unitval r;
    if (isVal) {
      r = value_unitval;
    } else {
      char *end = nullptr;
      const double v = std::strtod(value_str.c_str(), &end);
      if (end == value_str.c_str()) HXB_THROW("Could not convert '" + value_str + "' to a number");
      r = unitval(v, units_str.empty() ? U_UNDEFINED : unitval::parseUnitsName(units_str));
    }
    if (strict && r.units() != expected)
      HXB_THROW("Units: " + r.unitsName() + " do not match expected: " + unitval::unitsName(expected));
    r.expecting_unit(expected);
    return r;
  }
  double date;
  std::string value_str;
  unitval value_unitval;
  std::string units_str;
  bool isVal;
};

struct Logger { /* log levels only: the engine does not log (logger.hpp:47-54) */
  enum LogLevel { DEBUG, NOTICE, WARNING, SEVERE };
};

namespace detail {
struct VarUnit {
  const char *name;
  unit_types units;
};
/* units of the variables and parameters of this path (component_data.hpp + each component's
 * getData / setData) */
inline unit_types units_of(const std::string &v) {
  static const VarUnit tab[] = {
      {"CO2_concentration", U_PPMV_CO2}, {"global_tas", U_DEGC}, {"land_tas", U_DEGC},
      {"sst", U_DEGC}, {"RF_tot", U_W_M2}, {"RF_CO2", U_W_M2}, {"RF_CH4", U_W_M2},
      {"RF_N2O", U_W_M2}, {"heatflux", U_W_M2}, {"ocean_c", U_PGC}, {"atmos_co2", U_PGC},
      {"permafrost_c", U_PGC}, {"veg_c", U_PGC}, {"detritus_c", U_PGC}, {"soil_c", U_PGC},
      {"thawedp_c", U_PGC}, {"earth_c", U_PGC}, {"HL_ocean_c", U_PGC}, {"LL_ocean_c", U_PGC},
      {"IO_ocean_c", U_PGC}, {"DO_ocean_c", U_PGC}, {"NBP", U_PGC_YR}, {"ocean_uptake", U_PGC_YR},
      {"rh_ch4", U_PGC_YR}, {"HL_pH", U_PH}, {"LL_pH", U_PH}, {"HL_PCO2", U_UATM},
      {"LL_PCO2", U_UATM}, {"CH4_concentration", U_PPBV_CH4}, {"N2O_concentration", U_PPBV_N2O},
      {"O3_concentration", U_DU_O3}, {"ocean_timesteps", U_UNITLESS}, {"NPP", U_PGC_YR},
      {"RH", U_PGC_YR}, {"gmst", U_DEGC}, {"ocean_tas", U_DEGC}, {"heatflux_mixed", U_W_M2},
      {"heatflux_interior", U_W_M2},
      /* parameters */
      {"S", U_DEGC}, {"diff", U_CM2_S}, {"qco2", U_W_M2}, {"beta", U_UNITLESS},
      {"q10_rh", U_UNITLESS}, {"f_nppv", U_UNITLESS}, {"f_nppd", U_UNITLESS},
      {"f_litterd", U_UNITLESS}, {"npp_flux0", U_PGC_YR}, {"C0", U_PPMV_CO2},
      {"warmingfactor", U_UNITLESS}, {"rh_ch4_frac", U_UNITLESS}, {"pf_mu", U_UNITLESS},
      {"pf_sigma", U_UNITLESS}, {"fpf_static", U_UNITLESS}, {"tt", U_M3_S}, {"tu", U_M3_S},
      {"twi", U_M3_S}, {"tid", U_M3_S}, {"preind_surface_c", U_PGC}, {"preind_interdeep_c", U_PGC},
      {"aero_scalar", U_UNITLESS}, {"vol_scalar", U_UNITLESS}, {"delta_co2", U_UNITLESS},
      {"delta_ch4", U_UNITLESS}, {"delta_n2o", U_UNITLESS}, {"rho_bc", U_W_M2_TG},
      {"rho_oc", U_W_M2_TG}, {"rho_so2", U_W_M2_GG}, {"rho_nh3", U_W_M2_TG}, {"M0", U_PPBV_CH4},
      {"N0", U_PPBV_N2O}, {"Tsoil", U_YRS}, {"Tstrat", U_YRS}, {"TOH0", U_YRS},
      {"UC_CH4", U_TG_PPBV}, {"CNOX", U_UNITLESS}, {"CCO", U_UNITLESS}, {"CNMVOC", U_UNITLESS},
      {"CCH4", U_UNITLESS}, {"lo_warming_ratio", U_UNITLESS},
      /* user constraints (component_data.hpp:46, 263, 273, 379-381) and [core] trackingDate */
      {"CO2_constrain", U_PPMV_CO2}, {"tas_constrain", U_DEGC}, {"RF_tot_constrain", U_W_M2},
      {"CH4_constrain", U_PPBV_CH4}, {"N2O_constrain", U_PPBV_N2O}, {"NBP_constrain", U_PGC_YR},
      {"trackingDate", U_UNITLESS}};
  for (const VarUnit &e : tab)
    if (v == e.name) return e.units;
  /* <biome>.<name> (core.cpp:718-727): the units of <name>; "<gas>.tau" etc. are not in `tab` */
  const size_t dot = v.find('.');
  if (dot != std::string::npos)
    for (const VarUnit &e : tab)
      if (v.compare(dot + 1, std::string::npos, e.name) == 0) return e.units;
  const std::string suffix = "_constrain"; /* <gas>_constrain: halocarbon concentrations */
  if (v.size() > suffix.size() && v.compare(v.size() - suffix.size(), suffix.size(), suffix) == 0)
    return U_PPTV;
  /* outputs derived at fetch time: per-agent forcings, halocarbon forcings and concentrations */
  if (v.compare(0, 3, "RF_") == 0 || v.compare(0, 4, "Fadj") == 0) return U_W_M2;
  const std::string csuffix = "_concentration";
  if (v.size() > csuffix.size() && v.compare(v.size() - csuffix.size(), csuffix.size(), csuffix) == 0)
    return U_PPTV;
  return U_UNDEFINED;
}
inline const char *member_failure(int status) { /* the reference's exception text */
  switch (status) {
    case HX_MEMBER_NEGATIVE: return "Flux and pool values may not be negative";
    case HX_MEMBER_MASS: return "Mass not conserved in simpleNbox";
    case HX_MEMBER_RETRIES: return "solver failure: t != tnew";
    case HX_MEMBER_NOROOT: return "ocean_csys: no root found for [H+]";
    case HX_MEMBER_YEARFRACTION: return "yearfraction out of bounds";
    case HX_MEMBER_CO2SARF: return "CO2 SARF could not be calculated";
    case HX_MEMBER_STEPPER: return "Max number of iterations exceeded in odeint";
    case HX_MEMBER_SPINUP: return "spin-up did not converge";
    case HX_MEMBER_TRACKING: return "fractions must be 0-1";
    default: return "model failure";
  }
}
} // namespace detail

/* ---- M members behind Hector's message surface ---- */
class EnsembleCore {
 public:
  /* all 32 recorded variables unless `outputs` names a subset (fewer outputs = less HBM traffic) */
  explicit EnsembleCore(int n_members = 1, int device = 0, unsigned flags = 0)
      : n_(n_members), device_(device), flags_(flags) {}
  EnsembleCore(const EnsembleCore &) = delete;
  EnsembleCore &operator=(const EnsembleCore &) = delete;
  virtual ~EnsembleCore() { shutDown(); }

  void init() {} /* components are created with the engine (core.cpp:90-196) */

  /* INIToCoreReader::parse: one ini file per scenario */
  void parse(const std::vector<std::string> &ini_files) {
    shutDown();
    std::vector<const char *> p;
    for (const std::string &s : ini_files) p.push_back(s.c_str());
    if (hx_create_from_ini(p.data(), (int32_t)p.size(), n_, device_, flags_, &h_) != HX_OK)
      HXB_THROW(std::string(hx_last_error(nullptr)));
    prepared_ = false;
  }
  void selectOutputs(const std::vector<std::string> &names) {
    need();
    std::vector<const char *> p;
    for (const std::string &s : names) p.push_back(s.c_str());
    chk(hx_select_outputs(h_, (int32_t)p.size(), p.data()));
    outputs_selected_ = true;
  }
  void setMemberScenario(const std::vector<int32_t> &scenario_of_member) {
    need();
    chk(hx_set_member_scenario(h_, scenario_of_member.data(), (int32_t)scenario_of_member.size()));
  }

  /* Biomes: the reference grows biome_list as "<biome>.<name>" data arrive (simpleNbox.cpp:
   * 229-236); here the list is declared once, in creation order, before any such setData
   * (an ini file with <biome>.<name> lines declares it by itself).  getBiomeList: core.cpp:
   * 560-563. */
  void setBiomes(const std::vector<std::string> &names) {
    need();
    std::vector<const char *> p;
    for (const std::string &s : names) p.push_back(s.c_str());
    chk(hx_set_biomes(h_, (int32_t)p.size(), p.data()));
    biomes_ = names;
  }
  std::vector<std::string> getBiomeList() const {
    return biomes_.empty() ? std::vector<std::string>(1, "global") : biomes_;
  }

  /* Core::setData (core.cpp:219-268): the component name only routes in the reference */
  void setData(const std::string & /*componentName*/, const std::string &varName,
               const message_data &data) {
    need();
    const unit_types want = detail::units_of(varName);
    const unitval v = data.getUnitval(want);
    if (varName == "trackingDate") { /* [core] trackingDate, core.cpp:228-235 */
      chk(hx_set_tracking(h_, (int32_t)(double)v, 1));
    } else if (data.date != message_data::undefined()) {
      /* a dated input: one entry of a user constraint (emission series come from the ini) */
      const double d = v;
      chk(hx_set_scenario_series(h_, 0, varName.c_str(), (int32_t)data.date, 1, &d));
    } else {
      chk(hx_set_param_scalar(h_, varName.c_str(), (double)v));
    }
  }
  /* R setvar for a whole ensemble: one value per member */
  void setMembers(const std::string &varName, const std::vector<double> &per_member,
                  unit_types units = U_UNDEFINED) {
    need();
    check_units(varName, units);
    chk(hx_set_param(h_, varName.c_str(), per_member.data(), (int32_t)per_member.size()));
  }

  void prepareToRun() {
    need();
    if (!outputs_selected_) select_all();
    chk(hx_prepare(h_));
    prepared_ = true;
  }
  /* Core::run (core.cpp:448-509).  throw_on_failure mirrors the single-core behaviour: the first
   * failed member's exception is re-thrown; batch callers pass false and read memberStatus(). */
  void run(double runtodate = -1.0, bool throw_on_failure = true) {
    need();
    if (!prepared_) prepareToRun();
    if (runtodate >= 0 && runtodate > getEndDate()) HXB_THROW("Run-to date is after end date.");
    chk(hx_run(h_, runtodate));
    chk(hx_synchronize(h_));
    if (throw_on_failure) {
      std::vector<int32_t> st(n_), fy(n_);
      chk(hx_member_status(h_, st.data(), fy.data(), n_));
      for (int i = 0; i < n_; ++i)
        if (st[i] != HX_MEMBER_OK) {
          std::ostringstream os;
          os << detail::member_failure(st[i]) << " (member " << i << ", year " << fy[i] << ")";
          HXB_THROW(os.str());
        }
    }
  }
  /* Core::reset (core.cpp:511-549): to (or before) the start date -- parameters changed since
   * the last spin-up trigger a new one --, or to a year inside the run already made */
  void reset(double resetdate) {
    need();
    if (!prepared_) return;
    chk(hx_reset_date(h_, resetdate));
  }
  void shutDown() {
    if (h_) hx_destroy(h_);
    h_ = nullptr;
    prepared_ = outputs_selected_ = false;
  }

  double getStartDate() const { return h_ ? start_of(h_) : message_data::undefined(); }
  double getEndDate() const { return end_; }
  double getCurrentDate() const { return h_ ? hx_current_date(h_) : message_data::undefined(); }
  int members() const { return n_; }
  hx_handle handle() const { return h_; }

  /* Core::sendMessage (core.cpp:716-778) for member `member` */
  unitval sendMessage(const std::string &message, const std::string &datum,
                      const message_data &info = message_data(), int member = 0) {
    need();
    if (message == HXB_M_GETDATA) return getData(datum, info.date, member);
    if (message == HXB_M_SETDATA) {
      const unit_types want = detail::units_of(datum);
      const unitval v = info.getUnitval(want, /*strict*/ want != U_UNDEFINED);
      if (n_ == 1) {
        chk(hx_set_param_scalar(h_, datum.c_str(), (double)v));
      } else {
        std::vector<double> cur(n_);
        chk(hx_get_param(h_, datum.c_str(), cur.data(), n_));
        cur.at(member) = (double)v;
        chk(hx_set_param(h_, datum.c_str(), cur.data(), n_));
      }
      return v;
    }
    HXB_THROW("Unknown message: " + message);
  }
  unitval getData(const std::string &varName, double date, int member = 0) {
    need();
    if (member < 0 || member >= n_) HXB_THROW("member index out of range");
    const unit_types u = detail::units_of(varName);
    if (date == message_data::undefined()) { /* a parameter */
      std::vector<double> cur(n_);
      chk(hx_get_param(h_, varName.c_str(), cur.data(), n_));
      return unitval(cur[member], u);
    }
    std::vector<double> col(n_);
    chk(hx_fetch(h_, varName.c_str(), &date, 1, col.data()));
    return unitval(col[member], u);
  }
  /* Core::getTrackingData (core.cpp:199-209): the CSVFluxPoolVisitor's text for one member,
   * "year,component,pool_name,pool_value,pool_units,source_name,source_fraction" for every
   * recorded year up to the current date (csv_tracking_visitor.cpp:80-137); "" when tracking
   * is off.  Numbers are printed with the stream's default 6 significant digits, like there. */
  std::string getTrackingData(int member = 0) {
    need();
    if (member < 0 || member >= n_) HXB_THROW("member index out of range");
    static const char *const pools[HX_TRACK_NPOOL] = {"atmos_co2", "earth_c", "veg_c", "detritus_c",
                                                      "soil_c", "permafrost_c", "thawedp_c", "HL",
                                                      "LL", "intermediate", "deep"};
    static const char *const pool_var[HX_TRACK_NPOOL] = {
        "atmos_co2", "earth_c", "veg_c", "detritus_c", "soil_c", "permafrost_c", "thawedp_c",
        "HL_ocean_c", "LL_ocean_c", "IO_ocean_c", "DO_ocean_c"};
    const int ny = prepared_ ? hx_tracking_years(h_, nullptr, 0) : 0;
    if (ny <= 0) return std::string();
    std::vector<int32_t> years(ny);
    hx_tracking_years(h_, years.data(), ny);
    std::vector<double> frac((size_t)n_ * HX_TRACK_NPOOL * HX_TRACK_NSRC), col(n_);
    std::vector<uint32_t> mask((size_t)n_ * HX_TRACK_NPOOL);
    std::ostringstream os;
    os << "year,component,pool_name,pool_value,pool_units,source_name,source_fraction\n";
    for (int32_t y : years) {
      if (y > getCurrentDate()) break;
      chk(hx_fetch_tracking(h_, (double)y, frac.data(), mask.data()));
      for (int k = 0; k < HX_TRACK_NPOOL; ++k) {
        const double date = y;
        chk(hx_fetch(h_, pool_var[k], &date, 1, col.data()));
        for (int s = 0; s < HX_TRACK_NSRC; ++s)
          if (mask[(size_t)member * HX_TRACK_NPOOL + k] >> s & 1u)
            os << y << "," << (k < 7 ? "simpleNbox" : "ocean") << "," << pools[k] << ","
               << col[member] << ",Pg C," << (s < HX_TRACK_NPOOL ? pools[s] : "untracked") << ","
               << frac[((size_t)member * HX_TRACK_NPOOL + k) * HX_TRACK_NSRC + s] << "\n";
      }
    }
    return os.str();
  }
  /* R fetchvars for the whole ensemble: out[member][date] */
  void fetch(const std::string &varName, const std::vector<double> &dates, double *out) {
    need();
    chk(hx_fetch(h_, varName.c_str(), dates.data(), (int32_t)dates.size(), out));
  }
  void memberStatus(std::vector<int32_t> &status, std::vector<int32_t> &fail_year) {
    need();
    status.resize(n_);
    fail_year.resize(n_);
    chk(hx_member_status(h_, status.data(), fail_year.data(), n_));
  }

 protected:
  void need() const {
    if (!h_) HXB_THROW("no input has been parsed yet (INIToCoreReader::parse)");
  }
  void chk(int rc) const {
    if (rc != HX_OK) HXB_THROW(std::string(hx_last_error(h_)));
  }
  void check_units(const std::string &var, unit_types given) const {
    const unit_types want = detail::units_of(var);
    if (given != U_UNDEFINED && want != U_UNDEFINED && given != want)
      HXB_THROW("Units: " + unitval::unitsName(given) + " do not match expected: " + unitval::unitsName(want));
  }
  void select_all() {
    static const char *const all[] = {
        "CO2_concentration", "global_tas", "RF_tot", "RF_CO2", "heatflux", "ocean_c", "HL_pH",
        "atmos_co2", "sst", "permafrost_c", "CH4_concentration", "N2O_concentration",
        "O3_concentration", "land_tas", "veg_c", "detritus_c", "soil_c", "thawedp_c", "earth_c",
        "NBP", "ocean_uptake", "LL_pH", "HL_PCO2", "LL_PCO2", "HL_ocean_c", "LL_ocean_c",
        "IO_ocean_c", "DO_ocean_c", "RF_CH4", "RF_N2O", "rh_ch4", "NPP", "RH", "gmst",
        "ocean_tas", "heatflux_mixed", "heatflux_interior", "ocean_timesteps"};
    chk(hx_select_outputs(h_, (int32_t)(sizeof all / sizeof all[0]), all));
    outputs_selected_ = true;
  }
  double start_of(hx_handle) const { return start_; }

  friend class INIToCoreReader;
  hx_handle h_ = nullptr;
  int n_, device_;
  unsigned flags_;
  bool prepared_ = false, outputs_selected_ = false;
  std::vector<std::string> biomes_;
  double start_ = message_data::undefined(), end_ = message_data::undefined();
};

/* ---- the single-member face, with the reference's registry (core.cpp:813-857) ---- */
class Core : public EnsembleCore {
 public:
  Core(Logger::LogLevel = Logger::DEBUG, bool /*echotoscreen*/ = true, bool /*echotofile*/ = true)
      : EnsembleCore(1) {}
  static double undefinedIndex() { return message_data::undefined(); }
  static int mkcore(bool /*logtofile*/ = false, Logger::LogLevel lvl = Logger::NOTICE,
                    bool /*logtoscrn*/ = false) {
    registry().push_back(new Core(lvl, false, false));
    return (int)registry().size() - 1;
  }
  static Core *getcore(std::vector<Core *>::size_type idx) {
    return idx < registry().size() ? registry()[idx] : nullptr;
  }
  static void delcore(std::vector<Core *>::size_type idx) { /* slots are never reused */
    if (idx < registry().size() && registry()[idx]) {
      delete registry()[idx];
      registry()[idx] = nullptr;
    }
  }

 private:
  static std::vector<Core *> &registry() {
    static std::vector<Core *> r;
    return r;
  }
};

/* INIToCoreReader(core).parse(file) (ini_to_core_reader.cpp:74-180): the engine's own ini/csv
 * reader fills the scenario tables and scalar parameters */
class INIToCoreReader {
  EnsembleCore *core_;

 public:
  explicit INIToCoreReader(EnsembleCore *core) : core_(core) {}
  void parse(const std::string &filename) { parse(std::vector<std::string>(1, filename)); }
  void parse(const std::vector<std::string> &filenames) {
    int32_t s = 0, e = 0;
    if (filenames.empty()) HXB_THROW("no ini file given");
    if (hx_ini_read(filenames[0].c_str(), &s, &e, nullptr, 0) != HX_OK)
      HXB_THROW(std::string(hx_last_error(nullptr)));
    core_->parse(filenames);
    core_->start_ = s;
    core_->end_ = e;
  }
};

} // namespace hector_b200

#ifdef HECTOR_B200_AS_HECTOR
namespace Hector = hector_b200;
#endif

#endif /* HECTOR_B200_CORE_HPP */
