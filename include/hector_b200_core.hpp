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
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "hector_b200.h"

/* ---- h_exception (h_exception.hpp:26-102): message + where it was raised.  Like the reference's
 * it lives in the GLOBAL namespace: the R glue catches it unqualified (rcpp_hector.cpp:56). ---- */
#ifndef HECTOR_B200_H_EXCEPTION
#define HECTOR_B200_H_EXCEPTION
class h_exception : public std::exception {
  std::string msg_, func_, file_;
  int line_;

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
#endif

namespace hector_b200 {

using ::h_exception;

#define HXB_M_GETDATA "getData"
#define HXB_M_SETDATA "setData"
#ifndef M_GETDATA
#define M_GETDATA HXB_M_GETDATA
#define M_SETDATA HXB_M_SETDATA
#endif

#define HXB_THROW(m) throw ::h_exception((m), __func__, __FILE__, __LINE__)

/* ---- units (unitval.hpp:68-130): the ones variables of this path carry ---- */
enum unit_types {
  U_UNITLESS, U_PPMV_CO2, U_PPBV_CH4, U_PPBV_N2O, U_DU_O3, U_TG_PPBV, U_DEGC, U_CM2_S, U_PGC,
  U_PGC_YR, U_W_M2, U_W_M2_TG, U_W_M2_GG, U_M3_S, U_PH, U_UATM, U_YRS, U_PPTV, U_UMOL_KG,
  U_UNDEFINED
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
                                        "pptv", "umol/kg", "(undefined)"};
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
    if (expected != U_UNDEFINED) r.expecting_unit(expected); /* no opinion: keep the caller's */
    return r;
  }
  double date;
  std::string value_str;
  unitval value_unitval;
  std::string units_str;
  bool isVal;
};

/* Logger (logger.hpp:35-118): levels, H_LOG's shouldWrite / write pair, close.  The engine itself
 * does not log; this serves callers that write to the core's global log (src/main.cpp:43-113). */
class Logger {
 public:
  enum LogLevel { DEBUG, NOTICE, WARNING, SEVERE };
  Logger() : min_(WARNING), screen_(false) {}
  void open(const std::string & /*logName*/, bool echoToScreen, bool /*echoToFile*/, LogLevel minLogLevel) {
    min_ = minLogLevel;
    screen_ = echoToScreen;
  }
  bool shouldWrite(const LogLevel writeLevel) const { return screen_ && writeLevel >= min_; }
  std::ostream &write(const LogLevel writeLevel, const std::string &functionInfo) {
    static const char *const names[] = {"DEBUG", "NOTICE", "WARNING", "SEVERE"};
    return std::clog << names[(int)writeLevel] << ":" << functionInfo << ": ";
  }
  void close() { std::clog.flush(); }

 private:
  LogLevel min_;
  bool screen_;
};

class Core;
/* AVisitor (avisitor.hpp:44-88): what Core::addVisitor takes.  Only visit(Core*) is ever called
 * here -- the components behind the engine are not objects a visitor could be handed. */
class AVisitor {
 public:
  virtual ~AVisitor() {}
  virtual bool shouldVisit(const bool in_spinup, const double date) = 0;
  virtual void reset(const double /*reset_date*/) {}
  virtual void outputTrackingData(std::ostream & /*tracking_out*/) const {}
  virtual void visit(Core * /*core*/) {}
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
      /* functions of recorded outputs, evaluated by hx_fetch (include/hector_b200.h) */
      {"HL_sst", U_DEGC}, {"LL_sst", U_DEGC}, {"HL_DIC", U_UMOL_KG}, {"LL_DIC", U_UMOL_KG},
      {"DIC", U_UMOL_KG}, {"pH", U_PH}, {"PCO2", U_UATM}, {"ML_ocean_c", U_PGC}, {"TAU_OH", U_YRS},
      {"HL_ocean_uptake", U_PGC_YR}, {"LL_ocean_uptake", U_PGC_YR}, {"rh_det", U_PGC_YR},
      {"rh_soil", U_PGC_YR},
      {"HL_OmegaCa", U_UNITLESS}, {"LL_OmegaCa", U_UNITLESS}, {"HL_OmegaAr", U_UNITLESS},
      {"LL_OmegaAr", U_UNITLESS}, {"baseyear", U_UNITLESS},
      {"f_frozen", U_UNITLESS}, {"HL_CO3", U_UMOL_KG}, {"LL_CO3", U_UMOL_KG}, {"CO3", U_UMOL_KG},
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
      {"trackingDate", U_UNITLESS},
      /* emission series (component_data.hpp; unit checks of each component's setData) */
      {"ffi_emissions", U_PGC_YR}, {"daccs_uptake", U_PGC_YR}, {"luc_emissions", U_PGC_YR},
      {"luc_uptake", U_PGC_YR}};
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
    case HX_MEMBER_NEEDS_EXACT:
      return "an ODE attempt the engine skips could have raised 'Flux and pool values may not be "
             "negative'; re-run with HX_FLAG_EXACT_ATTEMPTS";
    default: return "model failure";
  }
}
} // namespace detail

/* ---- M members behind Hector's message surface ---- */
class EnsembleCore {
 public:
  /* all recorded variables unless selectOutputs names a subset (fewer outputs = less HBM traffic) */
  explicit EnsembleCore(int n_members = 1, int device = 0, unsigned flags = 0)
      : n_(n_members), device_(device), flags_(flags) {}
  EnsembleCore(const EnsembleCore &) = delete;
  EnsembleCore &operator=(const EnsembleCore &) = delete;
  virtual ~EnsembleCore() { shutDown(); }

  void init() {} /* components are created with the engine (core.cpp:90-196) */
  Logger &getGlobalLogger() { return glog_; }

  /* INIToCoreReader::parse: one ini file per scenario */
  void parse(const std::vector<std::string> &ini_files) {
    shutDown();
    ini_files_ = ini_files;
    journal_.clear();
    biomes_.clear();
    biomes_edited_ = false;
    open_engine();
    /* biomes an ini file declares ("<biome>.<name>" lines, simpleNbox.cpp:229-236): remember the
     * list and their values, so that a later createBiome / renameBiome can rebuild the engine */
    const int nb = hx_biome_count(h_);
    if (nb > 1) {
      for (int i = 0; i < nb; ++i) {
        char buf[256];
        chk(hx_biome_name(h_, i, buf, (int32_t)sizeof buf));
        biomes_.push_back(buf);
      }
      for (const std::string &b : biomes_) capture_biome(b + ".", b + ".");
    }
  }
  void selectOutputs(const std::vector<std::string> &names) {
    need();
    apply_outputs(names);
    outputs_ = names;
  }
  void setMemberScenario(const std::vector<int32_t> &scenario_of_member) {
    need();
    chk(hx_set_member_scenario(h_, scenario_of_member.data(), (int32_t)scenario_of_member.size()));
    member_scenario_ = scenario_of_member;
  }

  /* ---- biomes (core.cpp:560-599 -> simpleNbox.cpp:864-1100) ----
   * The engine fixes its biome list when it is prepared, so an edit rebuilds it: the ini files
   * are read again, the new list is declared and every setData / sendMessage(SETDATA) made so
   * far is replayed; it is prepared again (set-up + spin-up) when next needed.  A core whose only
   * biome has been renamed keeps running the global code path under the new name. */
  std::vector<std::string> getBiomeList() const {
    return biomes_.empty() ? std::vector<std::string>(1, "global") : biomes_;
  }
  void setBiomes(const std::vector<std::string> &names) { /* declare the whole list at once */
    need();
    for (const std::string &b : names)
      if (b.empty() || b.find('.') != std::string::npos) HXB_THROW("bad biome name '" + b + "'");
    rebuild(names);
  }
  void createBiome(const std::string &biome) {
    need();
    if (has_biome(biome)) HXB_THROW("Assertion failed: Biome '" + biome + "' is already in `biome_list`.");
    if (biomes_.empty())
      HXB_THROW("Assertion failed: If one of the biomes is 'global', you cannot add other biomes.");
    std::vector<std::string> nl = biomes_;
    nl.push_back(biome);
    /* new pools and initial NPP are zero (simpleNbox.cpp:873-900); parameters stay unset */
    static const char *const zeroed[] = {"veg_c", "detritus_c", "soil_c", "permafrost_c", "npp_flux0"};
    for (const char *p : zeroed) journal_.push_back(SetCall::scalar(biome + "." + p, 0.0));
    rebuild(nl);
  }
  void deleteBiome(const std::string &biome) {
    need();
    if (!has_biome(biome)) HXB_THROW("Assertion failed: Biome '" + biome + "' not found in `biome_list`.");
    if (getBiomeList().size() == 1) HXB_THROW("cannot delete the only biome");
    std::vector<std::string> nl;
    for (const std::string &b : biomes_)
      if (b != biome) nl.push_back(b);
    drop_journal_prefix(biome + ".");
    rebuild(nl);
  }
  void renameBiome(const std::string &oldname, const std::string &newname) {
    need();
    if (!has_biome(oldname)) HXB_THROW("Assertion failed: Biome '" + oldname + "' not found in `biome_list`.");
    if (has_biome(newname)) HXB_THROW("Assertion failed: Biome '" + newname + "' already exists in `biome_list`.");
    if (newname.empty() || newname.find('.') != std::string::npos) HXB_THROW("bad biome name '" + newname + "'");
    std::vector<std::string> nl = getBiomeList();
    for (std::string &b : nl)
      if (b == oldname) b = newname;
    /* the old biome's pools and parameters, as they stand, become the new one's */
    const std::string from = engine_prefix(oldname);
    drop_journal_prefix(oldname + ".");
    if (from.empty()) drop_global_land_params();
    capture_biome(from, newname + ".");
    if (nl.size() == 1 && nl[0] == "global") nl.clear();
    rebuild(nl);
  }

  /* Core::setData (core.cpp:219-268): the component name only routes in the reference */
  void setData(const std::string & /*componentName*/, const std::string &varName,
               const message_data &data) {
    need();
    const unit_types want = detail::units_of(varName);
    const unitval v = data.getUnitval(want);
    SetCall c = varName == "trackingDate"                    ? SetCall::tracking((double)v)
                : data.date != message_data::undefined()     ? SetCall::dated(varName, data.date, (double)v)
                                                             : SetCall::scalar(varName, (double)v);
    apply(c);
    journal_.push_back(c);
  }
  /* R setvar for a whole ensemble: one value per member */
  void setMembers(const std::string &varName, const std::vector<double> &per_member,
                  unit_types units = U_UNDEFINED) {
    need();
    check_units(varName, units);
    SetCall c = SetCall::members(varName, per_member);
    apply(c);
    journal_.push_back(c);
  }

  void prepareToRun() {
    need();
    ensure_prepared();
  }
  /* Core::run (core.cpp:448-509).  throw_on_failure mirrors the single-core behaviour: the first
   * failed member's exception is re-thrown; batch callers pass false and read memberStatus(). */
  void run(double runtodate = -1.0, bool throw_on_failure = true) {
    need();
    ensure_prepared();
    if (runtodate >= 0 && runtodate > getEndDate()) HXB_THROW("Run-to date is after end date.");
    const double from = hx_current_date(h_);
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
    after_run(from, hx_current_date(h_));
  }
  /* Core::reset (core.cpp:511-549): to (or before) the start date -- parameters changed since
   * the last spin-up trigger a new one --, or to a year inside the run already made */
  void reset(double resetdate) {
    need();
    if (prepared_) chk(hx_reset_date(h_, resetdate)); /* an engine not yet prepared is at the start */
    after_reset(resetdate);
  }
  void shutDown() {
    if (h_) hx_destroy(h_);
    h_ = nullptr;
    prepared_ = false;
  }

  double getStartDate() const { return h_ ? start_ : message_data::undefined(); }
  double getEndDate() const { return end_; }
  double getCurrentDate() const { return h_ ? hx_current_date(h_) : message_data::undefined(); }
  /* Core::getTrackingDate (core.hpp:85): 9999 when tracking is off (core.cpp:60) */
  double getTrackingDate() const { return h_ ? (double)hx_tracking_date(h_) : 9999.0; }
  /* [core] run_name of the (first) ini file (core.hpp:88) */
  std::string getRun_name() const {
    char buf[512] = "";
    if (!ini_files_.empty()) hx_ini_string(ini_files_[0].c_str(), "run_name", buf, (int32_t)sizeof buf);
    return buf;
  }
  bool outputEnabled(const std::string & /*componentName*/) const { return true; }
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
      SetCall c;
      if (info.date != message_data::undefined()) {
        /* a dated input: one year of an emission series or of a user constraint -- R's
         * setvar(core, dates, var, values, unit) arrives here once per date */
        c = SetCall::dated(datum, info.date, (double)v);
      } else if (n_ == 1) {
        c = SetCall::scalar(datum, (double)v);
      } else {
        /* one member of an ensemble: one element on the host and on the device (a journal
         * entry of a few bytes, not the ensemble's whole vector) */
        if (member < 0 || member >= n_) HXB_THROW("member index out of range");
        c = SetCall::one(datum, member, (double)v);
      }
      apply(c);
      journal_.push_back(c);
      return v;
    }
    HXB_THROW("Unknown message: " + message);
  }
  unitval getData(const std::string &varName, double date, int member = 0) {
    need();
    if (member < 0 || member >= n_) HXB_THROW("member index out of range");
    const unit_types u = detail::units_of(varName);
    const std::string name = engine_name(varName);
    std::vector<double> col(n_);
    if (date == message_data::undefined()) {
      /* a parameter -- or, for a variable, its value at the current date (the reference's
       * undated getData; misc/main-api.cpp:136-138) */
      if (hx_get_param(h_, name.c_str(), col.data(), n_) == HX_OK) return unitval(col[member], u);
      date = getCurrentDate();
    }
    ensure_prepared();
    chk(hx_fetch(h_, name.c_str(), &date, 1, col.data()));
    return unitval(col[member], u);
  }
  /* Core::getTrackingData (core.cpp:199-209): the CSVFluxPoolVisitor's text for one member,
   * "year,component,pool_name,pool_value,pool_units,source_name,source_fraction" for every
   * recorded year up to the current date (csv_tracking_visitor.cpp:80-137); "" when tracking
   * is off.  Numbers are printed with the stream's default 6 significant digits, like there. */
  std::string getTrackingData(int member = 0) {
    need();
    if (member < 0 || member >= n_) HXB_THROW("member index out of range");
    const int ny = prepared_ ? hx_tracking_years(h_, nullptr, 0) : 0;
    if (ny <= 0) return std::string();
    std::vector<int32_t> years(ny);
    hx_tracking_years(h_, years.data(), ny);
    std::ostringstream os;
    os << "year,component,pool_name,pool_value,pool_units,source_name,source_fraction\n";
    for (int32_t y : years) {
      if (y > getCurrentDate()) break;
      trackingRows(os, y, member);
    }
    return os.str();
  }
  /* the tracking rows of one recorded year (or of the current date) */
  void trackingRows(std::ostream &os, int year, int member = 0) {
    static const char *const pools[HX_TRACK_NPOOL] = {"atmos_co2", "earth_c", "veg_c", "detritus_c",
                                                      "soil_c", "permafrost_c", "thawedp_c", "HL",
                                                      "LL", "intermediate", "deep"};
    static const char *const pool_var[HX_TRACK_NPOOL] = {
        "atmos_co2", "earth_c", "veg_c", "detritus_c", "soil_c", "permafrost_c", "thawedp_c",
        "HL_ocean_c", "LL_ocean_c", "IO_ocean_c", "DO_ocean_c"};
    std::vector<double> frac((size_t)n_ * HX_TRACK_NPOOL * HX_TRACK_NSRC), col(n_);
    std::vector<uint32_t> mask((size_t)n_ * HX_TRACK_NPOOL);
    chk(hx_fetch_tracking(h_, (double)year, frac.data(), mask.data()));
    for (int k = 0; k < HX_TRACK_NPOOL; ++k) {
      const double date = year;
      chk(hx_fetch(h_, pool_var[k], &date, 1, col.data()));
      for (int s = 0; s < HX_TRACK_NSRC; ++s)
        if (mask[(size_t)member * HX_TRACK_NPOOL + k] >> s & 1u)
          os << year << "," << (k < 7 ? "simpleNbox" : "ocean") << "," << pools[k] << ","
             << col[member] << ",Pg C," << (s < HX_TRACK_NPOOL ? pools[s] : "untracked") << ","
             << frac[((size_t)member * HX_TRACK_NPOOL + k) * HX_TRACK_NSRC + s] << "\n";
    }
  }
  /* R fetchvars for the whole ensemble: out[member][date] */
  void fetch(const std::string &varName, const std::vector<double> &dates, double *out) {
    need();
    ensure_prepared();
    chk(hx_fetch(h_, engine_name(varName).c_str(), dates.data(), (int32_t)dates.size(), out));
  }
  void memberStatus(std::vector<int32_t> &status, std::vector<int32_t> &fail_year) {
    need();
    ensure_prepared();
    status.resize(n_);
    fail_year.resize(n_);
    chk(hx_member_status(h_, status.data(), fail_year.data(), n_));
  }
  /* is `var` among the recorded outputs (the visitors print what is there) */
  bool isRecorded(const std::string &var) const {
    if (outputs_.empty()) return true; /* everything */
    for (const std::string &o : outputs_)
      if (o == var) return true;
    return false;
  }

 protected:
  /* one setData / sendMessage(SETDATA) / setMembers, kept so that an engine rebuilt for a new
   * biome list can be brought back to the same inputs */
  struct SetCall {
    enum Kind { SCALAR, DATED, MEMBERS, TRACKING, MEMBER } kind = SCALAR;
    std::string name;
    double date = 0.0, value = 0.0;
    int member = 0;
    std::vector<double> per_member;
    static SetCall scalar(const std::string &n, double v) { SetCall c; c.kind = SCALAR; c.name = n; c.value = v; return c; }
    static SetCall dated(const std::string &n, double d, double v) { SetCall c; c.kind = DATED; c.name = n; c.date = d; c.value = v; return c; }
    static SetCall members(const std::string &n, const std::vector<double> &v) { SetCall c; c.kind = MEMBERS; c.name = n; c.per_member = v; return c; }
    static SetCall tracking(double v) { SetCall c; c.kind = TRACKING; c.value = v; return c; }
    static SetCall one(const std::string &n, int m, double v) { SetCall c; c.kind = MEMBER; c.name = n; c.member = m; c.value = v; return c; }
  };
  void apply(const SetCall &c) {
    const std::string name = engine_name(c.name);
    switch (c.kind) {
      case SetCall::TRACKING: chk(hx_set_tracking(h_, (int32_t)c.value, 1)); break; /* core.cpp:228-235 */
      case SetCall::DATED: chk(hx_set_scenario_series(h_, 0, name.c_str(), (int32_t)c.date, 1, &c.value)); break;
      case SetCall::MEMBERS: chk(hx_set_param(h_, name.c_str(), c.per_member.data(), (int32_t)c.per_member.size())); break;
      case SetCall::MEMBER: chk(hx_set_param_member(h_, name.c_str(), (int32_t)c.member, c.value)); break;
      default: chk(hx_set_param_scalar(h_, name.c_str(), c.value));
    }
  }
  void open_engine() {
    std::vector<const char *> p;
    for (const std::string &f : ini_files_) p.push_back(f.c_str());
    if (hx_create_from_ini(p.data(), (int32_t)p.size(), n_, device_, flags_, &h_) != HX_OK)
      HXB_THROW(std::string(hx_last_error(nullptr)));
    prepared_ = false;
    if (!member_scenario_.empty())
      chk(hx_set_member_scenario(h_, member_scenario_.data(), (int32_t)member_scenario_.size()));
    if (biomes_edited_) {
      if (biomes_.size() > 1) {
        std::vector<const char *> b;
        for (const std::string &s : biomes_) b.push_back(s.c_str());
        chk(hx_set_biomes(h_, (int32_t)b.size(), b.data()));
      } else {
        chk(hx_set_biomes(h_, 1, nullptr)); /* one biome, whatever its name: the global code path */
      }
    }
    for (const SetCall &c : journal_) apply(c);
    if (!outputs_.empty()) apply_outputs(outputs_);
  }
  void rebuild(const std::vector<std::string> &new_list) {
    if (h_) hx_destroy(h_);
    h_ = nullptr;
    biomes_ = new_list;
    biomes_edited_ = true;
    /* per-biome outputs of biomes that no longer exist cannot be selected any more */
    std::vector<std::string> keep;
    for (const std::string &o : outputs_) {
      const size_t dot = o.find('.');
      if (dot == std::string::npos || has_biome(o.substr(0, dot))) keep.push_back(o);
    }
    outputs_ = keep;
    open_engine();
  }
  void ensure_prepared() {
    if (prepared_) return;
    if (outputs_.empty()) select_all();
    chk(hx_prepare(h_));
    prepared_ = true;
  }
  bool has_biome(const std::string &b) const {
    for (const std::string &x : getBiomeList())
      if (x == b) return true;
    return false;
  }
  /* how the engine spells the inputs of biome `b`: "<b>." with several biomes, "" for the only one */
  std::string engine_prefix(const std::string &b) const { return biomes_.size() > 1 ? b + "." : std::string(); }
  /* "<biome>.<name>" of a core with ONE (renamed) biome is the engine's plain <name> */
  std::string engine_name(const std::string &name) const {
    if (biomes_.size() == 1 && name.compare(0, biomes_[0].size() + 1, biomes_[0] + ".") == 0)
      return name.substr(biomes_[0].size() + 1);
    return name;
  }
  /* journal the 15 per-biome inputs as the engine holds them under `from_prefix`, renamed */
  void capture_biome(const std::string &from_prefix, const std::string &to_prefix) {
    static const char *const names[] = {"veg_c", "detritus_c", "soil_c", "permafrost_c", "npp_flux0",
                                        "beta", "q10_rh", "warmingfactor", "f_nppv", "f_nppd",
                                        "f_litterd", "rh_ch4_frac", "pf_mu", "pf_sigma", "fpf_static"};
    std::vector<double> cur(n_);
    for (const char *p : names) {
      chk(hx_get_param(h_, (from_prefix + p).c_str(), cur.data(), n_));
      bool same = true, unset = false;
      for (int i = 0; i < n_; ++i) { same = same && cur[i] == cur[0]; unset = unset || cur[i] != cur[i]; }
      if (unset) continue; /* a created biome's parameter that was never given */
      journal_.push_back(same ? SetCall::scalar(to_prefix + p, cur[0]) : SetCall::members(to_prefix + p, cur));
    }
  }
  void drop_journal_prefix(const std::string &prefix) {
    std::vector<SetCall> keep;
    for (const SetCall &c : journal_)
      if (c.name.compare(0, prefix.size(), prefix) != 0) keep.push_back(c);
    journal_.swap(keep);
  }
  void drop_global_land_params() { /* they live on under the biome's name */
    static const char *const names[] = {"veg_c", "detritus_c", "soil_c", "permafrost_c", "npp_flux0",
                                        "beta", "q10_rh", "warmingfactor", "f_nppv", "f_nppd",
                                        "f_litterd", "rh_ch4_frac", "pf_mu", "pf_sigma", "fpf_static"};
    std::vector<SetCall> keep;
    for (const SetCall &c : journal_) {
      bool land = false;
      for (const char *p : names) land = land || c.name == p;
      if (!land) keep.push_back(c);
    }
    journal_.swap(keep);
  }
  virtual void after_run(double /*from*/, double /*to*/) {}
  virtual void after_reset(double /*date*/) {}
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
  void apply_outputs(const std::vector<std::string> &names) {
    std::vector<std::string> en;
    for (const std::string &s : names) en.push_back(engine_name(s));
    std::vector<const char *> p;
    for (const std::string &s : en) p.push_back(s.c_str());
    chk(hx_select_outputs(h_, (int32_t)p.size(), p.data()));
  }
  void select_all() {
    static const char *const all[] = {
        "CO2_concentration", "global_tas", "RF_tot", "RF_CO2", "heatflux", "ocean_c", "HL_pH",
        "atmos_co2", "sst", "permafrost_c", "CH4_concentration", "N2O_concentration",
        "O3_concentration", "land_tas", "veg_c", "detritus_c", "soil_c", "thawedp_c", "earth_c",
        "NBP", "ocean_uptake", "LL_pH", "HL_PCO2", "LL_PCO2", "HL_ocean_c", "LL_ocean_c",
        "IO_ocean_c", "DO_ocean_c", "RF_CH4", "RF_N2O", "rh_ch4", "NPP", "RH", "gmst",
        "ocean_tas", "heatflux_mixed", "heatflux_interior", "ocean_timesteps"};
    std::vector<std::string> names(all, all + sizeof all / sizeof all[0]);
    for (const char *v : {"HL_ocean_uptake", "LL_ocean_uptake", "rh_det", "rh_soil"}) names.push_back(v);
    if (biomes_.size() > 1) { /* "<biome>.<name>": every biome's own pools and fluxes */
      static const char *const own[] = {"veg_c", "detritus_c", "soil_c", "permafrost_c", "thawedp_c", "NPP", "RH"};
      for (const std::string &b : biomes_)
        for (const char *v : own) names.push_back(b + "." + v);
    }
    std::vector<const char *> p;
    for (const std::string &n : names) p.push_back(n.c_str());
    chk(hx_select_outputs(h_, (int32_t)p.size(), p.data()));
  }

  friend class INIToCoreReader;
  hx_handle h_ = nullptr;
  int n_, device_;
  unsigned flags_;
  bool prepared_ = false, biomes_edited_ = false;
  std::vector<std::string> ini_files_, biomes_, outputs_;
  std::vector<int32_t> member_scenario_;
  std::vector<SetCall> journal_;
  Logger glog_;
  double start_ = message_data::undefined(), end_ = message_data::undefined();
};

/* ---- the single-member face, with the reference's registry (core.cpp:813-857) ---- */
class Core : public EnsembleCore {
 public:
  Core(Logger::LogLevel loglvl = Logger::DEBUG, bool echotoscreen = true, bool echotofile = true)
      : EnsembleCore(1) {
    glog_.open("hector", echotoscreen, echotofile, loglvl);
  }
  static double undefinedIndex() { return message_data::undefined(); }
  /* Core::addVisitor (core.cpp:429-436): visitors stay the caller's.  After every run() they are
   * shown each year just computed, oldest first: shouldVisit(false, year), then visit(this) --
   * what the reference does at the end of every time step (core.cpp:497-503); the spin-up is
   * not shown. */
  void addVisitor(AVisitor *visitor) { visitors_.push_back(visitor); }
  /* the year a visitor is being shown (getCurrentDate() is already the end of the run) */
  double visitDate() const { return visit_date_; }
  static int mkcore(bool /*logtofile*/ = false, Logger::LogLevel lvl = Logger::NOTICE,
                    bool logtoscrn = false) {
    registry().push_back(new Core(lvl, logtoscrn, false));
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

 protected:
  void after_run(double from, double to) override {
    if (visitors_.empty()) return;
    for (double y = from + 1; y <= to; y += 1.0) {
      visit_date_ = y;
      for (AVisitor *v : visitors_)
        if (v->shouldVisit(false, y)) v->visit(this);
    }
  }
  void after_reset(double date) override {
    for (AVisitor *v : visitors_) v->reset(date);
  }

 private:
  std::vector<AVisitor *> visitors_;
  double visit_date_ = 0.0;
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
