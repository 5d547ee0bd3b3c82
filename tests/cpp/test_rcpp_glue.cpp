// tests/cpp/test_rcpp_glue.cpp -- drives the reference's R glue (src/rcpp_hector.cpp, compiled
// UNMODIFIED against include/compat and tests/cpp/rcpp_stub/Rcpp.h) the way R/hector.R and
// R/messages.R do: newcore -> run -> fetchvars (sendmessage GETDATA) -> setvar (sendmessage
// SETDATA, undated and dated) -> reset -> run, biome edits, shutdown.  Prints key=value lines that
// tests/test_compat.py compares with the CPU oracle.
#include <Rcpp.h>

#include <cstdio>
#include <string>
#include <vector>

using namespace Rcpp;

// the exported functions of src/rcpp_hector.cpp
Environment newcore_impl(String inifile, int loglevel, bool suppresslogging, String name);
Environment shutdown(Environment core);
Environment reset(Environment core, double date);
Environment run(Environment core, double runtodate);
double getdate(Environment core);
std::string get_tracking_data_impl(Environment core);
std::vector<std::string> get_biome_list(Environment core);
Environment create_biome_impl(Environment core, std::string biome);
Environment delete_biome_impl(Environment core, std::string biome);
Environment rename_biome(Environment core, std::string oldname, std::string newname);
DataFrame sendmessage(Environment core, String msgtype, String capability, NumericVector date,
                      NumericVector value, String unit);
bool chk_core_valid(Environment core);

static const double NA = NumericVector::get_na();

static double fetch1(Environment core, const char *var, double year) {
  DataFrame d = sendmessage(core, "getData", var, NumericVector{year}, NumericVector{NA}, "");
  return d.col("value").num.at(0);
}
static void setvar(Environment core, const char *var, double value, const char *unit) {
  sendmessage(core, "setData", var, NumericVector{NA}, NumericVector{value}, unit);
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  try {
    Environment core = newcore_impl(argv[1], 3, true, "glue");
    std::printf("STRT=%.0f\nEND=%.0f\nTRACK=%.0f\n", (double)core["strtdate"], (double)core["enddate"],
                (double)core["trackdate"]);
    run(core, 2100);
    std::printf("DATE=%.0f\n", getdate(core));
    DataFrame d = sendmessage(core, "getData", "global_tas", NumericVector{2000, 2100}, NumericVector{NA}, "");
    std::printf("TAS_2000=%.17g\nTAS_2100=%.17g\nTAS_UNITS=%s\n", d.col("value").num[0], d.col("value").num[1],
                d.col("units").str[0].c_str());
    std::printf("CO2_2100=%.17g\n", fetch1(core, "CO2_concentration", 2100));
    // setvar(core, NA, ECS(), 4.5, "degC"); reset(core); run(core)
    setvar(core, "S", 4.5, "degC");
    core["clean"] = false;  // what R's setvar does; run() then auto-resets
    run(core, -1.0);
    std::printf("S45_TAS_2300=%.17g\nS45_CO2_2300=%.17g\n", fetch1(core, "global_tas", 2300),
                fetch1(core, "CO2_concentration", 2300));
    // setvar(core, 2030:2040, FFI_EMISSIONS(), 0, "Pg C/yr"): dated inputs, one message per year
    NumericVector yrs(11), zero(11);
    for (int i = 0; i < 11; ++i) { yrs[i] = 2030 + i; zero[i] = 0.0; }
    sendmessage(core, "setData", "ffi_emissions", yrs, zero, "Pg C/yr");
    reset(core, 0);
    run(core, 2100);
    std::printf("FFI0_CO2_2100=%.17g\n", fetch1(core, "CO2_concentration", 2100));
    // a wrong unit is refused, the core stays usable (tests/testthat/test_wrapper.R:135-145)
    bool refused = false;
    try { setvar(core, "S", 3.0, "Pg C"); } catch (Rcpp::exception &) { refused = true; }
    std::printf("BAD_UNIT_REFUSED=%d\nSTILL_VALID=%d\n", refused ? 1 : 0, chk_core_valid(core) ? 1 : 0);
    // biomes: rename the global biome, add a second one, give it half of everything
    Environment bc = newcore_impl(argv[1], 3, true, "biomes");
    std::printf("BIOMES0=%s\n", get_biome_list(bc)[0].c_str());
    rename_biome(bc, "global", "boreal");
    create_biome_impl(bc, "tropical");
    std::vector<std::string> bl = get_biome_list(bc);
    std::printf("BIOMES=%s,%s\n", bl[0].c_str(), bl.size() > 1 ? bl[1].c_str() : "");
    const char *pools[] = {"veg_c", "detritus_c", "soil_c", "permafrost_c", "npp_flux0"};
    const char *punit[] = {"Pg C", "Pg C", "Pg C", "Pg C", "Pg C/yr"};
    for (int k = 0; k < 5; ++k) {
      DataFrame cur = sendmessage(bc, "getData", std::string("boreal.") + pools[k], NumericVector{NA}, NumericVector{NA}, "");
      const double v = cur.col("value").num[0];
      setvar(bc, (std::string("boreal.") + pools[k]).c_str(), 0.4 * v, punit[k]);
      setvar(bc, (std::string("tropical.") + pools[k]).c_str(), 0.6 * v, punit[k]);
    }
    const char *pars[] = {"beta", "q10_rh", "f_nppv", "f_nppd", "f_litterd"};
    const double parv[] = {0.5, 2.2, 0.35, 0.60, 0.98};
    for (int k = 0; k < 5; ++k) setvar(bc, (std::string("tropical.") + pars[k]).c_str(), parv[k], "(unitless)");
    setvar(bc, "boreal.warmingfactor", 1.8, "(unitless)");
    reset(bc, 0);
    run(bc, -1.0);
    std::printf("BIO_CO2_2300=%.17g\nBIO_BOREAL_VEG_2100=%.17g\nBIO_VEG_2100=%.17g\n",
                fetch1(bc, "CO2_concentration", 2300), fetch1(bc, "boreal.veg_c", 2100), fetch1(bc, "veg_c", 2100));
    delete_biome_impl(bc, "tropical");
    std::printf("BIOMES_AFTER_DELETE=%zu\n", get_biome_list(bc).size());
    shutdown(bc);
    shutdown(core);
    std::printf("VALID_AFTER_SHUTDOWN=%d\n", chk_core_valid(core) ? 1 : 0);
    std::printf("DONE=1\n");
  } catch (std::exception &e) {
    std::printf("EXCEPTION=%s\n", e.what());
    return 1;
  }
  return 0;
}
