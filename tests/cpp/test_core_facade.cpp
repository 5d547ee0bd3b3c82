// tests/cpp/test_core_facade.cpp -- drives the C++ facade (include/hector_b200_core.hpp) the way
// the reference's own callers drive Hector::Core: src/main.cpp:41-112 (init / parse /
// prepareToRun / run), src/rcpp_hector.cpp:31-365 (mkcore / getcore / sendMessage / reset) and
// tests/testthat/test_wrapper.R:135-145 (a core stays usable after an exception).
//
//   test_core_facade <ini> [nogpu]
// prints KEY=VALUE lines (doubles as %.17g) that tests/test_core_facade.py checks against the
// oracle; exits non-zero on any behavioural mismatch.
#define HECTOR_B200_AS_HECTOR
#include "hector_b200_core.hpp"

#include <cstdio>
#include <cstdlib>

using namespace Hector;

static int failures = 0;
#define EXPECT(cond)                                                     \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond);       \
      ++failures;                                                        \
    }                                                                    \
  } while (0)

template <class F>
static bool throws(F f) {
  try {
    f();
  } catch (const h_exception &) {
    return true;
  }
  return false;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string ini = argv[1];
  const bool nogpu = argc > 2 && std::string(argv[2]) == "nogpu";

  if (nogpu) { /* the product has no CPU path: parse must throw, naming CUDA */
    Core core(Logger::SEVERE, false, false);
    core.init();
    bool thrown = false;
    try {
      INIToCoreReader(&core).parse(ini);
    } catch (const h_exception &e) {
      thrown = true;
      std::printf("NOGPU_MSG=%s\n", e.what());
    }
    EXPECT(thrown);
    EXPECT(throws([&] { core.run(); })); /* nothing parsed */
    return failures ? 1 : 0;
  }

  /* ---- single core through the registry, like the R glue ---- */
  const int idx = Core::mkcore(false, Logger::SEVERE, false);
  Core *core = Core::getcore(idx);
  EXPECT(core != nullptr);
  EXPECT(Core::getcore(idx + 7) == nullptr);
  core->init();
  INIToCoreReader(core).parse(ini);
  EXPECT(core->getStartDate() == 1745 && core->getEndDate() == 2300);
  core->prepareToRun();
  core->run(2100);
  EXPECT(core->getCurrentDate() == 2100);
  std::printf("D_CO2_2100=%.17g\n", (double)core->sendMessage(M_GETDATA, "CO2_concentration", message_data(2100.0)));
  core->run(); /* resume to the end date */
  const unitval tas = core->sendMessage(M_GETDATA, "global_tas", message_data(2300.0));
  EXPECT(tas.units() == U_DEGC);
  EXPECT(tas.unitsName() == "degC");
  std::printf("D_TAS_2300=%.17g\n", tas.value(U_DEGC));
  std::printf("D_CO2_2300=%.17g\n", (double)core->sendMessage(M_GETDATA, "CO2_concentration", message_data(2300.0)));
  std::printf("D_PH_2000=%.17g\n", (double)core->sendMessage(M_GETDATA, "HL_pH", message_data(2000.0)));
  /* parameters read back with their units */
  const unitval S0 = core->sendMessage(M_GETDATA, "S");
  EXPECT((double)S0 == 3.0 && S0.units() == U_DEGC);

  /* ---- errors: same places as the reference, and the core stays usable ---- */
  EXPECT(throws([&] { core->sendMessage(M_GETDATA, "no_such_variable", message_data(2000.0)); }));
  EXPECT(throws([&] { core->sendMessage(M_GETDATA, "global_tas", message_data(2301.0)); }));
  EXPECT(throws([&] { core->sendMessage(M_GETDATA, "global_tas", message_data(1744.0)); }));
  /* the start date is a valid date, as in the reference (tests/testthat/test_parameters.R:46):
   * the preindustrial concentration, a zero temperature */
  EXPECT((double)core->sendMessage(M_GETDATA, "global_tas", message_data(1745.0)) == 0.0);
  EXPECT((double)core->sendMessage(M_GETDATA, "CO2_concentration", message_data(1745.0)) == 277.15);
  EXPECT(throws([&] { core->sendMessage("noSuchMessage", "global_tas"); }));
  EXPECT(throws([&] { core->sendMessage(M_SETDATA, "S", message_data(unitval(3.0, U_PGC))); }));
  EXPECT(throws([&] { tas.value(U_PGC); }));
  EXPECT(throws([&] { core->run(2400); }));
  EXPECT(throws([&] { unitval::parseUnitsName("furlongs"); }));

  /* ---- setvar + reset + run (R: setvar(core, NA, ECS(), 4.5, "degC"); reset(core); run(core)) ---- */
  core->sendMessage(M_SETDATA, "S", message_data(unitval(4.5, unitval::parseUnitsName("degC"))));
  core->sendMessage(M_SETDATA, "beta", message_data(unitval(0.4, U_UNITLESS)));
  core->reset(0);
  EXPECT(core->getCurrentDate() == 1745);
  core->run();
  std::printf("S45_CO2_2300=%.17g\n", (double)core->sendMessage(M_GETDATA, "CO2_concentration", message_data(2300.0)));
  std::printf("S45_TAS_2300=%.17g\n", (double)core->sendMessage(M_GETDATA, "global_tas", message_data(2300.0)));

  /* a member the reference aborts: run() must throw, and the core must survive it */
  core->sendMessage(M_SETDATA, "beta", message_data(unitval(50.0, U_UNITLESS)));
  core->reset(0);
  bool failed = false;
  try {
    core->run();
  } catch (const h_exception &e) {
    failed = true;
    std::printf("FAIL_MSG=%s\n", e.what());
  }
  std::printf("BETA50_FAILED=%d\n", failed ? 1 : 0);
  core->sendMessage(M_SETDATA, "beta", message_data(unitval(0.65, U_UNITLESS)));
  core->sendMessage(M_SETDATA, "S", message_data(unitval(3.0, U_DEGC)));
  core->reset(0);
  core->run();
  std::printf("AGAIN_CO2_2300=%.17g\n", (double)core->sendMessage(M_GETDATA, "CO2_concentration", message_data(2300.0)));
  Core::delcore(idx);
  EXPECT(Core::getcore(idx) == nullptr);
  Core::delcore(idx); /* idempotent */

  /* ---- carbon tracking + a user constraint, like tests/testthat/test_tracking.R and
   * test_constraints.R drive them: setData before prepareToRun, getTrackingData after ---- */
  {
    Core tc(Logger::SEVERE, false, false);
    tc.init();
    INIToCoreReader(&tc).parse(ini);
    EXPECT(tc.getTrackingData().empty());
    tc.setData("core", "trackingDate", message_data(unitval(1990, U_UNITLESS)));
    for (int y = 2000; y <= 2010; ++y) /* dated entries accumulate, one setData per year */
      tc.setData("CH4", "CH4_constrain", message_data((double)y, unitval(1800.0, U_PPBV_CH4)));
    EXPECT(throws([&] { /* lo_warming_ratio is unitless (temperature_component.cpp:266-269) */
      tc.setData("temperature", "lo_warming_ratio", message_data(unitval(1.6, U_DEGC)));
    }));
    EXPECT(throws([&] {
      tc.setData("CH4", "CH4_constrain", message_data(2000.0, unitval(1800.0, U_PGC)));
    }));
    tc.prepareToRun();
    tc.run(2020);
    EXPECT((double)tc.sendMessage(M_GETDATA, "CH4_concentration", message_data(2005.0)) == 1800.0);
    EXPECT((double)tc.sendMessage(M_GETDATA, "CH4_concentration", message_data(2011.0)) != 1800.0);
    const std::string csv = tc.getTrackingData();
    EXPECT(csv.compare(0, 5, "year,") == 0);
    size_t rows = 0;
    for (char ch : csv) rows += ch == '\n';
    std::printf("TRACK_ROWS=%zu\n", rows);
    EXPECT(csv.find("\n1990,simpleNbox,atmos_co2,") != std::string::npos);
    EXPECT(csv.find("\n2020,ocean,deep,") != std::string::npos);
    EXPECT(csv.find("\n1989,") == std::string::npos && csv.find("\n2021,") == std::string::npos);
    std::printf("TRACK_CH4_2020=%.17g\n", (double)tc.sendMessage(M_GETDATA, "CH4_concentration", message_data(2020.0)));
  }

  /* ---- biomes (tests/testthat/test_biome.R territory): two biomes that split the global
   * pools 0.3 / 0.7, with their own beta, Q10 and warming factor -- the two_ssp245 case of
   * tests/golden/ref_biomes.npz ---- */
  {
    Core bc(Logger::SEVERE, false, false);
    bc.init();
    INIToCoreReader(&bc).parse(ini);
    EXPECT(bc.getBiomeList().size() == 1 && bc.getBiomeList()[0] == "global");
    bc.setBiomes({"boreal", "tropical"});
    EXPECT(bc.getBiomeList().size() == 2 && bc.getBiomeList()[1] == "tropical");
    bc.selectOutputs({"CO2_concentration", "veg_c", "boreal.veg_c", "tropical.veg_c"});
    EXPECT(throws([&] { bc.selectOutputs({"tundra.veg_c"}); })); /* not a biome */
    const struct { const char *name; double boreal, tropical; unit_types u; } in[] = {
        {"npp_flux0", 56.2 * 0.3, 56.2 * 0.7, U_PGC_YR}, {"veg_c", 550.0 * 0.3, 550.0 * 0.7, U_PGC},
        {"detritus_c", 55.0 * 0.3, 55.0 * 0.7, U_PGC},   {"soil_c", 917.0 * 0.3, 917.0 * 0.7, U_PGC},
        {"permafrost_c", 865.0, 0.0, U_PGC},             {"f_nppv", 0.30, 0.35, U_UNITLESS},
        {"f_nppd", 0.60, 0.60, U_UNITLESS},              {"f_litterd", 0.98, 0.95, U_UNITLESS},
        {"beta", 0.5, 0.7, U_UNITLESS},                  {"q10_rh", 2.2, 1.6, U_UNITLESS},
        {"warmingfactor", 1.8, 0.9, U_UNITLESS}};
    for (const auto &e : in) {
      bc.setData("simpleNbox", std::string("boreal.") + e.name, message_data(unitval(e.boreal, e.u)));
      bc.setData("simpleNbox", std::string("tropical.") + e.name, message_data(unitval(e.tropical, e.u)));
    }
    EXPECT(throws([&] { /* units are those of the plain name */
      bc.setData("simpleNbox", "boreal.veg_c", message_data(unitval(1.0, U_DEGC)));
    }));
    EXPECT(throws([&] { /* global and biome-specific data do not mix */
      bc.setData("simpleNbox", "beta", message_data(unitval(0.5, U_UNITLESS)));
    }));
    bc.run();
    std::printf("BIOME_CO2_2300=%.17g\n", (double)bc.sendMessage(M_GETDATA, "CO2_concentration", message_data(2300.0)));
    std::printf("BIOME_VEG_2100=%.17g\n", (double)bc.sendMessage(M_GETDATA, "veg_c", message_data(2100.0)));
    const unitval bv = bc.sendMessage(M_GETDATA, "boreal.veg_c", message_data(2100.0));
    EXPECT(bv.units() == U_PGC);
    std::printf("BIOME_BOREAL_VEG_2100=%.17g\n", (double)bv);
    const double tv = bc.sendMessage(M_GETDATA, "tropical.veg_c", message_data(2100.0));
    const double gv = bc.sendMessage(M_GETDATA, "veg_c", message_data(2100.0));
    EXPECT(std::fabs((double)bv + tv - gv) < 1e-9);
  }

  /* ---- the batch face: 4 members, per-member S ---- */
  EnsembleCore ens(4);
  INIToCoreReader(&ens).parse(ini);
  ens.selectOutputs({"CO2_concentration", "global_tas"});
  ens.setMembers("S", {2.0, 3.0, 4.5, 6.0}, U_DEGC);
  EXPECT(throws([&] { ens.setMembers("S", {2.0, 3.0, 4.5, 6.0}, U_PGC); }));
  ens.sendMessage(M_SETDATA, "beta", message_data(unitval(0.4, U_UNITLESS)), 2); /* member 2 only */
  ens.run(-1, false);
  std::vector<double> out(4 * 2);
  ens.fetch("global_tas", {2100.0, 2300.0}, out.data());
  for (int m = 0; m < 4; ++m) std::printf("ENS_TAS_2300_%d=%.17g\n", m, out[m * 2 + 1]);
  std::printf("ENS_CO2_2300_2=%.17g\n", (double)ens.sendMessage(M_GETDATA, "CO2_concentration", message_data(2300.0), 2));
  std::vector<int32_t> st, fy;
  ens.memberStatus(st, fy);
  for (int m = 0; m < 4; ++m) EXPECT(st[m] == 0);
  EXPECT(throws([&] { ens.fetch("RF_tot", {2100.0}, out.data()); })); /* not selected */
  ens.shutDown();
  ens.shutDown();

  std::printf("FAILURES=%d\n", failures);
  return failures ? 1 : 0;
}
