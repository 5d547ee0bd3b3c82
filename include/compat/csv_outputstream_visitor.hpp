/* include/compat/csv_outputstream_visitor.hpp -- stands in for
 * inst/include/csv_outputstream_visitor.hpp + src/csv_outputstream_visitor.cpp:55-343.
 *
 * Writes the reference's outputstream_<run>.csv: a comment line, the header
 * year,run_name,spinup,component,variable,value,units and the reference's rows of every model
 * year in the reference's order -- the components in the order the core runs them (OH, CH4,
 * ozone, N2O, the 26 halocarbons, ocean, simpleNbox, forcing from the base year on,
 * temperature), the forcings sorted by name like its std::map.  Three things the reference's
 * text does are kept because a reader of the file sees them: every value carries FOUR
 * significant digits (the forcing visitor sets precision(4) and returns before the base year
 * without restoring it, :143-147, and the spin-up comes first), the row "rh_ch4" prints RH's
 * value (:177), and the undated ocean_uptake row has the units "Pg C" (:268).  Pinned by
 * tests/golden/ref_outputstream_ssp245.txt (the unmodified reference's own file).
 * Rows the engine has no number for are left out: HL_downwelling, HL_Revelle, LL_Revelle,
 * atmos_c_residual, the sea-level component's; per-biome f_frozen / tempfert rows.  The spin-up
 * is not reported. */
#ifndef HECTOR_B200_COMPAT_CSV_OUTPUTSTREAM_VISITOR_HPP
#define HECTOR_B200_COMPAT_CSV_OUTPUTSTREAM_VISITOR_HPP
#include <ctime>
#include <ostream>
#include <string>

#include "h_util.hpp"

namespace hector_b200 {
class CSVOutputStreamVisitor : public AVisitor {
 public:
  CSVOutputStreamVisitor(std::ostream &outputStream, const bool printHeader = true) : csvFile(outputStream) {
    if (printHeader) {
      std::time_t now = std::time(nullptr);
      std::string when = std::ctime(&now);
      while (!when.empty() && when[when.size() - 1] == '\n') when.erase(when.size() - 1);
      csvFile << "# Output from " << MODEL_NAME << " version " << MODEL_VERSION << " on " << when << std::endl;
      csvFile << "year,run_name,spinup,component,variable,value,units" << std::endl;
    }
  }
  bool shouldVisit(const bool in_spinup, const double date) override {
    current_date = date;
    spinup = in_spinup;
    return true;
  }
  void visit(Core *c) override {
    static const char *const halo[] = {"CF4", "C2F6", "HFC23", "HFC32", "HFC4310", "HFC125", "HFC134a",
                                       "HFC143a", "HFC227ea", "HFC245fa", "SF6", "CFC11", "CFC12", "CFC113",
                                       "CFC114", "CFC115", "CCl4", "CH3CCl3", "HCFC22", "HCFC141b",
                                       "HCFC142b", "halon1211", "halon1301", "halon2402", "CH3Br", "CH3Cl"};
    struct Row { const char *component, *variable, *datum, *units; }; /* datum / units: null = as named */
    static const Row head[] = {{"OH", "TAU_OH", nullptr, nullptr}, {"CH4", "CH4_concentration", nullptr, nullptr},
                               {"ozone", "O3_concentration", nullptr, nullptr},
                               {"N2O", "N2O_concentration", nullptr, nullptr}};
    static const Row body[] = {
        {"ocean", "HL_ocean_uptake", nullptr, nullptr}, {"ocean", "LL_ocean_uptake", nullptr, nullptr},
        {"ocean", "DO_ocean_c", nullptr, nullptr}, {"ocean", "HL_ocean_c", nullptr, nullptr},
        {"ocean", "IO_ocean_c", nullptr, nullptr}, {"ocean", "LL_ocean_c", nullptr, nullptr},
        {"ocean", "HL_DIC", nullptr, nullptr}, {"ocean", "LL_DIC", nullptr, nullptr},
        {"ocean", "ocean_uptake", nullptr, "Pg C"},
        {"ocean", "HL_OmegaAr", nullptr, nullptr}, {"ocean", "LL_OmegaAr", nullptr, nullptr},
        {"ocean", "HL_OmegaCa", nullptr, nullptr}, {"ocean", "LL_OmegaCa", nullptr, nullptr},
        {"ocean", "HL_PCO2", nullptr, nullptr}, {"ocean", "LL_PCO2", nullptr, nullptr},
        {"ocean", "HL_pH", nullptr, nullptr}, {"ocean", "LL_pH", nullptr, nullptr},
        {"ocean", "HL_sst", nullptr, nullptr}, {"ocean", "LL_sst", nullptr, nullptr},
        {"ocean", "ocean_c", nullptr, nullptr}, {"ocean", "HL_CO3", nullptr, nullptr},
        {"ocean", "LL_CO3", nullptr, nullptr},
        {"simpleNbox", "NBP", nullptr, nullptr}, {"simpleNbox", "NPP", nullptr, nullptr},
        {"simpleNbox", "RH", nullptr, nullptr}, {"simpleNbox", "rh_det", nullptr, nullptr},
        {"simpleNbox", "rh_soil", nullptr, nullptr}, {"simpleNbox", "rh_ch4", "RH", nullptr},
        {"simpleNbox", "CO2_concentration", nullptr, nullptr}, {"simpleNbox", "atmos_co2", nullptr, nullptr},
        {"simpleNbox", "veg_c", nullptr, nullptr}, {"simpleNbox", "detritus_c", nullptr, nullptr},
        {"simpleNbox", "soil_c", nullptr, nullptr}, {"simpleNbox", "permafrost_c", nullptr, nullptr},
        {"simpleNbox", "thawedp_c", nullptr, nullptr}, {"simpleNbox", "f_frozen", nullptr, nullptr},
        {"simpleNbox", "earth_c", nullptr, nullptr}};
    static const Row tail[] = {
        {"temperature", "global_tas", nullptr, nullptr}, {"temperature", "gmst", nullptr, nullptr},
        {"temperature", "heatflux_mixed", nullptr, nullptr}, {"temperature", "heatflux_interior", nullptr, nullptr},
        {"temperature", "heatflux", nullptr, nullptr}, {"temperature", "land_tas", nullptr, nullptr},
        {"temperature", "sst", nullptr, nullptr}};
    /* ForcingComponent::forcings_t is a std::map: the keys in ASCII order */
    static const char *const agents[] = {
        "RF_BC", "RF_C2F6", "RF_CCl4", "RF_CF4", "RF_CFC11", "RF_CFC113", "RF_CFC114", "RF_CFC115",
        "RF_CFC12", "RF_CH3Br", "RF_CH3CCl3", "RF_CH3Cl", "RF_CH4", "RF_CO2", "RF_H2O_strat",
        "RF_HCFC141b", "RF_HCFC142b", "RF_HCFC22", "RF_HFC125", "RF_HFC134a", "RF_HFC143a",
        "RF_HFC227ea", "RF_HFC23", "RF_HFC245fa", "RF_HFC32", "RF_HFC4310", "RF_N2O", "RF_NH3",
        "RF_O3_trop", "RF_OC", "RF_SF6", "RF_SO2", "RF_aci", "RF_albedo", "RF_halon1211",
        "RF_halon1301", "RF_halon2402", "RF_misc", "RF_tot", "RF_vol"};
    const std::string run_name = c->getRun_name();
    const std::streamsize old = csvFile.precision(4);
    auto emit = [&](const std::string &component, const std::string &variable, const std::string &datum,
                    const char *units) {
      unitval x;
      try { /* a variable this run does not record (selected outputs, several biomes) has no row */
        x = c->sendMessage(M_GETDATA, datum, message_data(current_date));
      } catch (const h_exception &) {
        return;
      }
      csvFile << current_date << "," << run_name << "," << (spinup ? 1 : 0) << "," << component << ","
              << variable << "," << x.value(x.units()) << "," << (units ? std::string(units) : x.unitsName())
              << std::endl;
    };
    auto emit_rows = [&](const Row *rows, size_t n) {
      for (size_t i = 0; i < n; ++i)
        emit(rows[i].component, rows[i].variable, rows[i].datum ? rows[i].datum : rows[i].variable, rows[i].units);
    };
    emit_rows(head, sizeof head / sizeof head[0]);
    for (const char *g : halo) emit(std::string(g) + "_halocarbon", "hc_concentration", std::string(g) + "_concentration", nullptr);
    emit_rows(body, sizeof body / sizeof body[0]);
    const double baseyear = (double)c->sendMessage(M_GETDATA, "baseyear");
    if (current_date >= baseyear)
      for (const char *a : agents) {
        /* the halocarbons' entries are the adjusted forcings relative to the base year: the
         * engine's Fadj<gas> (its RF_<gas> is the absolute forcing) */
        std::string datum = a;
        for (const char *g : halo)
          if (datum == std::string("RF_") + g) datum = std::string("Fadj") + g;
        emit("forcing", a, datum, "W/m2");
      }
    emit_rows(tail, sizeof tail / sizeof tail[0]);
    csvFile.precision(old);
  }

 private:
  std::ostream &csvFile;
  double current_date = 0.0;
  bool spinup = false;
};
} // namespace hector_b200
#endif
