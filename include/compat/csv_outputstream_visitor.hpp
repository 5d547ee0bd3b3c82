/* include/compat/csv_outputstream_visitor.hpp -- stands in for
 * inst/include/csv_outputstream_visitor.hpp + src/csv_outputstream_visitor.cpp:55-143.
 *
 * Writes the reference's outputstream_<run>.csv: a comment line, the header
 * year,run_name,spinup,component,variable,value,units and one row per variable and year, values
 * with the stream's default six significant digits (forcings with four, :148), components in the
 * order the reference visits them (its component map is sorted by name).  The rows are those of
 * the variables the engine records; quantities it does not carry (DIC, Omega, Revelle factors,
 * sea level ...) have no row.  The spin-up is not reported. */
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
    struct Row { const char *component, *variable; int precision; };
    static const Row rows[] = {
        {"CH4", "CH4_concentration", 6}, {"N2O", "N2O_concentration", 6},
        {"forcing", "RF_CH4", 4}, {"forcing", "RF_CO2", 4}, {"forcing", "RF_N2O", 4}, {"forcing", "RF_tot", 4},
        {"o3", "O3_concentration", 6},
        {"ocean", "DO_ocean_c", 6}, {"ocean", "HL_ocean_c", 6}, {"ocean", "IO_ocean_c", 6},
        {"ocean", "LL_ocean_c", 6}, {"ocean", "ocean_uptake", 6}, {"ocean", "HL_PCO2", 6},
        {"ocean", "LL_PCO2", 6}, {"ocean", "HL_pH", 6}, {"ocean", "LL_pH", 6}, {"ocean", "ocean_c", 6},
        {"simpleNbox", "NBP", 6}, {"simpleNbox", "NPP", 6}, {"simpleNbox", "RH", 6},
        {"simpleNbox", "CO2_concentration", 6}, {"simpleNbox", "atmos_co2", 6}, {"simpleNbox", "veg_c", 6},
        {"simpleNbox", "detritus_c", 6}, {"simpleNbox", "soil_c", 6}, {"simpleNbox", "permafrost_c", 6},
        {"simpleNbox", "thawedp_c", 6}, {"simpleNbox", "earth_c", 6},
        {"temperature", "global_tas", 6}, {"temperature", "gmst", 6}, {"temperature", "heatflux_mixed", 6},
        {"temperature", "heatflux_interior", 6}, {"temperature", "heatflux", 6},
        {"temperature", "land_tas", 6}, {"temperature", "sst", 6}};
    const std::string run_name = c->getRun_name();
    for (const Row &r : rows) {
      if (!c->isRecorded(r.variable)) continue;
      const unitval x = c->sendMessage(M_GETDATA, r.variable, message_data(current_date));
      const std::streamsize old = csvFile.precision(r.precision);
      csvFile << current_date << "," << run_name << "," << (spinup ? 1 : 0) << "," << r.component << ","
              << r.variable << "," << x.value(x.units()) << "," << x.unitsName() << std::endl;
      csvFile.precision(old);
    }
  }

 private:
  std::ostream &csvFile;
  double current_date = 0.0;
  bool spinup = false;
};
} // namespace hector_b200
#endif
