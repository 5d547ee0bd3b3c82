/* include/compat/csv_tracking_visitor.hpp -- stands in for inst/include/csv_tracking_visitor.hpp +
 * src/csv_tracking_visitor.cpp:55-137: CSVFluxPoolVisitor writes
 * year,component,pool_name,pool_value,pool_units,source_name,source_fraction for every tracked
 * pool and source, from the tracking date on (empty file when tracking is off). */
#ifndef HECTOR_B200_COMPAT_CSV_TRACKING_VISITOR_HPP
#define HECTOR_B200_COMPAT_CSV_TRACKING_VISITOR_HPP
#include <ostream>

#include "core.hpp"

namespace hector_b200 {
class CSVFluxPoolVisitor : public AVisitor {
 public:
  CSVFluxPoolVisitor(std::ostream &outputStream, const bool printHeader = true)
      : csvFile(outputStream), header(printHeader) {}
  bool shouldVisit(const bool in_spinup, const double date) override {
    current_date = date;
    return !in_spinup;
  }
  void visit(Core *c) override {
    if (current_date < c->getTrackingDate()) return;
    if (header) {
      csvFile << "year,component,pool_name,pool_value,pool_units,source_name,source_fraction" << std::endl;
      header = false;
    }
    c->trackingRows(csvFile, (int)current_date);
  }

 private:
  std::ostream &csvFile;
  bool header;
  double current_date = 0.0;
};
} // namespace hector_b200
#endif
