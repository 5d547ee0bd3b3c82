/* hx_ini.h -- Hector ini/csv input reader (see hx_ini.cpp) */
#ifndef HX_INI_H
#define HX_INI_H
#include <map>
#include <string>
#include <vector>

namespace hx {
struct IniInputs {
  int start_year = 0, end_year = 0;
  bool do_spinup = true;
  double tracking_date = 9999;
  std::string run_name;
  std::map<std::string, double> scalars;   /* engine parameter name -> value */
  std::vector<std::vector<double>> series; /* [RAW_COUNT][nrow] dense per-year values */
  std::map<int, std::vector<double>> constraints; /* CN_* -> [nrow], NaN = no entry */
  /* <biome>.<name> entries of [simpleNbox] (simpleNbox.cpp:201-236): the biomes in the order
   * they first appear, and their values */
  std::vector<std::string> biomes;
  std::vector<std::pair<std::string, double>> biome_scalars; /* "<biome>.<name>" -> value */
  std::string error;
  bool unsupported = false;
};
bool read_ini(const std::string &path, IniInputs &out);
}

/* library-internal: lets other translation units report an hx_create-stage error */
extern "C" void hx_set_create_error(const char *msg);
#endif
