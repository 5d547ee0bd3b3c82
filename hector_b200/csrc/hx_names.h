/* hx_names.h -- name tables: the reference's capability / input strings
 * (inst/include/component_data.hpp) mapped to the engine's dense ids. */
#ifndef HX_NAMES_H
#define HX_NAMES_H

#include <cmath>

#include "hx_layout.h"

namespace hx {

static const char *const kHaloNames[HX_NHALO] = {
    "CF4", "C2F6", "HFC23", "HFC32", "HFC4310", "HFC125", "HFC134a", "HFC143a", "HFC227ea",
    "HFC245fa", "SF6", "CFC11", "CFC12", "CFC113", "CFC114", "CFC115", "CCl4", "CH3CCl3",
    "HCFC22", "HCFC141b", "HCFC142b", "halon1211", "halon1301", "halon2402", "CH3Cl", "CH3Br"};

/* raw series names in RAW_* order; the 26 halocarbon ones are "<gas>_emissions" */
static const char *const kRawNames[RAW_HALO0] = {
    "ffi_emissions", "daccs_uptake", "luc_emissions", "luc_uptake", "CH4_emissions", "CH4N",
    "NOX_emissions", "CO_emissions", "NMVOC_emissions", "BC_emissions", "OC_emissions",
    "SO2_emissions", "NH3_emissions", "SV", "RF_albedo", "RF_misc", "N2O_emissions",
    "N2O_natural_emissions"};

struct ParamInfo {
  const char *name;
  double dflt; /* inst/input/hector_ssp245.ini or the compiled-in default */
};
/* PI_* order */
static const ParamInfo kParams[PI_COUNT] = {
    {"S", 3.0}, {"diff", 1.042}, {"qco2", 3.75},
    {"q10_rh", 1.2}, {"C0", 277.15},
    {"veg_c", 550}, {"detritus_c", 55}, {"soil_c", 917}, {"permafrost_c", 865},
    {"tt", 72000000}, {"tu", 49000000}, {"twi", 12500000}, {"tid", 200000000},
    {"preind_surface_c", 900}, {"preind_interdeep_c", 37100},
    {"dt", 0.25}, {"eps_spinup", 0.001},
    {"aero_scalar", 1.0}, {"vol_scalar", 1.0}, {"delta_co2", 0.05}, {"delta_ch4", -.14},
    {"delta_n2o", 0.07},
    {"rho_bc", 0.06386286}, {"rho_oc", -0.006407143}, {"rho_so2", -7.469841e-06},
    {"rho_nh3", -0.002146032},
    {"M0", 731.41}, {"Tsoil", 120}, {"Tstrat", 150}, {"UC_CH4", 2.78},
    {"TOH0", 9.6}, {"CNOX", 8.4e-3}, {"CCO", -1.575e-4}, {"CNMVOC", -4.725e-4}, {"CCH4", -0.32},
    {"PO3", 30.0},
    {"N0", 273.87},
    {"lo_warming_ratio", 0.0},
    {"beta", 0.65}, {"warmingfactor", 1.0}, {"pf_mu", 1.67}, {"pf_sigma", 0.986},
    {"eps_rel", 1.0e-6}, {"eps_abs", 1.0e-6}, {"npp_flux0", 56.2}, {"f_nppv", 0.35}, {"f_nppd", 0.60},
    {"f_litterd", 0.98}, {"fpf_static", 0.74}, {"rh_ch4_frac", 0.023}};

/* BP_* order: the per-biome inputs, named as the reference names them after the "<biome>."
 * prefix (simpleNbox.cpp:281-396); NaN = must be given (simpleNbox-runtime.cpp:66-101), a number
 * = the default prepareToRun fills in (:102-143) */
static const ParamInfo kBiomeParams[BP_COUNT] = {
    {"veg_c", NAN}, {"detritus_c", NAN}, {"soil_c", NAN}, {"permafrost_c", NAN},
    {"npp_flux0", NAN}, {"beta", NAN}, {"q10_rh", NAN}, {"warmingfactor", 1.0},
    {"f_nppv", NAN}, {"f_nppd", NAN}, {"f_litterd", NAN}, {"rh_ch4_frac", 0.023},
    {"pf_mu", 1.67}, {"pf_sigma", 0.986}, {"fpf_static", 0.74}};

/* BO_* order: per-biome outputs and the BF_* field each one reports */
static const char *const kBiomeOutNames[BO_COUNT] = {"veg_c", "detritus_c", "soil_c",
                                                     "permafrost_c", "thawedp_c", "NPP", "RH"};

/* parameters that influence the spin-up / alkalinity equilibration; if all of them are
 * scalars the spin-up is computed once and broadcast */
static const int kSpinupParams[] = {
    PI_F_NPPV, PI_F_NPPD, PI_F_LITTERD, PI_NPP_FLUX0, PI_C0, PI_VEG_C0, PI_DET_C0, PI_SOIL_C0,
    PI_PERMAFROST_C0, PI_FPF_STATIC, PI_RH_CH4_FRAC, PI_TT, PI_TU, PI_TWI, PI_TID, PI_PREIND_SURF,
    PI_PREIND_ID, PI_EPS_ABS, PI_EPS_REL, PI_DT, PI_EPS_SPINUP};

/* OUT_* order */
static const char *const kOutNames[OUT_COUNT] = {
    "CO2_concentration", "global_tas", "RF_tot", "RF_CO2", "heatflux", "ocean_c", "HL_pH",
    "atmos_co2", "sst", "permafrost_c", "CH4_concentration", "N2O_concentration",
    "O3_concentration", "land_tas", "veg_c", "detritus_c", "soil_c", "thawedp_c", "earth_c", "NBP",
    "ocean_uptake", "LL_pH", "HL_PCO2", "LL_PCO2", "HL_ocean_c", "LL_ocean_c", "IO_ocean_c",
    "DO_ocean_c", "RF_CH4", "RF_N2O", "rh_ch4", "NPP", "RH", "gmst", "ocean_tas",
    "heatflux_mixed", "heatflux_interior", "ocean_timesteps",
    "HL_ocean_uptake", "LL_ocean_uptake", "rh_det", "rh_soil"};

/* halocarbon defaults (26 x [<gas>_halocarbon] sections of hector_ssp245.ini) */
static const double kHaloTau[HX_NHALO] = {50000.0, 10000.0, 228.0, 5.4, 17.0, 30.0, 14.0, 51.0,
                                          36.0, 7.9, 3200.0, 52.0, 102.0, 93.0, 189, 540, 32.0,
                                          5.0, 11.9, 9.4, 18.0, 16.0, 72.0, 28.0, 0.9, 0.8};
static const double kHaloRho[HX_NHALO] = {
    0.000099, 0.000261, 0.000191, 0.000111, 0.000357, 0.000234, 0.000167, 0.000168, 0.000273,
    0.000245, 0.000567, 0.000259, 0.00032, 0.000301, 0.000314, 0.000246, 0.000166, 0.000065,
    0.000214, 0.000161, 0.000193, 0.00003, 0.000299, 0.000312, 0.000005, 0.000004};
static const double kHaloMolarMass[HX_NHALO] = {
    88.0043, 138.01, 70.0, 52.0, 252.0, 120.02, 102.02, 84.04, 170.03, 134.0, 146.06, 137.35, 120.9,
    187.35, 170.9, 154.45, 153.8, 133.35, 86.45, 116.9, 100.45, 165.35, 148.9, 259.8, 50.45, 50.45};
static const double kHaloDelta[HX_NHALO] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.13, 0.13, 0, 0, 0, 0,
                                            0, 0, 0, 0, 0, 0, 0, 0, 0};
static const double kHaloH0[HX_NHALO] = {35.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                         0, 0, 0, 0, 0, 0, 504.0, 5.8};

} // namespace hx
#endif
