/* oracle/hector_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-member, CPU restatement of Hector's per-year coupled hot path
 * (JGCRI/hector v3.5.0).  It follows the reference statement by statement (same
 * evaluation order, same retry / doomed-attempt control flow, cold-start Newton) and is
 * pinned against (a) the reference's golden file tests/testthat/compdata/hector_comp.csv
 * and (b) the unmodified reference built as oracle/_ref/libhector_ref.so.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link or call
 * this.  The product (hector_b200/) never does.
 */
#ifndef HECTOR_ORACLE_H
#define HECTOR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HO_NHALO 26
#define HO_NPOOL 11 /* tracked pools */
#define HO_NSRC 12  /* source names: the 11 pool names + "untracked" */

/* raw scenario series: one dense value per integer year start..end, as the reference's
 * tseries::get(year) would return it (exact key, linear interpolation, or constant) */
enum {
  HO_RAW_FFI = 0, HO_RAW_DACCS, HO_RAW_LUC_E, HO_RAW_LUC_U,
  HO_RAW_CH4_E, HO_RAW_CH4N, HO_RAW_NOX, HO_RAW_CO, HO_RAW_NMVOC,
  HO_RAW_BC, HO_RAW_OC, HO_RAW_SO2, HO_RAW_NH3, HO_RAW_SV, HO_RAW_ALBEDO, HO_RAW_MISC,
  HO_RAW_N2O_E, HO_RAW_N2O_NAT,
  HO_RAW_HALO0, /* 26 halocarbon emission series, order of forcing_component.cpp:402-409 */
  HO_NRAW = HO_RAW_HALO0 + HO_NHALO
};

/* per-year outputs (index y - start - 1) */
enum {
  HO_OUT_CO2 = 0,      /* CO2_concentration, ppmv */
  HO_OUT_TAS,          /* global_tas */
  HO_OUT_RF_TOT,
  HO_OUT_RF_CO2,
  HO_OUT_HEATFLUX,
  HO_OUT_OCEAN_C,
  HO_OUT_HL_PH,
  HO_OUT_ATMOS_C,      /* atmos_co2, Pg C */
  HO_OUT_SST,
  HO_OUT_PERMAFROST_C,
  HO_OUT_CH4,
  HO_OUT_N2O,
  HO_OUT_O3,
  HO_OUT_LAND_TAS,
  HO_OUT_VEG_C,
  HO_OUT_DETRITUS_C,
  HO_OUT_SOIL_C,
  HO_OUT_THAWEDP_C,
  HO_OUT_EARTH_C,
  HO_OUT_NBP,
  HO_OUT_OCEAN_UPTAKE, /* annualflux_sum (atm_ocean_flux) */
  HO_OUT_LL_PH,
  HO_OUT_PCO2_HL,
  HO_OUT_PCO2_LL,
  HO_OUT_CARBON_HL,
  HO_OUT_CARBON_LL,
  HO_OUT_CARBON_IO,
  HO_OUT_CARBON_DO,
  HO_OUT_RF_CH4,
  HO_OUT_RF_N2O,
  HO_OUT_RH_CH4,       /* end-of-year annual permafrost CH4 flux (what CH4 reads next year) */
  HO_OUT_NPP,          /* final_npp: NPP of the year's last stash, Pg C/yr (simpleNbox.cpp:684-686) */
  HO_OUT_RH,           /* final_rh: detritus + soil + thawed-permafrost CO2 and CH4 respiration */
  HO_OUT_GMST,         /* gmst: flnd T_land + (1 - flnd) SST (temperature_component.cpp:506-507) */
  HO_OUT_OCEAN_TAS,    /* ocean_tas: bsi SST (:716-717) */
  HO_OUT_FLUX_MIXED,   /* heatflux_mixed */
  HO_OUT_FLUX_INTERIOR,/* heatflux_interior */
  HO_OUT_TIMESTEPS,    /* ocean sub-steps (stashes) in the year */
  HO_NOUT
};

/* member status (reference throws h_exception where we return non-zero) */
enum {
  HO_OK = 0,
  HO_ERR_NEGATIVE = 1,     /* fluxpool.hpp:100-102,121-123 "may not be negative" */
  HO_ERR_MASS = 2,         /* simpleNbox-runtime.cpp:556-563 */
  HO_ERR_RETRIES = 3,      /* carbon-cycle-solver.cpp:294 */
  HO_ERR_NOROOT = 4,       /* newton bracket lost */
  HO_ERR_YEARFRACTION = 5, /* simpleNbox-runtime.cpp:275, ocean_component.cpp:665 */
  HO_ERR_CO2SARF = 6,      /* forcing_component.cpp:353 */
  HO_ERR_STEPPER = 7,      /* odeint: 500 failed step-size searches */
  HO_ERR_TRACKING = 8,     /* fluxpool.hpp:105-112 source fractions out of range / tracking mismatch */
  HO_ERR_UNSUPPORTED = 9   /* an input combination this restatement does not cover (biomes + tracking) */
};

#define HO_MAX_BIOMES 4
/* what the reference keeps per biome (simpleNbox.hpp:236-290) */
typedef struct {
  double veg_c, detritus_c, soil_c, permafrost_c;
  double npp_flux0, beta, q10_rh, warmingfactor, f_nppv, f_nppd, f_litterd;
  double rh_ch4_frac, pf_mu, pf_sigma, fpf_static;
} ho_biome;

typedef struct {
  /* [core] */
  int start_year, end_year, do_spinup, max_spinup;
  /* [temperature] */
  double S, diff, qco2;
  /* [simpleNbox] */
  double beta, q10_rh, f_nppv, f_nppd, f_litterd, npp_flux0, C0;
  double veg_c, detritus_c, soil_c, permafrost_c;
  double warmingfactor, rh_ch4_frac, pf_mu, pf_sigma, fpf_static;
  /* [ocean] */
  double tt, tu, twi, tid, preind_C_surface, preind_C_ID;
  int spinup_chem;
  /* [carbon-cycle-solver] */
  double eps_abs, eps_rel, dt, eps_spinup;
  /* [forcing] */
  double baseyear, aero_scalar, vol_scalar, delta_co2, delta_ch4, delta_n2o;
  double rho_bc, rho_oc, rho_so2, rho_nh3;
  /* [CH4] [OH] [ozone] [N2O] */
  double M0, Tsoil, Tstrat, UC_CH4;
  double TOH0, CNOX, CCO, CNMVOC, CCH4;
  double PO3;
  double N0, UC_N2O, TN2O0;
  /* 26 x [<gas>_halocarbon] */
  double halo_tau[HO_NHALO], halo_rho[HO_NHALO], halo_delta[HO_NHALO], halo_H0[HO_NHALO],
      halo_molarMass[HO_NHALO];
  /* [temperature] lo_warming_ratio: 0 = off; otherwise land / ocean-air / sea-surface
   * temperatures as other components and callers see them are re-derived from global tas with
   * this land-ocean warming ratio (temperature_component.cpp:586-622, 722-739) */
  double lo_warming_ratio;
  /* biomes (simpleNbox.cpp:201-236): n_biomes <= 1 = the single "global" biome described by the
   * scalar [simpleNbox] fields above; otherwise biome[0..n_biomes) in biome_list (creation)
   * order replace them.  biome_order lists the same indices sorted by biome NAME: the
   * reference keeps pools in std::map<string, ...>, so sum_map() adds in that order
   * (simpleNbox.cpp:428-438) while the per-biome loops run in creation order. */
  int n_biomes;
  int biome_order[HO_MAX_BIOMES];
  ho_biome biome[HO_MAX_BIOMES];
} ho_params;

/* User constraints: dense per-model-year series [nrow] (row = year - start_year), NaN = no
 * entry for that year; NULL = no such constraint.
 *   co2, nbp, ch4, n2o, halo[nrow][26]  applied in the years that have an entry
 *        (tseries::exists(year): simpleNbox-runtime.cpp:345, 568, 871; ch4_component.cpp:141,
 *        156; n2o_component.cpp:141, 157; halocarbon_component.cpp:189)
 *   rf_tot   applied in every year <= rf_tot_last_row (forcing_component.cpp:498); the caller
 *            fills the rows below the first entry flat and interpolates gaps linearly, which
 *            is what Ftot_constrain.get() returns
 *   tas      applied for tas_first_row <= row <= tas_last_row (temperature_component.cpp:
 *            510-512), gaps interpolated by the caller */
typedef struct {
  const double *co2, *nbp, *ch4, *n2o, *halo, *rf_tot, *tas;
  int rf_tot_last_row, tas_first_row, tas_last_row;
} ho_constraints;

typedef struct {
  uint64_t rhs_evals, steps_accepted, steps_rejected, integrate_calls, newton_iterations,
      newton_calls, spinup_steps;
} ho_counters;

/* post-spin-up snapshot (what a shared spin-up broadcasts), for cross-checks */
typedef struct {
  double atmos, veg, det, soil, permafrost, thawed, earth;
  double ocean[4];   /* HL, LL, intermediate, deep */
  double alk_HL, alk_LL; /* after chem_equilibrate in the first model year */
  int spinup_steps;
} ho_spinup_state;

/* defaults = inst/input/hector_ssp245.ini + compiled-in defaults */
void ho_default_params(ho_params *p);

/* Run one member start_year -> run_to (<= end_year).
 *   raw   : [end_year-start_year+1][HO_NRAW] row-major scenario table
 *   out   : [HO_NOUT][nyears_cap], column i = year start_year+1+i (NaN-filled beyond failure)
 * returns member status (HO_OK or HO_ERR_*); *fail_year = year being computed on failure. */
int ho_run_member(const ho_params *p, const double *raw, int run_to, double *out,
                  int nyears_cap, int *fail_year, ho_counters *counters,
                  ho_spinup_state *spin);

/* Same with carbon tracking switched on in year `tracking_date` (core.cpp:228-235; 9999 =
 * never).  For every year >= tracking_date:
 *   track_frac [nyears_cap][HO_NPOOL][HO_NSRC]  source fractions of the 11 tracked pools
 *   track_mask [nyears_cap][HO_NPOOL]           bit s set = source s is a key of the pool's map
 * pool / source order: atmos_co2 earth_c veg_c detritus_c soil_c permafrost_c thawedp_c HL LL
 * intermediate deep (+ source 11 "untracked").  Either pointer may be NULL. */
int ho_run_member_tracked(const ho_params *p, const double *raw, int run_to, double *out,
                          int nyears_cap, int *fail_year, ho_counters *counters,
                          ho_spinup_state *spin, int tracking_date, double *track_frac,
                          uint32_t *track_mask);

/* everything at once: optional constraints and carbon tracking */
int ho_run_member_ex(const ho_params *p, const double *raw, const ho_constraints *cn, int run_to,
                     double *out, int nyears_cap, int *fail_year, ho_counters *counters,
                     ho_spinup_state *spin, int tracking_date, double *track_frac,
                     uint32_t *track_mask);
/* a run that also reports every biome's own pools and final fluxes:
 * bio_out[biome][k][nyears_cap], k = veg_c detritus_c soil_c permafrost_c thawedp_c NPP RH
 * (what getData("<biome>.<name>", date) returns, simpleNbox.cpp:533-697) */
#define HO_NBIOME_OUT 7
int ho_run_member_biomes(const ho_params *p, const double *raw, const ho_constraints *cn, int run_to,
                         double *out, int nyears_cap, int *fail_year, double *bio_out);
double ho_gas_series_constrained(const ho_params *p, const double *raw, const ho_constraints *cn,
                                 double *n2o, double *halo_rf);

/* fluxpool operator+ on explicit maps (unit-test hook, cf. src/unit-testing/test_tracking.cpp) */
int ho_tm_add(double a, double *fa, uint32_t *mask_a, double b, const double *fb, uint32_t mask_b);

/* carbonate chemistry spot check (ocean_csys.cpp:166-366): returns [H+]; fills 8 outputs
 * {pH, PCO2o, Tr, K0, CO3, TCO2o, HCO3, OmegaCa} and the Newton iteration count */
double ho_csys(double Tbox, double carbon_pgc, double alk, double volume, double S, double U,
               double *out8, int *iters);

/* member-independent series (n2o_component.cpp:150-191, halocarbon_component.cpp:181-229):
 * n2o[nrow] ppbv, halo_rf[nrow][26] W/m2 (row 0 = start year) */
void ho_gas_series(const ho_params *p, const double *raw, double *n2o, double *halo_rf);

/* DOECLIM kernel Ker[ns] (temperature_component.cpp:303-371) */
void ho_doeclim_kernel(double diff, int ns, double *ker);

#ifdef __cplusplus
}
#endif
#endif
