/* hx_layout.h -- data layout shared by the host engine and the device kernels.
 *
 * Per-member data is structure-of-arrays, tiled by CTA ("CTA-tiled SoA"): the HX_TILE = 128
 * members of one CTA own a contiguous block
 *   params + derived constants  P[tile][PI_COUNT | DI_COUNT][128]     state  S[tile][SI_COUNT][128]
 *   histories  sst_hist / tland_hist [tile][nrow][128],  ker [tile][HX_KER_ROWS(nrow)][128]
 *   scratch    conv [tile][HX_SLAB_YEARS][128]   (per-slab partial convolution sums)
 * so (a) a warp touches 32 consecutive doubles (256 B) per access, (b) every field of a thread
 * sits at a COMPILE-TIME offset (field * 1 KiB) from one base pointer -- no per-field address
 * registers or 64-bit index arithmetic in the kernels -- and (c) consecutive history rows of a
 * tile are contiguous (1 KiB per row) for bulk copies.  Outputs stay O[nsel][nyears][Mpad]
 * (what fetches and the NCCL all-gather consume).  Scenario tables are
 * [scenario][row][SC_STRIDE] doubles, row = year - start_year, staged to shared memory in
 * slabs of consecutive rows with one bulk copy each.
 */
#ifndef HX_LAYOUT_H
#define HX_LAYOUT_H

#include <stdint.h>

#define HX_NHALO 26
#define HX_TILE 128 /* members per tile = threads per CTA (HX_BLOCK) */
#define HX_SLAB_YEARS 16 /* years per work item = scenario rows staged per bulk copy */
/* rows of the DOECLIM lag-kernel table K(0..nrow), zero-padded so the slab prepass of the
 * convolution may read K(j) up to j = nrow + HX_SLAB_YEARS without a bounds test */
#define HX_KER_ROWS(nrow) ((nrow) + 1 + HX_SLAB_YEARS)

/* element index of (field, member) in a tiled array with `nfields` rows per tile */
#define HX_TILED(field, m, nfields) \
  ((((size_t)(m) / HX_TILE) * (size_t)(nfields) + (size_t)(field)) * HX_TILE + ((size_t)(m) % HX_TILE))

/* ---- raw scenario series (what callers hand in), reference input names ---- */
enum {
  RAW_FFI = 0, RAW_DACCS, RAW_LUC_E, RAW_LUC_U, RAW_CH4_E, RAW_CH4N, RAW_NOX, RAW_CO, RAW_NMVOC,
  RAW_BC, RAW_OC, RAW_SO2, RAW_NH3, RAW_SV, RAW_ALBEDO, RAW_MISC, RAW_N2O_E, RAW_N2O_NAT,
  RAW_HALO0,
  RAW_COUNT = RAW_HALO0 + HX_NHALO
};

/* ---- device scenario table columns ---- */
enum {
  SC_FFI = 0, SC_DACCS, SC_LUC_E, SC_LUC_U, SC_CH4_E, SC_CH4N, SC_NOX, SC_CO, SC_NMVOC,
  SC_BC, SC_OC, SC_SO2, SC_NH3, SC_SV, SC_ALBEDO, SC_MISC,
  SC_N2O,   /* N2O concentration, host-precomputed (member independent) */
  SC_HALO0, /* 26 halocarbon forcings, host-precomputed */
  /* user constraints of the year (NaN = none): atmospheric CO2 [ppmv], CH4 [ppbv], total
   * forcing [W/m2], global mean temperature [degC], net biome production [Pg C/yr]; N2O and halocarbon constraints are already
   * folded into SC_N2O / SC_HALO0.. on the host */
  SC_C_CO2 = SC_HALO0 + HX_NHALO, SC_C_CH4, SC_C_RFTOT, SC_C_TAS, SC_C_NBP,
  /* member-independent pieces of the forcing, host-precomputed with the C library (the doubles
   * the reference computes): -aci_beta log(1 + SO2 / s_SO2 + (BC + OC) / s_BCOC)
   * (forcing_component.cpp:465-467) and sqrt(N2O concentration) */
  SC_ACI, SC_SQRT_N2O,
  SC_USED,                       /* 50 */
  SC_STRIDE = 50                 /* 400 B per row: multiple of 16 B for cp.async.bulk */
};

/* ---- constraint series (what callers hand in), reference input names ---- */
enum {
  CN_CO2 = 0, CN_NBP, CN_CH4, CN_N2O, CN_RFTOT, CN_TAS, CN_HALO0,
  CN_COUNT = CN_HALO0 + HX_NHALO
};

/* ---- per-member parameters ----
 * The last HX_HOT_PI of them and the first HX_HOT_DI derived constants (below) are contiguous in
 * the tiled array P | D: the run kernel keeps that stretch -- what every sub-step and stash of
 * the year reads -- in shared memory (HX_HOT_*). */
enum {
  PI_S = 0, PI_DIFF, PI_QCO2,
  PI_Q10, PI_C0,
  PI_VEG_C0, PI_DET_C0, PI_SOIL_C0, PI_PERMAFROST_C0,
  PI_TT, PI_TU, PI_TWI, PI_TID, PI_PREIND_SURF, PI_PREIND_ID,
  PI_DT, PI_EPS_SPINUP,
  PI_AERO, PI_VOL, PI_DELTA_CO2, PI_DELTA_CH4, PI_DELTA_N2O,
  PI_RHO_BC, PI_RHO_OC, PI_RHO_SO2, PI_RHO_NH3,
  PI_M0, PI_TSOIL, PI_TSTRAT, PI_UC_CH4, PI_TOH0, PI_CNOX, PI_CCO, PI_CNMVOC, PI_CCH4, PI_PO3,
  PI_N0,
  PI_LO_RATIO, /* [temperature] lo_warming_ratio, 0 = off */
  /* the hot stretch */
  PI_BETA, PI_WARMINGFACTOR, PI_PF_MU, PI_PF_SIGMA, /* slowparameval, once a year */
  PI_EPS_REL, PI_EPS_ABS, PI_NPP_FLUX0, PI_F_NPPV, PI_F_NPPD, PI_F_LITTERD, PI_FPF_STATIC, PI_RH_CH4_FRAC,
  PI_COUNT
};
#define HX_HOT_PI 12
#define HX_HOT_DI 8                       /* DI_K_LL_HL .. DI_K_DO_IO, DI_LNQ10 */
#define HX_HOT_FIRST (PI_COUNT - HX_HOT_PI) /* first field of the stretch in P | D */
#define HX_HOT_COUNT (HX_HOT_PI + HX_HOT_DI)

/* ---- per-member dynamic state ----
 * The first SI_REG_COUNT fields live in registers while a work item runs (they are read once
 * when it starts and written once when it ends), so the run kernel's shared-memory copy of the
 * state leaves them out; the first SI_SPINUP_ROWS fields are what the spin-up touches. */
enum {
  SI_ATMOS = 0, SI_VEG, SI_DET, SI_SOIL, SI_PERMAFROST, SI_THAWED, SI_EARTH,
  SI_BOX_HL, SI_BOX_LL, SI_BOX_IO, SI_BOX_DO,
  SI_MAX_TIMESTEP, SI_TIMEOUT, SI_SOLVER_DT,
  SI_ALK_HL, SI_ALK_LL, SI_H_HL, SI_H_LL,        /* alkalinity, last [H+] root (warm start) */
  SI_TEMPFERTS, SI_F_FROZEN, SI_CUM_LUC_VA, SI_EOS_VEGC, SI_MASSTOT, SI_CUM_PF_CH4, SI_RH_CH4,
  SI_LASTFLUX_ANN,
  SI_CH4, SI_TLAND, SI_SST, SI_HEAT_MIXED, SI_HEAT_INTERIOR, SI_RF_PREV,
  SI_BASE_TOT, SI_BASE_CO2, SI_BASE_CH4, SI_BASE_N2O,
  SI_TLAND_WSUM, SI_TLAND_WCOMP, /* 200-year land-temperature window: compensated running sum */
  SI_DPAST_RAW, /* last year's unscaled DOECLIM convolution sum_i sst[i] K(t-i+1): next year's
                   interior-flux sum is this plus one term */
  /* per-year scratch (slowparameval results, emissions, annual sums) parked between phases */
  SI_X_CO2FERT, SI_X_TFD, SI_X_TFS, SI_X_FNEWTHAW, SI_X_NPPLUC, SI_X_FFI, SI_X_DACCS, SI_X_NBP,
  SI_X_FLUXSUM,
  SI_TLAND_C, SI_SST_C, /* land / sea-surface temperature as the carbon cycle sees them: DOECLIM's
                           own, or re-derived from tas with lo_warming_ratio (constraint builds) */
  SI_X_NPP, SI_X_RH, /* final_npp / final_rh of the year's last stash (outputs NPP, RH) */
  SI_X_C_CO2, /* this year's CO2 constraint (NaN = none), read by the year's last stash */
  SI_X_C_NBP0, SI_X_C_NBP1, /* NBP constraints of year y-1 and y: round(t) picks one */
  /* the solver's own thawed-permafrost and ocean totals after the last sub-step: a sub-step that
   * follows a stash without a retry continues from them, not from the pools (the stash zeroes a
   * thawed pool below 1e-10 and the four boxes sum to the solver's total only to the last ulps;
   * carbon-cycle-solver.cpp:232, 279) */
  SI_X_SOLVER_TPF, SI_X_SOLVER_OCEAN,
  /* log(CH4) and log(CO2 / C0) as the year's end left them: the OH lifetime and the CO2
   * fertilisation of the next year take the logarithms of the same numbers */
  SI_LOG_CH4, SI_LOG_CO2R,
  SI_COUNT
};
#define SI_REG_COUNT 14   /* SI_ATMOS .. SI_SOLVER_DT */
#define SI_SPINUP_ROWS 26 /* SI_ATMOS .. SI_LASTFLUX_ANN */

/* ---- per-member scratch rows X[tile][XS_COUNT][128] in global memory (allocated only when one
 * of OUT_UPTAKE_HL .. OUT_RH_SOIL is recorded): the per-box air-sea flux sums of the year
 * (ocean_component.cpp:737-738) and the last stash's detritus / soil respiration
 * (simpleNbox-runtime.cpp:445-446), written by the stashes, read by the year's output stage.  The
 * shared-memory state block has no room for them and they are off the year's critical path. */
enum { XS_UPTAKE_HL = 0, XS_UPTAKE_LL, XS_RH_DET, XS_RH_SOIL, XS_COUNT };

/* ---- per-member derived constants (set-up kernel) ---- */
enum {
  DI_K_LL_HL = 0, DI_K_LL_IO, DI_K_HL_DO, DI_K_IO_LL, DI_K_IO_HL, DI_K_IO_DO, DI_K_DO_IO,
  DI_LNQ10,       /* log(q10_rh): pow(q10, x) is evaluated as exp(x * lnq10) */
  DI_A0, DI_A1, DI_A2, DI_A3, DI_IB0, DI_IB1, DI_IB2, DI_IB3,
  DI_TAUCFL, DI_TAUKLS, DI_TAUCFS, DI_TAUKSL,
  DI_SQDT_TAUDIF, /* pow(dt/taudif, 0.5)            temperature_component.cpp:491 */
  DI_HF_INT,      /* cas*fso/pow(taudif*dt, 0.5)    temperature_component.cpp:539 */
  DI_QC1, DI_QC2, /* forcing-increment correction per unit dQ (temperature_component.cpp:471-475) */
  DI_INV_UC_CH4, DI_INV_TSOIL, DI_INV_TSTRAT, /* reciprocals of the CH4 constants */
  DI_LOG_M0, DI_SQRT_M0,                      /* log and sqrt of the preindustrial CH4 */
  DI_COUNT
};
/* parameters and derived constants live in ONE tiled array, [tile][PI_COUNT | DI_COUNT][128] */
#define PD_COUNT (PI_COUNT + DI_COUNT)
#define PD_OF(di) (PI_COUNT + (di)) /* a derived constant's field index in that array */

/* ---- recorded outputs ---- */
enum {
  OUT_CO2 = 0, OUT_TAS, OUT_RF_TOT, OUT_RF_CO2, OUT_HEATFLUX, OUT_OCEAN_C, OUT_HL_PH, OUT_ATMOS_C,
  OUT_SST, OUT_PERMAFROST_C, OUT_CH4, OUT_N2O, OUT_O3, OUT_LAND_TAS, OUT_VEG_C, OUT_DETRITUS_C,
  OUT_SOIL_C, OUT_THAWEDP_C, OUT_EARTH_C, OUT_NBP, OUT_OCEAN_UPTAKE, OUT_LL_PH, OUT_PCO2_HL,
  OUT_PCO2_LL, OUT_CARBON_HL, OUT_CARBON_LL, OUT_CARBON_IO, OUT_CARBON_DO, OUT_RF_CH4,
  OUT_RF_N2O, OUT_RH_CH4, OUT_NPP, OUT_RH, OUT_GMST, OUT_OCEAN_TAS, OUT_FLUX_MIXED,
  OUT_FLUX_INTERIOR, OUT_TIMESTEPS,
  /* per-stash quantities the year loop parks in the scratch rows X (below) when asked for */
  OUT_UPTAKE_HL, OUT_UPTAKE_LL, OUT_RH_DET, OUT_RH_SOIL,
  OUT_COUNT
};

/* ---- carbon tracking (fluxpool source maps, inst/include/fluxpool.hpp) ----
 * A tracked pool's unordered_map<source name, fraction> is a vector over the HX_NSRC possible
 * source names plus a key mask (a key with fraction 0 is not an absent key: it counts in the
 * 1/n split of a zero total and the tracking visitor prints it).  Pool / source order:
 * atmos_co2 earth_c veg_c detritus_c soil_c permafrost_c thawedp_c HL LL intermediate deep
 * (+ source "untracked").  Per member the maps live in a CTA-tiled array
 * T[tile][TS_COUNT * HX_NSRC][128] (+ masks K[tile][TS_COUNT][128]): the 11 pools, the ocean's
 * year-start copy of the atmosphere (ocean_component.hpp:78,106), the stash-start copies
 * that fluxes made by flux_from_fluxpool() carry (scratch), and the four CarbonAdditions maps. */
#define HX_NPOOL 11
#define HX_NSRC 12
enum {
  TS_ATMOS = 0, TS_EARTH, TS_VEG, TS_DET, TS_SOIL, TS_PERM, TS_THAWED, TS_HL, TS_LL, TS_IO, TS_DO,
  TS_ATM_CPOOL,
  TS_ATM0, TS_PERM0, TS_OA, /* scratch of one stash: live in the replay kernel's registers only */
  TS_ADD_HL, TS_ADD_LL, TS_ADD_IO, TS_ADD_DO,
  TS_COUNT
};
#define HX_SRC_UNTRACKED 11

/* Biomes (simpleNbox.cpp:201-236): with more than one, the land pools, their parameters and the
 * yearly land factors exist once per biome.  Per-biome arrays are tiled like P and S:
 * [tile][biome * COUNT + field][HX_TILE].  The per-member state keeps the across-biome sums
 * (SI_VEG .. SI_THAWED, SI_RH_CH4, SI_X_NPP, SI_X_RH), which is all the other components see. */
#define HX_MAX_BIOMES 4
enum { /* per-biome parameters, the order of the oracle's ho_biome */
  BP_VEG_C0 = 0, BP_DET_C0, BP_SOIL_C0, BP_PERMAFROST_C0, BP_NPP_FLUX0, BP_BETA, BP_Q10_RH,
  BP_WARMINGFACTOR, BP_F_NPPV, BP_F_NPPD, BP_F_LITTERD, BP_RH_CH4_FRAC, BP_PF_MU, BP_PF_SIGMA,
  BP_FPF_STATIC,
  BP_COUNT
};
enum { /* per-biome state */
  BF_VEG = 0, BF_DET, BF_SOIL, BF_PERMAFROST, BF_THAWED,
  BF_TEMPFERTS, /* tempferts_tv of last year (sticky soil Q10 factor) */
  BF_F_FROZEN,
  BF_X_CO2FERT, BF_X_TFD, BF_X_TFS, BF_X_FNEWTHAW, /* this year's slow parameters */
  BF_X_NPP, BF_X_RH,                               /* final_npp / final_rh of the last stash */
  BF_RH_CH4,
  /* the biome's fluxes of the current sub-step (constant between two stashes): computed once
   * with the sub-step constants, read again by the stash that ends the sub-step */
  BF_S_NPP, BF_S_RH_FDA, BF_S_RH_FSA, BF_S_RH_CO2, BF_S_RH_CH4,
  BF_COUNT
};
/* per-biome outputs "<biome>.<name>" (getData with a biome prefix, simpleNbox.cpp:533-697):
 * output id OUT_COUNT + biome * BO_COUNT + k */
enum { BO_VEG = 0, BO_DET, BO_SOIL, BO_PERMAFROST, BO_THAWED, BO_NPP, BO_RH, BO_COUNT };
#define HX_OUT_IDS (OUT_COUNT + HX_MAX_BIOMES * BO_COUNT)

/* ---- per-member N2O / halocarbon parameters and state (the GAS build of the run kernel) ----
 * By default the N2O concentration and the 26 halocarbon forcings are member-independent series
 * computed on the host per scenario (E-5).  When one of their parameters is given per member
 * (n2o_component.cpp:98-116: N0, UC_N2O, TN2O0; halocarbon_component.cpp:127-136: tau, rho, delta,
 * H0, molarMass of a gas) the 27 recurrences run per member on the device: parameters
 * GP[tile][GP_COUNT][128], state GF[tile][GF_COUNT][128], and the emissions come from a second
 * scenario table [scenario][row][HX_GAS_COLS] (N2O_emissions, N2O_natural_emissions, 26 x
 * <gas>_emissions). */
enum {
  GP_UC_N2O = 0, GP_TN2O0,
  GP_HALO0, /* gas g: GP_HALO0 + 5 g + {0 tau, 1 rho, 2 delta, 3 H0, 4 molarMass} */
  GP_COUNT = GP_HALO0 + 5 * HX_NHALO
};
enum {
  GF_N2O = 0,                      /* N2O concentration of the last year */
  GF_HA0 = 1,                      /* 26 halocarbon concentrations */
  GF_EXPFAC0 = GF_HA0 + HX_NHALO,  /* 26 x exp(-1 / tau), set-up */
  GF_RF0 = GF_EXPFAC0 + HX_NHALO,  /* 26 x this year's adjusted forcing rho Ha (1 + delta) */
  GF_COUNT = GF_RF0 + HX_NHALO
};
#define HX_GAS_COLS (2 + HX_NHALO)

/* ---- engine-wide constants handed to every kernel ---- */
struct HxConst {
  int32_t start_year, end_year, nrow; /* nrow = end - start + 1 */
  int32_t baseyear;
  int32_t max_spinup;
  uint32_t flags;
  /* carbon tracking: first tracked year (core.cpp:228-235; 9999 = off); recorded years are
   * tracking_date + k * track_every (k >= 0) and end_year (track_every = 0: end_year only) */
  int32_t tracking_date, track_every, track_nrec;
  /* salinity-only chemistry constants, computed on the host with the C library so they are
   * the doubles the reference computes (ocean_csys.cpp:225-287) */
  double S, sqrtS, S15, bor;
  /* geometry (ocean_component.cpp:207-303) */
  double vol_HL, vol_LL, vol_IO, vol_DO, As_HL, As_LL, U;
  double inv_vol_HL, inv_vol_LL; /* 1 / volume of the surface boxes (convertToDIC) */
  double spy_ocean;
  /* DOECLIM (temperature_component.hpp:77-98) */
  double powtoheat;
  /* odeint default_step_adjuster growth at the error floor: 0.9 * pow(pow(5,-5), -1/5) */
  double rk_grow_max;
  /* biomes: count (1 = the single "global" biome of the scalar parameters) and, for sum_map
   * (simpleNbox.cpp:428-438, a std::map walk), the biome indices sorted by name */
  int32_t n_biomes;
  int32_t biome_order[HX_MAX_BIOMES];
};

/* status words live next to the state */
struct HxStatus {
  int32_t *status;    /* [Mpad] */
  int32_t *fail_year; /* [Mpad] */
};

#endif
