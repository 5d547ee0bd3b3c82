/* include/hector_b200.h -- C ABI of libhector_b200.so, the B200 ensemble engine for Hector's
 * per-year coupled hot path (carbon-cycle ODE + ocean boxes/chemistry + forcing + DOECLIM),
 * batched over ensemble members.
 *
 * Plain C types, caller-allocated buffers, FP64 throughout.  Each entry point replaces a
 * piece of the reference's C++ surface for this path (paths relative to JGCRI/hector v3.5.0):
 *
 *   hx_create / hx_destroy     Core::mkcore / Core::delcore + Core::init   src/core.cpp:817-857, 90-196
 *   hx_set_scenario_series     Core::setData(component, var, (date,value)) src/core.cpp:219-268 as driven by
 *                              CSVTableReader::process                     src/csv_table_reader.cpp:115-196
 *   hx_set_param[_scalar]      Core::sendMessage(M_SETDATA, var, value)    src/core.cpp:752-772
 *                              (R: setvar, R/messages.R:107-140)
 *   hx_prepare                 Core::prepareToRun incl. run_spinup         src/core.cpp:302-420
 *   hx_run                     Core::run(runtodate)                        src/core.cpp:448-509
 *   hx_reset                   Core::reset(date < start)                   src/core.cpp:511-549
 *   hx_fetch                   Core::sendMessage(M_GETDATA, var, date)     src/core.cpp:716-778
 *                              (R: fetchvars, R/messages.R:46-88)
 *   hx_member_status           h_exception propagation                     inst/include/h_exception.hpp:26-102
 *
 * Errors never cross the ABI as exceptions: every call returns HX_OK or a negative code and
 * hx_last_error() describes it.  Members that the reference would abort with an h_exception
 * get a non-zero per-member status and NaN outputs from the failing year on; the rest of the
 * batch continues.
 */
#ifndef HECTOR_B200_H
#define HECTOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hx_engine *hx_handle;

/* return codes */
#define HX_OK 0
#define HX_ERR_ARG (-1)     /* bad argument / unknown name */
#define HX_ERR_STATE (-2)   /* call out of order (e.g. run before prepare) */
#define HX_ERR_CUDA (-3)    /* CUDA runtime error, see hx_last_error */
#define HX_ERR_UNSUPPORTED (-4)

/* per-member status words (0 = ok); the comment names the reference exception */
#define HX_MEMBER_OK 0
#define HX_MEMBER_NEGATIVE 1     /* "Flux and pool values may not be negative" fluxpool.hpp:100-102 */
#define HX_MEMBER_MASS 2         /* "Mass not conserved" simpleNbox-runtime.cpp:556-563 */
#define HX_MEMBER_RETRIES 3      /* "solver failure: t != tnew" carbon-cycle-solver.cpp:294 */
#define HX_MEMBER_NOROOT 4       /* newton_raphson_iterate lost its bracket / produced NaN */
#define HX_MEMBER_YEARFRACTION 5 /* "yearfraction out of bounds" ocean_component.cpp:665 */
#define HX_MEMBER_CO2SARF 6      /* forcing_component.cpp:353 */
#define HX_MEMBER_STEPPER 7      /* odeint: 500 failed step-size searches */
#define HX_MEMBER_SPINUP 8       /* spin-up did not converge within max_spinup (reference only logs) */
#define HX_MEMBER_NEEDS_EXACT 10  /* not a reference failure: an ODE attempt the reference abandons
                                    (and the engine skips) could have raised its negativity
                                    exception for this member; re-run with HX_FLAG_EXACT_ATTEMPTS */
#define HX_MEMBER_TRACKING 9     /* "fractions must be 0-1" / "pool_map must sum to ~1.0" fluxpool.hpp:105-112 */

/* hx_config.flags */
#define HX_FLAG_COLD_NEWTON 1u   /* start every [H+] solve from the Fujiwara bound like the
                                    reference (ocean_csys.cpp:141-153) instead of last root */
#define HX_FLAG_NO_SPINUP 2u     /* do_spinup = 0 */
#define HX_FLAG_EXACT_ATTEMPTS 4u /* execute the ODE attempts the reference abandons whenever one of
                                    their stage states could go negative (slower build of the run
                                    kernel; without it such a member stops with
                                    HX_MEMBER_NEEDS_EXACT).  Every combination of features has such a build. */
#define HX_FLAG_KEEP_ORDER 8u    /* keep members in caller order on the device (default: members of a
                                    scenario are re-ordered so that the members of a warp behave
                                    alike; outputs are in caller order either way).  Runs that
                                    record more than 20 variables keep the caller's order anyway:
                                    their output rows then store coalesced, which is worth more */

typedef struct {
  int32_t n_members;   /* ensemble members owned by this engine (this GPU's shard) */
  int32_t n_scenarios; /* distinct scenario tables */
  int32_t start_year;  /* [core] startDate, e.g. 1745 */
  int32_t end_year;    /* [core] endDate,   e.g. 2300 */
  int32_t device;      /* CUDA device ordinal */
  uint32_t flags;
} hx_config;

/* counters returned by hx_counters (sums over members of the last hx_run segment) */
#define HX_NCOUNTERS 8
#define HX_CNT_RHS_EVALS 0
#define HX_CNT_RK_STEPS 1
#define HX_CNT_RK_REJECTED 2
#define HX_CNT_STASHES 3
#define HX_CNT_NEWTON_ITERS 4
#define HX_CNT_NEWTON_CALLS 5
#define HX_CNT_FAILED_MEMBERS 6
#define HX_CNT_MEMBER_YEARS 7

int hx_create(const hx_config *cfg, hx_handle *out);
/* newcore(inifile): create an engine from n_inis Hector ini files (one scenario each; csv:
 * tables resolved like the reference does, src/ini_to_core_reader.cpp:134-167).  Scalar
 * parameters come from the first file, start/end dates must agree.  Replaces
 * INIToCoreReader::parse + Core::setData (src/ini_to_core_reader.cpp:74-180). */
int hx_create_from_ini(const char *const *ini_paths, int32_t n_inis, int32_t n_members,
                       int32_t device, uint32_t flags, hx_handle *out);
/* host-only access to the ini/csv reader (no GPU needed): run dates, the dense raw table
 * [end-start+1][44] in the series order listed at hx_set_scenario_series, scalar values by
 * engine parameter name ("S", "beta", "CF4.tau", ...) */
int hx_ini_read(const char *ini_path, int32_t *start_year, int32_t *end_year, double *table,
                int32_t table_rows);
int hx_ini_scalar(const char *ini_path, const char *name, double *out);
/* the one string input of an ini file: [core] run_name (Core::getRun_name, core.hpp:88) */
int hx_ini_string(const char *ini_path, const char *name, char *buf, int32_t cap);
int hx_destroy(hx_handle h); /* idempotent on NULL */
const char *hx_last_error(hx_handle h); /* h may be NULL: error of the last failed hx_create */

/* Use an existing CUDA stream (a cudaStream_t passed as void*) for all engine work; NULL
 * selects the engine's own stream. */
int hx_set_stream(hx_handle h, void *cuda_stream);

/* Raw scenario series, one value per integer year year0 .. year0+n-1.  The first call for a
 * series must cover start_year..end_year; later calls may overwrite any sub-range of years (R's
 * setvar(core, dates, var, values) editing a few years of an emission series; NaN entries are
 * skipped).  Names are the reference's input names: ffi_emissions, daccs_uptake,
 * luc_emissions, luc_uptake, CH4_emissions, CH4N, NOX_emissions, CO_emissions,
 * NMVOC_emissions, BC_emissions, OC_emissions, SO2_emissions, NH3_emissions, SV, RF_albedo,
 * RF_misc, N2O_emissions, N2O_natural_emissions, <gas>_emissions for the 26 halocarbons.
 *
 * User constraints use the same call with the reference's names CO2_constrain, NBP_constrain,
 * tas_constrain, RF_tot_constrain, CH4_constrain, N2O_constrain and <gas>_constrain
 * (simpleNbox-runtime.cpp:343-383, 567-603, 871-898, temperature_component.cpp:510-525, forcing_component.cpp:498-505, ch4_component.cpp:
 * 141-158, n2o_component.cpp:141-158, halocarbon_component.cpp:189-192): any sub-range of years,
 * NaN = no entry for that year.  Entries act where the reference's tseries would: CO2 / NBP /
 * CH4 / N2O / halocarbons in the years that have an entry; RF_tot in every year up to its last
 * entry and tas between its first and last entry, gaps interpolated linearly.
 *
 * Series may also be (re)set after hx_prepare; like the reference's setvar they take effect at
 * the next hx_reset / hx_run, which re-runs set-up and spin-up. */
int hx_set_scenario_series(hx_handle h, int32_t scenario_id, const char *name, int32_t year0,
                           int32_t n, const double *values);
/* whole table at once: values[n_years][n_names] row-major */
int hx_set_scenario_table(hx_handle h, int32_t scenario_id, int32_t n_names,
                          const char *const *names, int32_t year0, int32_t n_years,
                          const double *values);
/* scenario of every member (default: all members use scenario 0) */
int hx_set_member_scenario(hx_handle h, const int32_t *scenario_of_member, int32_t n);

/* Parameters by the reference's names (S, diff, qco2, beta, q10_rh, f_nppv, f_nppd, f_litterd,
 * npp_flux0, C0, veg_c, detritus_c, soil_c, permafrost_c, warmingfactor, rh_ch4_frac, pf_mu,
 * pf_sigma, fpf_static, tt, tu, twi, tid, preind_surface_c, preind_interdeep_c, eps_abs,
 * eps_rel, dt, eps_spinup, aero_scalar, vol_scalar, delta_co2, delta_ch4, delta_n2o, rho_bc,
 * rho_oc, rho_so2, rho_nh3, M0, Tsoil, Tstrat, UC_CH4, TOH0, CNOX, CCO, CNMVOC, CCH4, PO3, N0,
 * lo_warming_ratio per member or scalar; baseyear and max_spinup scalar only).  Defaults =
 * inst/input/hector_ssp245.ini.
 *
 * The N2O and halocarbon parameters -- N0, UC_N2O, TN2O0 (n2o_component.cpp:98-116) and
 * <gas>.tau / .rho / .delta / .H0 / .molarMass for the 26 halocarbons (halocarbon_component.cpp:
 * 127-136) -- are scalars by default, and their 27 series are then computed once per scenario on
 * the host.  Given per member (hx_set_param, before hx_prepare) they move the 27 recurrences
 * into the run kernel (its GAS builds): with CO2 / NBP / CH4 / RF_tot / tas constraints,
 * lo_warming_ratio, carbon tracking, biomes and HX_FLAG_EXACT_ATTEMPTS too; with an N2O or
 * halocarbon concentration constraint hx_prepare returns HX_ERR_UNSUPPORTED. */
int hx_set_param_scalar(hx_handle h, const char *name, double value);
int hx_set_param(hx_handle h, const char *name, const double *per_member, int32_t n);
/* one member's value (the other members keep theirs; a scalar parameter becomes per-member):
 * Core::sendMessage(M_SETDATA, var, value) addressed to one core of an ensemble */
int hx_set_param_member(hx_handle h, const char *name, int32_t member, double value);
/* same, from a DEVICE pointer (no host round trip) */
int hx_set_param_device(hx_handle h, const char *name, const double *per_member_dev, int32_t n);
int hx_get_param(hx_handle h, const char *name, double *per_member_out, int32_t n);

/* Variables recorded per year (subset of: CO2_concentration, global_tas, RF_tot, RF_CO2,
 * heatflux, ocean_c, HL_pH, atmos_co2, sst, permafrost_c, CH4_concentration,
 * N2O_concentration, O3_concentration, land_tas, veg_c, detritus_c, soil_c, thawedp_c, earth_c,
 * NBP, ocean_uptake, LL_pH, HL_PCO2, LL_PCO2, HL_ocean_c, LL_ocean_c, IO_ocean_c, DO_ocean_c,
 * RF_CH4, RF_N2O, rh_ch4, NPP, RH, gmst, ocean_tas, heatflux_mixed, heatflux_interior,
 * ocean_timesteps).  Default: CO2_concentration, global_tas.
 * Must be called before hx_prepare. */
int hx_select_outputs(hx_handle h, int32_t n, const char *const *names);

int hx_prepare(hx_handle h);
/* Core::run(runtodate) (core.cpp:448-509).  < 0: run to end_year; resumes where it left off.
 * After a parameter or input change (hx_set_param*, hx_set_scenario_series) the next run starts
 * over from start_year with a fresh set-up and spin-up, like R's setvar + reset + run; when the
 * change is made in the MIDDLE of a run (current date > start_year) hx_run and hx_run_stream
 * refuse with HX_ERR_STATE until hx_reset / hx_reset_date is called -- the reference would carry
 * the old state on with the new values, which needs a state history the engine does not keep. */
int hx_run(hx_handle h, double run_to_date);
int hx_reset(hx_handle h);                   /* back to the post-spin-up state at start_year */
/* Core::reset(resetdate) (core.cpp:511-549): date <= start_year is hx_reset; a date inside the
 * run restores the state of that year so that hx_run continues from it.  The engine keeps no
 * per-year state history (the reference keeps one tvector per state variable); it re-derives the
 * state by re-running to `date`, which is exact because runs are bit-reproducible.  After input
 * changes it stays exact -- and is allowed -- when every change made since the last run concerns
 * years after `date` only: the dated hx_set_scenario_series edits of R's setvar(core, dates, ...)
 * followed by reset(core, min(dates) - 1).  After a change that reaches back to `date` or
 * earlier (any parameter) it is refused with HX_ERR_UNSUPPORTED: reset to the start instead. */
int hx_reset_date(hx_handle h, double date);
int hx_synchronize(hx_handle h);
/* hx_run + fetch of every year of the segment in one call, with the device-to-host copies
 * overlapped with the computation.  segments > 1: ONE launch of the persistent run kernel; the
 * kernel raises a flag in mapped host memory when every tile has finished a 16-year slab, and
 * the calling thread queues that slab's rows on the copy engine while later slabs compute, so
 * only the last slab's copy is exposed (a failed member's NaNs are written by the kernel).
 * Carbon-tracking runs are cut into `segments` launches of halving length instead (their slabs
 * are separate launches already); segments = 1 runs, then copies.  outs[v] receives
 * names[v] YEAR-major: outs[v][(year - first_year) * n_members + member], first_year = the date
 * before the call + 1 -- the long format R's fetchvars returns (R/messages.R:46-88), and the
 * device layout, so no transpose stands between the kernel and the copy.  Use pinned host
 * memory for the copies to be asynchronous.  Needs members in API order on the device (a single
 * scenario, or members listed scenario by scenario); returns when everything has arrived. */
int hx_run_stream(hx_handle h, double run_to_date, int32_t n_vars, const char *const *names,
                  double *const *outs, int32_t segments);

/* out[member][date] (row-major, n_members x n_dates) on the host; dates before start_year or
 * beyond the current date are an error, as in the reference.  The START date itself is a valid
 * date for the recorded variables, as it is there (R's fetchvars keeps dates >= startdate,
 * R/messages.R:66): the pools as the spin-up left them, the preindustrial concentrations
 * (C0, M0, N0, PO3), zero for temperatures, heat fluxes, forcings, pH / pCO2 and the ocean
 * uptake, the spin-up's last NPP and RH; NBP has no entry there (HX_ERR_ARG; the reference
 * throws).  The derived outputs below start in start_year + 1.
 * An INPUT series is read back under its own name (ffi_emissions, luc_emissions, CH4_emissions,
 * <gas>_emissions, SV, RF_albedo ...; any year of the run) like the getData of the component
 * that owns it; a constraint series (CO2_constrain ...) answers NaN where it has no entry.
 * Besides the recorded variables, hx_fetch answers for the outputs that need no kernel -- it
 * derives them from the scenario series, per-member parameters and recorded outputs with
 * ForcingComponent::run's expressions (forcing_component.cpp:410-524): RF_BC, RF_OC, RF_SO2,
 * RF_NH3, RF_aci, RF_vol, RF_albedo, RF_misc, RF_O3_trop (needs O3_concentration recorded),
 * RF_H2O_strat (needs CH4_concentration recorded), and per halocarbon RF_<gas> (absolute),
 * Fadj<gas> (relative to the base year) and <gas>_concentration.
 * And for the rest of the reference's outputstream variables that are plain functions of
 * recorded ones (each needs those recorded): HL_sst, LL_sst (the box temperatures the year's
 * chemistry ran at, from sst), HL_DIC, LL_DIC, DIC, ML_ocean_c (from HL_ocean_c / LL_ocean_c),
 * pH, PCO2 (area-weighted surface means), TAU_OH (from CH4_concentration), f_frozen (from
 * land_tas and permafrost_c; with biomes "<biome>.f_frozen" from "<biome>.permafrost_c" and
 * f_frozen their mean weighted with the permafrost of the current date), HL_CO3, LL_CO3, CO3 and the calcite / aragonite
 * saturation states HL_OmegaCa, LL_OmegaCa, HL_OmegaAr, LL_OmegaAr (from the recorded pCO2, pH
 * and sst).  HL_ocean_uptake, LL_ocean_uptake, rh_det and rh_soil are RECORDED outputs: select them
 * with hx_select_outputs (they cost scratch rows of global memory traffic per stash, so they are
 * not part of the default set; with biomes rh_det and rh_soil are the sums over the biomes). */
int hx_fetch(hx_handle h, const char *name, const double *dates, int32_t n_dates, double *out);
/* device-resident view: pointer to the [year][member_stride] block of `name` (year index 0 =
 * start_year+1); valid until hx_destroy.  For NCCL gathers / zero-copy consumers. */
int hx_output_device(hx_handle h, const char *name, const double **dev_ptr,
                     int64_t *member_stride, int32_t *n_years);

/* Carbon tracking ([core] trackingDate, Core::getTrackingData: src/core.cpp:199-209, 228-235;
 * CSVFluxPoolVisitor: src/csv_tracking_visitor.cpp:80-137; the algebra: inst/include/fluxpool.hpp:
 * 197-257).  From year `tracking_date` on (startDate < tracking_date <= endDate) the source
 * fractions of the 11 tracked pools are carried per member and recorded for the years
 * tracking_date + k * record_every and for end_year (record_every = 0: end_year only).  Call
 * before hx_prepare; "trackingDate" via hx_set_param_scalar / the ini reader is the same switch
 * with record_every = 1.  Pool and source order: atmos_co2 earth_c veg_c detritus_c soil_c
 * permafrost_c thawedp_c HL LL intermediate deep (+ source 11 "untracked"). */
#define HX_TRACK_NPOOL 11
#define HX_TRACK_NSRC 12
int hx_set_tracking(hx_handle h, int32_t tracking_date, int32_t record_every);
/* Core::getTrackingDate (core.hpp:85): 9999 = tracking off */
int hx_tracking_date(hx_handle h);
/* frac[member][pool][source] (n_members x 11 x 12) and, if mask != NULL, mask[member][pool] whose
 * bit s says that source s is a key of the pool's map (the rows the reference's tracking CSV
 * prints; a key can carry fraction 0).  `date` must be a recorded year or the current date. */
int hx_fetch_tracking(hx_handle h, double date, double *frac, uint32_t *mask);
/* the recorded years, ascending; returns their count (0 when tracking is off) or a negative
 * error; at most `cap` are written */
int hx_tracking_years(hx_handle h, int32_t *years, int32_t cap);

/* Biomes (simpleNbox.cpp:201-236, biome-split pools simpleNbox-runtime.cpp:399-531).  Replaces the
 * single "global" biome by 2 .. HX_MAX_BIOMES named ones, in creation (biome_list) order; call
 * before hx_prepare.  Afterwards the land inputs exist per biome and are set through
 * hx_set_param_scalar / hx_set_param under the reference's names "<biome>.<name>", name in
 *   veg_c detritus_c soil_c permafrost_c npp_flux0 beta q10_rh f_nppv f_nppd f_litterd   (required)
 *   warmingfactor rh_ch4_frac pf_mu pf_sigma fpf_static                       (default like the reference)
 * (after hx_prepare they take effect at the next reset / run, like every parameter); the global
 * spellings are then refused ("cannot have both global and
 * biome-specific data").  The plain output names stay the across-biome totals the reference
 * reports for the global datum; "<biome>.<name>", name in veg_c detritus_c soil_c permafrost_c
 * thawedp_c NPP RH, selects and fetches one biome's own (hx_select_outputs after hx_set_biomes).
 * Not combined with carbon tracking (HX_ERR_UNSUPPORTED at hx_prepare).
 * n_biomes <= 1 with names == NULL returns to the global biome. */
#define HX_MAX_BIOMES 4
int hx_set_biomes(hx_handle h, int32_t n_biomes, const char *const *names);
int hx_biome_count(hx_handle h);
/* name of biome i in creation order ("global" when there is only the global one):
 * Core::getBiomeList (core.cpp:560-563) */
int hx_biome_name(hx_handle h, int32_t i, char *buf, int32_t cap);

/* Exchange of recorded outputs between the GPUs of one node WITHOUT kernels: every rank exports
 * its output block through CUDA IPC (64-byte handle; pass the handles around with any host-side
 * collective), opens its peers' blocks, and pulls finished year ranges with copy-engine
 * transfers over NVLink while its own run kernel keeps every SM (a collective kernel finds no
 * room next to the persistent run kernel).  All ranks must have the
 * same outputs, years and padded member count.  dst_dev is a device buffer
 * [n_peers][n_years][member_stride] (hx_output_device gives n_years and member_stride); rows
 * year_a .. year_b (inclusive, calendar years) of every peer land at their place in it.  The
 * caller orders things across ranks: pull a range only after every rank has finished it
 * (hx_event_record after the run segment, hx_event_synchronize, then a host barrier). */
int hx_ipc_export(hx_handle h, void *handle64, int64_t *bytes);
int hx_ipc_open(hx_handle h, int32_t n_peers, const void *handles, int32_t self_index);
int hx_ipc_pull(hx_handle h, const char *name, int32_t year_a, int32_t year_b, double *dst_dev);
int hx_ipc_wait(hx_handle h);  /* all pulls issued so far have landed */
int hx_ipc_close(hx_handle h);
/* ---- the job's one exchange, PUSHED while the run kernel computes (members
 * shard over the GPUs of a node; afterwards every GPU holds every member's recorded outputs).
 * No reference counterpart: the reference is single-process.  One process per GPU, all engines
 * with the same outputs, years and padded member count:
 *   hx_xchg_create   allocates this rank's gather block [n_peers][n_out][n_years][stride] (stride
 *                    and n_years as hx_output_device reports them) and returns its CUDA IPC handle;
 *   (the caller exchanges the 64-byte handles between the ranks -- MPI_Allgather, a torch.
 *    distributed all_gather_object, a file: any transport)
 *   hx_xchg_open     opens the peers' blocks (handles[n_peers][64], own slot ignored);
 *   hx_run_exchange  hx_run as ONE launch of the persistent kernel; each 16-year slab, once every
 *                    tile has finished it (flag in mapped host memory), is copied into slot
 *                    `self_index` of EVERY rank's block by the copy engines, over NVLink for the
 *                    peers, while the kernel computes the later slabs; returns when the kernel and
 *                    this rank's copies are complete.  A barrier across the ranks (the caller's:
 *                    MPI_Barrier, gloo ...) then makes every block complete; another one must
 *                    precede the next run, which overwrites the blocks.
 *   hx_xchg_block    device pointer of this rank's block, elements per rank slot.
 * Only the last slab's copy and one host barrier stay exposed; no collective kernel has to find
 * room next to the run kernel.  Untracked runs only. ---- */
int hx_xchg_create(hx_handle h, int32_t n_peers, int32_t self_index, void *handle64, int64_t *bytes);
int hx_xchg_open(hx_handle h, int32_t n_peers, const void *handles);
int hx_run_exchange(hx_handle h, double run_to_date);
int hx_xchg_block(hx_handle h, const double **dev_ptr, int64_t *elems_per_rank);
int hx_xchg_close(hx_handle h);

/* markers on the engine's stream (idx 0..15): record after a launch, wait for it on the host */
int hx_event_record(hx_handle h, int32_t idx);
int hx_event_synchronize(hx_handle h, int32_t idx);

int hx_member_status(hx_handle h, int32_t *status, int32_t *fail_year, int32_t n);
int hx_counters(hx_handle h, uint64_t *out, int32_t n);
double hx_current_date(hx_handle h);
/* device time of the last hx_run segment measured with CUDA events on the engine stream */
double hx_last_run_ms(hx_handle h);
/* post-spin-up snapshot of member m: atmos, veg, det, soil, permafrost, thawed, earth,
 * HL, LL, IO, DO, alk_HL, alk_LL, spinup_steps (14 doubles; alk valid after the first year) */
int hx_spinup_state(hx_handle h, int32_t member, double *out14);

/* ---- diagnostics: the denominators of the roofline bench.py reports ("FP64
 * peak is not in MEASURED_PEAKS.json; the builder must microbenchmark it").  No reference
 * counterpart.  hx_measure_fp64_peak: dependent-FMA chains on every SM, best of three launches,
 * TFLOP/s (2 flops per FMA) and FMAs per clock per SM at the nominal SM clock;
 * hx_measure_hbm_copy: device-to-device copy of 1 GiB, read + write GB/s. ---- */
int hx_measure_fp64_peak(int32_t device, double *tflops, double *fma_per_clk_per_sm);
int hx_measure_hbm_copy(int32_t device, double *gbs);
/* the run kernel's interleaved exp (which = 0) / exp10 (1) / log (2) next to the CUDA library's,
 * for n host values: the two output arrays must be bit-identical (hx_model.cuh, hx_exp_n) */
int hx_diag_transcendentals(int32_t device, int32_t which, const double *x, double *fast,
                            double *lib, int32_t n);

const char *hx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HECTOR_B200_H */
