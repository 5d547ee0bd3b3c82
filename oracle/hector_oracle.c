/* oracle/hector_oracle.c -- TEST INFRASTRUCTURE ONLY (see hector_oracle.h).
 *
 * Plain-C restatement of Hector v3.5.0's per-year coupled step for ONE member, written to
 * follow the reference's evaluation order statement by statement.  Citations are
 * file:line under /root/reference.  Third-party numerics (Boost, un-vendored; comments in
 * the reference mention 1.81) are restated from their published algorithms:
 *   - odeint controlled runge_kutta_dopri5 + integrate_adaptive  (carbon-cycle-solver.cpp:257-261)
 *   - math::tools::newton_raphson_iterate                        (ocean_csys.cpp:152-153)
 *   - math::tools::brent_find_minima                             (oceanbox.cpp:437-438)
 *   - math::lognormal cdf                                        (simpleNbox-runtime.cpp:1028)
 * Pinned by tests/test_oracle.py against the reference's golden file
 * tests/testthat/compdata/hector_comp.csv and against oracle/_ref (the unmodified reference).
 *
 * Compile with -ffp-contract=off: the reference is built without FMA contraction
 * (src/makefile.standalone:26-31).
 */
#include "hector_oracle.h"

#include <float.h>
#include <math.h>
#include <setjmp.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define PGC_TO_PPMVCO2 (1.0 / 2.13)            /* carbon-cycle-model.hpp:29 */
#define PPMVCO2_TO_PGC (1.0 / PGC_TO_PPMVCO2)  /* carbon-cycle-model.hpp:30 */
#define MAX_CARBON_MODEL_RETRIES 8             /* carbon-cycle-solver.hpp:23 */
#define CARBON_CYCLE_RETRY 1234                /* carbon-cycle-model.hpp:34 */
#define MB_EPSILON 0.001                       /* simpleNbox.hpp:37 */
#define OCEAN_MAX_TIMESTEP 1.0                 /* ocean_component.hpp:25-31 */
#define OCEAN_MIN_TIMESTEP 0.3
#define OCEAN_TSR_FACTOR 0.5
#define OCEAN_TSR_TIMEOUT 20
#define OCEAN_TSR_TRIGGER1 0.1
#define MEAN_TOS_TEMP 18                       /* oceanbox.hpp:35 */

enum { C_ATMOS = 0, C_VEG, C_DET, C_SOIL, C_PERMAFROST, C_THAWEDP, C_OCEAN, C_EARTH, NC };
enum { HL = 0, LL = 1, IO = 2, DO = 3 };

/* ---------------------------------------------------------------------------------- */
/* ocean carbonate chemistry: oceancsys (ocean_csys.hpp, ocean_csys.cpp)               */
typedef struct {
  double S, alk, As, U, volumeofbox;
  double K0, Tr, PCO2o, pH, CO3, TCO2o, HCO3, OmegaCa, OmegaAr, Kh, Kw;
} csys_t;

/* Carbon tracking (fluxpool.hpp): a fluxpool's unordered_map<source name, fraction> restated as
 * a fixed vector over the 12 possible source names plus a presence mask (a key that is in
 * the map with fraction 0 is not the same as an absent key: it counts in the 1/n split of a
 * zero total and it is printed by the tracking visitor).  Source / pool order:
 * atmos_co2 earth_c veg_c detritus_c soil_c permafrost_c thawedp_c HL LL intermediate deep
 * untracked. */
typedef struct {
  double f[HO_NSRC];
  unsigned mask;
} tmap_t;
enum { TP_ATMOS = 0, TP_EARTH, TP_VEG, TP_DET, TP_SOIL, TP_PERMAFROST, TP_THAWEDP, TP_BOX0,
       TP_UNTRACKED = 11 };

/* one ocean box (oceanbox.hpp, oceanbox.cpp) */
typedef struct {
  tmap_t cmap, addmap, aomap, oamap; /* source maps of carbon, CarbonAdditions, ao_flux, oa_flux */
  int tracking, ao_tracking;
  double carbon, additions, subtractions;
  double Tbox, deltaT, atmosphere_flux, preindustrial_flux, ao_flux, oa_flux;
  int surfacebox, active_chemistry;
  int nconn, conn_to[3];
  double conn_k[3];
  double CO2_conc;
  csys_t chem;
} box_t;

/* one biome's pools, slow parameters and recorded fluxes (simpleNbox.hpp:236-290) */
typedef struct {
  ho_biome par;
  double veg_c, detritus_c, soil_c, permafrost_c, thawed_permafrost_c;
  double co2fert, tempfertd, tempferts, f_frozen, f_new_thaw;
  double tempferts_last_year; /* tempferts_tv[t] */
  double RH_ch4, final_npp, final_rh;
} bio_t;

typedef struct {
  const ho_params *p;
  const double *raw; /* [nrow][HO_NRAW] */
  int nrow;
  ho_counters cnt;
  jmp_buf fail;      /* stands in for h_exception propagation */
  int status;

  /* core */
  int in_spinup;

  /* simpleNbox state (simpleNbox.hpp) */
  double atmos_c, earth_c;
  int nb;                    /* biome_list.size() */
  bio_t bio[HO_MAX_BIOMES];  /* biome_list order */
  int border[HO_MAX_BIOMES]; /* std::map (name) order, for sum_map */
  double masstot, cum_luc_va, end_of_spinup_vegc, cumulative_pf_ch4, npp_luc_adjust;
  int have_tempferts_last;
  double current_luc_e, current_luc_u, current_ffi_e, current_daccs_u;
  double nbp;
  double snbox_ODEstartdate;
  double *Tland_record; /* [nrow], index = year - start; first key is start+1 */
  int has_been_run_before;

  /* ocean component state (ocean_component.hpp) */
  box_t box[4];
  double max_timestep, lastflux_annualized, annualflux_sum, annualflux_sumHL, annualflux_sumLL;
  int reduced_timestep_timeout, timesteps;
  double ocean_ODEstartdate, SST, ocean_CO2_conc;
  int ocean_in_spinup;

  /* solver (carbon-cycle-solver.hpp:86-98) */
  double c[NC], t, dt;

  /* user constraints (NULL = none) and the preindustrial values they may overwrite */
  const ho_constraints *cn;
  double M0_ch4, N0_n2o; /* CH4Component::M0 / N2OComponent::N0 after prepareToRun */

  /* carbon tracking: maps of atmos_c, earth_c and the five land pools (TP_* order), and the
   * ocean's copy of the atmosphere (set_atmosphere_sources, ocean_component.hpp:78,106) */
  int tracking_date, tracking;
  tmap_t tm[7];
  tmap_t atmosphere_cpool;
  int atmosphere_cpool_tracking;

  /* gas components */
  double *CH4, *O3, *N2O, *halo_rf; /* per-row series */
  double tau_oh;

  /* forcing */
  double base_tot, base_co2, base_ch4, base_n2o;
  double *rf_tot, *rf_co2, *rf_ch4, *rf_n2o; /* relative forcings, per row */
  double *co2_ts;                            /* atmos_c_ts in Pg C, per row */

  /* DOECLIM (temperature_component.hpp) */
  int ns;
  double *Ker, *temp, *temp_landair, *temp_sst, *heatflux_mixed, *heatflux_interior, *heat_mixed,
      *heat_interior, *forcing;
  double A[4], B[4], Cc[4], IB[4];
  double taucfl, taukls, taucfs, tauksl, taudif, powtoheat;
  double tas, tas_land, sst, heatflux, tas_ocean;
} member_t;

static void fail_member(member_t *m, int code) {
  m->status = code;
  longjmp(m->fail, 1);
}

/* fluxpool construction check: fluxpool.hpp:100-102, 121-123 (NaN passes, as in C++) */
static inline double FP(member_t *m, double v) {
  if (v < 0) fail_member(m, HO_ERR_NEGATIVE);
  return v;
}

/* fluxpool::set(v, u, track, name): ctmap[name] = 1.0 -- other keys are NOT erased
 * (fluxpool.hpp:118-127) */
static void tm_set_self(tmap_t *t, int self) {
  t->f[self] = 1.0;
  t->mask |= 1u << self;
}
static void tm_init(tmap_t *t, int self) {
  memset(t, 0, sizeof *t);
  tm_set_self(t, self);
}
/* private constructor checks (fluxpool.hpp:93-113): every fraction in [0, 1], sum - 1 < 1e-6 */
static void tm_check(member_t *m, const tmap_t *t) {
  double frac = 0.0;
  for (int s = 0; s < HO_NSRC; ++s)
    if (t->mask >> s & 1u) {
      if (!(t->f[s] >= 0 && t->f[s] <= 1)) fail_member(m, HO_ERR_TRACKING);
      frac += t->f[s];
    }
  if (!(frac - 1.0 < 1e-6)) fail_member(m, HO_ERR_TRACKING);
}
/* operator+(fluxpool, fluxpool) with tracking on (fluxpool.hpp:197-257): the map of
 * (a, A) + (b, B) is written to A.  Per source in the union of the key sets:
 * (a fa + b fb) / (a + b), or 1/n for every key when the new total is zero. */
static void tm_add(member_t *m, double a, tmap_t *A, double b, const tmap_t *B) {
  const double new_total = a + b;
  const unsigned un = A->mask | B->mask;
  int n = 0;
  for (int s = 0; s < HO_NSRC; ++s) n += (int)(un >> s & 1u);
  for (int s = 0; s < HO_NSRC; ++s)
    if (un >> s & 1u) {
      const double pool = a * A->f[s] + b * B->f[s]; /* get_fraction() is 0 for absent keys */
      A->f[s] = new_total ? pool / new_total : 1.0 / n;
    }
  A->mask = un;
  tm_check(m, A);
}

/* ---------------------------------------------------------------------------------- */
/* Boost newton_raphson_iterate (>= 1.7x), f = (poly, poly')                            */
static void poly_eval(const double a[6], double x, double *f0, double *f1) {
  /* polynomial::evaluate = Horner from the top (ocean_csys.cpp:104-118) */
  double d[5];
  for (int i = 1; i < 6; ++i) d[i - 1] = a[i] * (double)i;
  double s = a[5];
  for (int i = 4; i >= 0; --i) { s *= x; s += a[i]; }
  double sd = d[4];
  for (int i = 3; i >= 0; --i) { sd *= x; sd += d[i]; }
  *f0 = s;
  *f1 = sd;
}

static double sgn(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

static double newton_raphson_iterate(member_t *m, const double a[6], double guess, double min,
                                     double max, int digits, int *iters) {
  double f0 = 0, f1, last_f0 = 0;
  double result = guess;
  double factor = ldexp(1.0, 1 - digits);
  double delta = DBL_MAX, delta1 = DBL_MAX, delta2 = DBL_MAX;
  double max_range_f = 0, min_range_f = 0;
  int n = 0;
  do {
    last_f0 = f0;
    delta2 = delta1;
    delta1 = delta;
    poly_eval(a, result, &f0, &f1);
    ++n;
    if (0 == f0) break;
    if (f1 == 0) {
      /* handle_zero_derivative (never reached in practice) */
      double g0, g1;
      if (last_f0 == 0) {
        guess = (result == min) ? max : min;
        poly_eval(a, guess, &g0, &g1);
        last_f0 = g0;
        delta = guess - result;
      }
      if (sgn(last_f0) * sgn(f0) < 0) {
        delta = (delta < 0) ? (result - min) / 2 : (result - max) / 2;
      } else {
        delta = (delta < 0) ? (result - max) / 2 : (result - min) / 2;
      }
    } else {
      delta = f0 / f1;
    }
    if (fabs(delta * 2) > fabs(delta2)) {
      double shift = (delta > 0) ? (result - min) / 2 : (result - max) / 2;
      if ((result != 0) && (fabs(shift) > fabs(result))) {
        delta = sgn(delta) * fabs(result) * 1.1f;
      } else {
        delta = shift;
      }
      delta1 = 3 * delta;
      delta2 = 3 * delta;
    }
    guess = result;
    result -= delta;
    if (result <= min) {
      delta = 0.5F * (guess - min);
      result = guess - delta;
      if ((result == min) || (result == max)) break;
    } else if (result >= max) {
      delta = 0.5F * (guess - max);
      result = guess - delta;
      if ((result == min) || (result == max)) break;
    }
    if (delta > 0) {
      max = guess;
      max_range_f = f0;
    } else {
      min = guess;
      min_range_f = f0;
    }
    if (max_range_f * min_range_f > 0) {
      if (m) fail_member(m, HO_ERR_NOROOT);
      break;
    }
  } while (fabs(result * factor) < fabs(delta));
  if (iters) *iters = n;
  if (m) {
    m->cnt.newton_iterations += (uint64_t)n;
    m->cnt.newton_calls += 1;
  }
  return result;
}

/* find_largest_root: ocean_csys.cpp:134-156 */
static double find_largest_root(member_t *m, double a[6], int *iters) {
  const int degree = 5;
  double max = pow(fabs(a[0] / (2.0 * a[degree])), 1.0 / degree);
  for (int i = 1; i < degree; ++i) {
    double c = pow(fabs(a[i] / a[degree]), 1.0 / (double)(degree - i));
    max = max > c ? max : c; /* std::max(max, c) */
  }
  max *= 2.0;
  const int digits = 53;
  int get_digits = (int)(digits * 0.6);
  return newton_raphson_iterate(m, a, max - 0.001, 0.0, max, get_digits, iters);
}

/* convertToDIC: ocean_csys.cpp:403-408 (returns umol/kg) */
static double convertToDIC(const csys_t *c, double carbon) {
  const double dic =
      ((carbon * 1e15) * (1.0 / 12.01) * (1.0 / 1027.0) * (1.0 / c->volumeofbox));
  return dic * 1e6;
}

/* ocean_csys_run: ocean_csys.cpp:166-366 */
static double ocean_csys_run(member_t *m, csys_t *c, double tbox, double carbon, int *iters) {
  double tmp, tmp1, tmp2, tmp3;
  const double S = c->S, alk = c->alk;
  const double dic = convertToDIC(c, carbon) / 1e6;
  const double Tc = tbox;
  const double Tk = Tc + 273.15;

  tmp1 = -58.0931 + 90.5069 * (100 / Tk) + 22.2940 * log(Tk / 100);
  tmp2 = S * (0.027766 - 0.025888 * (Tk / 100) + 0.0050578 * ((Tk / 100) * (Tk / 100)));
  const double lnK0 = tmp1 + tmp2;
  c->K0 = exp(lnK0);

  const double Sc = 2073.1 - (125.62 * Tc) + (3.6276 * Tc * Tc) - (0.043219 * Tc * Tc * Tc);

  tmp1 = -13847.26 / Tk + 148.96502 - 23.6521 * log(Tk);
  tmp2 = +(118.67 / Tk - 5.977 + 1.0495 * log(Tk)) * sqrt(S) - 0.01615 * S;
  const double lnKw = tmp1 + tmp2;
  c->Kw = exp(lnKw);

  tmp = 9345.17 / Tk - 60.2409 + 23.3585 * log(Tk / 100);
  const double nKhwe74 = tmp + S * (0.023517 - 0.00023656 * Tk + 0.0047036e-4 * Tk * Tk);
  c->Kh = exp(nKhwe74);

  const double pK1mehr =
      3633.86 / Tk - 61.2172 + 9.6777 * log(Tk) - 0.011555 * S + 0.0001152 * S * S;
  const double K1 = pow(10, -pK1mehr);

  const double pK2mehr =
      471.78 / Tk + 25.9290 - 3.16967 * log(Tk) - 0.01781 * S + 0.0001122 * S * S;
  const double K2 = pow(10.0, -pK2mehr);

  tmp1 = (-8966.90 - 2890.53 * sqrt(S) - 77.942 * S + 1.728 * pow(S, (3.0 / 2.0)) -
          0.0996 * S * S) /
         Tk;
  tmp2 = +148.0248 + 137.1942 * sqrt(S) + 1.62142 * S;
  tmp3 = +(-24.4344 - 25.085 * sqrt(S) - 0.2474 * S) * log(Tk) + 0.053105 * sqrt(S) * Tk;
  const double lnKb = tmp1 + tmp2 + tmp3;
  const double Kb = exp(lnKb);

  tmp1 = -171.9065 - 0.077993 * Tk + 2839.319 / Tk + 71.595 * log10(Tk);
  tmp2 = +(-0.77712 + 0.0028426 * Tk + 178.34 / Tk) * sqrt(S);
  tmp3 = -0.07711 * S + 0.0041249 * pow(S, 1.5);
  const double log10Kspc = tmp1 + tmp2 + tmp3;
  const double Kspc = pow(10.0, log10Kspc);

  tmp1 = -171.945 - 0.077993 * Tk + 2903.293 / Tk + 71.595 * log10(Tk);
  tmp2 = +(-0.068393 + 0.0017276 * Tk + 88.135 / Tk) * sqrt(S);
  tmp3 = -0.10018 * S + 0.0059415 * pow(S, 1.5);
  const double log10Kspa = tmp1 + tmp2 + tmp3;
  const double Kspa = pow(10.0, log10Kspa);

  const double bor = 1 * (416.0 * (S / 35.0)) * 1.e-6;

  const double Kb_val = Kb, K1_val = K1, K2_val = K2, Kw_val = c->Kw;
  double a[6];
  const double p5 = -1.0;
  const double p4 = -alk - Kb_val - K1_val;
  const double p3 = dic * K1_val - alk * (Kb_val + K1_val) + Kb_val * bor + Kw_val -
                    Kb_val * K1_val - K1_val * K2_val;
  tmp = dic * (Kb_val * K1_val + 2.0 * K1_val * K2_val) -
        alk * (Kb_val * K1_val + K1_val * K2_val) + Kb_val * bor * K1_val;
  const double p2 = tmp + (Kw_val * Kb_val + Kw_val * K1_val - Kb_val * K1_val * K2_val);
  tmp = 2.0 * dic * Kb_val * K1_val * K2_val - alk * Kb_val * K1_val * K2_val +
        Kb_val * bor * K1_val * K2_val;
  const double p1 = tmp + (Kw_val * Kb_val * K1_val + Kw_val * K1_val * K2_val);
  const double p0 = Kw_val * Kb_val * K1_val * K2_val;
  a[0] = p0; a[1] = p1; a[2] = p2; a[3] = p3; a[4] = p4; a[5] = p5;

  const double h = find_largest_root(m, a, iters);

  const double co2st = dic / (1.0 + K1_val / h + K1_val * K2_val / h / h);
  const double hco3 = dic / (1.0 + h / K1_val + K2_val / h);
  const double co3 = dic / (1.0 + h / K2_val + h * h / K1_val / K2_val);
  const double million = 1e6;
  c->TCO2o = co2st * million;
  c->HCO3 = hco3 * million;
  c->CO3 = co3 * million;
  c->PCO2o = co2st * million / c->Kh;
  c->pH = -log10(h);
  c->Tr = (0.585 * c->K0 * pow(Sc, -0.5) * c->U * c->U);
  const double calcium = 0.02128 / 40.087 * (S / 1.80655);
  c->OmegaCa = ((co3 * calcium) / Kspc);
  c->OmegaAr = ((co3 * calcium) / Kspa);
  return h;
}

/* ocean_csys.cpp:375-396 */
static double calc_annual_surface_flux(const csys_t *c, double CO2_conc, double cpoolscale) {
  double monthly = ((CO2_conc - c->PCO2o * cpoolscale) * c->Tr);
  return (monthly * c->As * 12.0) / 1e15;
}

double ho_csys(double Tbox, double carbon_pgc, double alk, double volume, double S, double U,
               double *out8, int *iters) {
  csys_t c;
  memset(&c, 0, sizeof c);
  c.S = S; c.alk = alk; c.U = U; c.volumeofbox = volume; c.As = 1.0;
  double h = ocean_csys_run(NULL, &c, Tbox, carbon_pgc, iters);
  if (out8) {
    out8[0] = c.pH; out8[1] = c.PCO2o; out8[2] = c.Tr; out8[3] = c.K0;
    out8[4] = c.CO3; out8[5] = c.TCO2o; out8[6] = c.HCO3; out8[7] = c.OmegaCa;
  }
  return h;
}

/* ---------------------------------------------------------------------------------- */
/* oceanbox                                                                            */
static void box_separate_surface_fluxes(member_t *m, box_t *b) { /* oceanbox.cpp:262-271 */
  /* ao_flux = atmosphere_pool.flux_from_unitval(..), oa_flux = carbon.flux_from_unitval(..):
   * the fluxes inherit the source map (and tracking flag) of the pool they leave */
  b->aomap = m->atmosphere_cpool;
  b->ao_tracking = m->atmosphere_cpool_tracking;
  b->oamap = b->cmap;
  if (b->atmosphere_flux > 0) {
    b->ao_flux = FP(m, b->atmosphere_flux);
    b->oa_flux = FP(m, 0.0);
  } else {
    b->ao_flux = FP(m, 0.0);
    b->oa_flux = FP(m, -b->atmosphere_flux);
  }
}

/* oceanbox.cpp:203-260 */
static void box_compute_fluxes(member_t *m, int ib, double current_Ca, double yf, int do_circ) {
  box_t *b = &m->box[ib];
  b->CO2_conc = current_Ca;
  if (b->active_chemistry) {
    ocean_csys_run(m, &b->chem, b->Tbox, b->carbon, NULL);
    b->atmosphere_flux = calc_annual_surface_flux(&b->chem, b->CO2_conc, 1.0);
  } else {
    b->atmosphere_flux = b->surfacebox ? b->preindustrial_flux : 0.0;
  }
  b->atmosphere_flux = b->atmosphere_flux * yf;
  box_separate_surface_fluxes(m, b);
  if (do_circ) {
    for (int i = 0; i < b->nconn; ++i) {
      double closs = FP(m, FP(m, b->carbon * b->conn_k[i]) * yf);
      box_t *dst = &m->box[b->conn_to[i]];
      if (b->tracking) tm_add(m, dst->additions, &dst->addmap, closs, &b->cmap);
      dst->additions = FP(m, dst->additions + closs);   /* add_carbon, oceanbox.cpp:85-90 */
      b->subtractions = FP(m, b->subtractions + closs);
    }
  }
}

/* oceanbox.cpp:297-303 */
static void box_update_state(member_t *m, int ib) {
  box_t *b = &m->box[ib];
  if (b->tracking) tm_add(m, b->carbon, &b->cmap, b->additions, &b->addmap);
  double v = FP(m, b->carbon + b->additions);
  if (b->tracking) {
    tm_add(m, v, &b->cmap, b->ao_flux, &b->aomap);
  }
  if (b->tracking != b->ao_tracking) fail_member(m, HO_ERR_TRACKING); /* "tracking mismatch" */
  v = FP(m, v + b->ao_flux);
  v = FP(m, v - b->oa_flux);
  v = FP(m, v - b->subtractions);
  b->carbon = v;
  b->additions = 0.0;
  b->subtractions = 0.0;
  tm_set_self(&b->addmap, TP_BOX0 + ib); /* CarbonAdditions.set(0.0, U_PGC, tracking, Name) */
}

/* oceanbox.cpp:309-323 */
static void box_new_year(box_t *b, double SST) {
  b->Tbox = SST + (double)MEAN_TOS_TEMP + b->deltaT; /* compute_tabsC, oceanbox.cpp:97-99 */
  if (b->surfacebox) b->atmosphere_flux = 0.0;
}

/* fmin: oceanbox.cpp:335-349 */
static double box_fmin(member_t *m, box_t *b, double alk, double f_target) {
  b->chem.alk = alk;
  ocean_csys_run(m, &b->chem, b->Tbox, b->carbon, NULL);
  return fabs(calc_annual_surface_flux(&b->chem, b->CO2_conc, 1.0) - f_target);
}

/* Boost brent_find_minima (bits clamped to 53/2) driving fmin: oceanbox.cpp:382-445 */
static void box_chem_equilibrate(member_t *m, box_t *b, double current_Ca) {
  b->CO2_conc = current_Ca;
  double alk_min = 2100e-6, alk_max = 2750e-6;
  double f_target = b->preindustrial_flux;
  /* best-guess scan (its only lasting effect is chemistry calls; Brent overwrites alk) */
  double min_diff = 1e6;
  double min_point = (alk_min + alk_max) / 2.0;
  for (double alk1 = alk_min; alk1 <= alk_max; alk1 += (alk_max - alk_min) / 20) {
    double diff = box_fmin(m, b, alk1, f_target);
    if (diff < min_diff) { min_diff = diff; min_point = alk1; }
  }
  (void)min_point;

  int bits = (int)(53 * 0.6);
  bits = (53 / 2) < bits ? (53 / 2) : bits;
  double tolerance = ldexp(1.0, 1 - bits);
  double min = alk_min, max = alk_max;
  double x, w, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2;
  const double golden = 0.3819660f;
  x = w = v = max;
  fw = fv = fx = box_fmin(m, b, x, f_target);
  delta2 = delta = 0;
  for (;;) {
    mid = (min + max) / 2;
    fract1 = tolerance * fabs(x) + tolerance / 4;
    fract2 = 2 * fract1;
    if (fabs(x - mid) <= (fract2 - (max - min) / 2)) break;
    if (fabs(delta2) > fract1) {
      double r = (x - w) * (fx - fv);
      double q = (x - v) * (fx - fw);
      double pp = (x - v) * q - (x - w) * r;
      q = 2 * (q - r);
      if (q > 0) pp = -pp;
      q = fabs(q);
      double td = delta2;
      delta2 = delta;
      if ((fabs(pp) >= fabs(q * td / 2)) || (pp <= q * (min - x)) || (pp >= q * (max - x))) {
        delta2 = (x >= mid) ? min - x : max - x;
        delta = golden * delta2;
      } else {
        delta = pp / q;
        u = x + delta;
        if (((u - min) < fract2) || ((max - u) < fract2))
          delta = (mid - x) < 0 ? -fabs(fract1) : fabs(fract1);
      }
    } else {
      delta2 = (x >= mid) ? min - x : max - x;
      delta = golden * delta2;
    }
    u = (fabs(delta) >= fract1) ? (x + delta)
                                : (delta > 0 ? (x + fabs(fract1)) : (x - fabs(fract1)));
    fu = box_fmin(m, b, u, f_target);
    if (fu <= fx) {
      if (u >= x) min = x; else max = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) min = u; else max = u;
      if ((fu <= fw) || (w == x)) {
        v = w; w = u;
        fv = fw; fw = fu;
      } else if ((fu <= fv) || (v == x) || (v == w)) {
        v = u;
        fv = fu;
      }
    }
  }
  /* the reference discards (x, fx): alk stays at the last point fmin saw (oceanbox.cpp:338) */
}

/* ---------------------------------------------------------------------------------- */
/* OceanComponent                                                                      */
static void ocean_prepareToRun(member_t *m) { /* ocean_component.cpp:202-319 */
  const ho_params *p = m->p;
  const double spy = 60 * 60 * 24 * 365.25;
  const double part_high = 0.15, part_low = 1 - part_high; /* ocean_component.hpp:89-90 */
  const double thick_LL = 100, thick_HL = 100;
  const double thick_inter = 1000 - thick_LL;
  const double thick_deep = 3777 - thick_inter - thick_LL;
  const double ocean_area = 3.6e14;
  const double LL_volume = ocean_area * part_low * thick_LL;
  const double HL_volume = ocean_area * part_high * thick_HL;
  const double I_volume = ocean_area * thick_inter;
  const double D_volume = ocean_area * thick_deep;
  const double LL_vol_frac = LL_volume / (LL_volume + HL_volume);
  const double HL_vol_frac = 1 - LL_vol_frac;
  const double I_vol_frac = I_volume / (I_volume + D_volume);
  const double D_vol_frac = 1 - I_vol_frac;
  const double LL_preind_C = LL_vol_frac * p->preind_C_surface;
  const double HL_preind_C = HL_vol_frac * p->preind_C_surface;
  const double I_preind_C = I_vol_frac * p->preind_C_ID;
  const double D_preind_C = D_vol_frac * p->preind_C_ID;

  memset(m->box, 0, sizeof m->box);
  for (int i = 0; i < 4; ++i) {
    m->box[i].Tbox = -999;
    tm_init(&m->box[i].cmap, TP_BOX0 + i); /* initbox: carbon.set(boxc, U_PGC, false, name) */
    tm_init(&m->box[i].addmap, TP_BOX0 + i);
    tm_init(&m->box[i].aomap, TP_BOX0 + i);
    tm_init(&m->box[i].oamap, TP_BOX0 + i);
  }
  m->box[HL].carbon = FP(m, HL_preind_C);
  m->box[HL].surfacebox = 1;
  m->box[HL].preindustrial_flux = 1.000;
  m->box[HL].active_chemistry = p->spinup_chem;
  m->box[LL].carbon = FP(m, LL_preind_C);
  m->box[LL].surfacebox = 1;
  m->box[LL].preindustrial_flux = -1.000;
  m->box[LL].active_chemistry = p->spinup_chem;
  m->box[IO].carbon = FP(m, I_preind_C);
  m->box[DO].carbon = FP(m, D_preind_C);

  double LL_HL = (p->tt * spy) / LL_volume;
  double HL_DO = ((p->tt + p->tu) * spy) / HL_volume;
  double DO_IO = ((p->tt + p->tu) * spy) / D_volume;
  double IO_HL = (p->tu * spy) / I_volume;
  double IO_LL = (p->tt * spy) / I_volume;
  double IO_LLex = (p->twi * spy) / I_volume;
  double LL_IOex = (p->twi * spy) / LL_volume;
  double DO_IOex = (p->tid * spy) / D_volume;
  double IO_DOex = (p->tid * spy) / I_volume;

  /* make_connection order: ocean_component.cpp:278-284 */
  box_t *b;
  b = &m->box[LL]; b->nconn = 2; b->conn_to[0] = HL; b->conn_k[0] = LL_HL;
  b->conn_to[1] = IO; b->conn_k[1] = LL_IOex;
  b = &m->box[HL]; b->nconn = 1; b->conn_to[0] = DO; b->conn_k[0] = HL_DO;
  b = &m->box[IO]; b->nconn = 3; b->conn_to[0] = LL; b->conn_k[0] = IO_LL + IO_LLex;
  b->conn_to[1] = HL; b->conn_k[1] = IO_HL;
  b->conn_to[2] = DO; b->conn_k[2] = IO_DOex;
  b = &m->box[DO]; b->nconn = 1; b->conn_to[0] = IO; b->conn_k[0] = DO_IO + DO_IOex;

  m->box[HL].deltaT = -16.4;
  m->box[HL].chem.S = 34.5;
  m->box[HL].chem.volumeofbox = HL_volume;
  m->box[HL].chem.As = ocean_area * part_high;
  m->box[HL].chem.U = 6.7;
  m->box[LL].deltaT = 2.9;
  m->box[LL].chem.S = 34.5;
  m->box[LL].chem.volumeofbox = LL_volume;
  m->box[LL].chem.As = ocean_area * part_low;
  m->box[LL].chem.U = 6.7;

  m->annualflux_sum = m->annualflux_sumHL = m->annualflux_sumLL = 0.0;
  m->SST = 0.0;
  m->lastflux_annualized = 0.0;
  m->max_timestep = OCEAN_MAX_TIMESTEP; /* init(): ocean_component.cpp:76-77 */
  m->reduced_timestep_timeout = 0;
}

static double ocean_totalcpool(member_t *m) { /* ocean_component.cpp:325-328 */
  double v = FP(m, m->box[DO].carbon + m->box[IO].carbon);
  v = FP(m, v + m->box[LL].carbon);
  return FP(m, v + m->box[HL].carbon);
}

/* ocean_component.cpp:337-352 */
static double ocean_annual_totalcflux(member_t *m, double CO2_conc, double cpoolscale) {
  if (m->ocean_in_spinup && !m->p->spinup_chem)
    return m->box[HL].preindustrial_flux + m->box[LL].preindustrial_flux;
  return calc_annual_surface_flux(&m->box[HL].chem, CO2_conc, cpoolscale) +
         calc_annual_surface_flux(&m->box[LL].chem, CO2_conc, cpoolscale);
}

/* ocean_component.cpp:356-407; co2_conc = D_CO2_CONC(runToDate), sst = current D_SST */
static void ocean_run(member_t *m, double runToDate, double co2_conc, double sst) {
  /* tracking start: ocean_component.cpp:358-366 (ocean_in_spinup is last call's flag) */
  if (!m->ocean_in_spinup && runToDate == (double)m->tracking_date)
    for (int i = 0; i < 4; ++i) m->box[i].tracking = 1;
  m->ocean_CO2_conc = co2_conc;
  m->SST = sst;
  m->ocean_in_spinup = m->in_spinup;
  m->annualflux_sum = m->annualflux_sumHL = m->annualflux_sumLL = 0.0;
  m->timesteps = 0;
  box_new_year(&m->box[HL], m->SST);
  box_new_year(&m->box[LL], m->SST);
  box_new_year(&m->box[IO], m->SST);
  box_new_year(&m->box[DO], m->SST);
  if (!m->p->spinup_chem && !m->ocean_in_spinup && !m->box[HL].active_chemistry) {
    m->box[HL].active_chemistry = 1;
    m->box[LL].active_chemistry = 1;
    box_chem_equilibrate(m, &m->box[HL], m->ocean_CO2_conc);
    box_chem_equilibrate(m, &m->box[LL], m->ocean_CO2_conc);
  }
  box_compute_fluxes(m, HL, m->ocean_CO2_conc, 1.0, 0);
  box_compute_fluxes(m, LL, m->ocean_CO2_conc, 1.0, 0);
}

/* ocean_component.cpp:603-626 */
static int ocean_calcderivs(member_t *m, double t, const double c[], double dcdt[]) {
  const double yearfraction = (t - m->ocean_ODEstartdate);
  const double cpooldiff = c[C_OCEAN] - ocean_totalcpool(m);
  const double surfacepools = FP(m, m->box[LL].carbon + m->box[HL].carbon);
  const double cpoolscale = (surfacepools + cpooldiff) / surfacepools;
  const double CO2_conc = c[C_ATMOS] * PGC_TO_PPMVCO2;
  dcdt[C_OCEAN] = ocean_annual_totalcflux(m, CO2_conc, cpoolscale);
  if (yearfraction > m->max_timestep) return CARBON_CYCLE_RETRY;
  return 0;
}

/* ocean_component.cpp:653-763 */
static void ocean_stashCValues(member_t *m, double t, const double c[]) {
  const double yearfraction = (t - m->ocean_ODEstartdate);
  if (!(yearfraction >= 0 && yearfraction <= 1)) fail_member(m, HO_ERR_YEARFRACTION);
  m->timesteps++;
  const int in_partial_year = (t != (int)t);
  const double CO2_conc = c[C_ATMOS] * PGC_TO_PPMVCO2;

  box_compute_fluxes(m, HL, CO2_conc, yearfraction, 1);
  box_compute_fluxes(m, LL, CO2_conc, yearfraction, 1);
  box_compute_fluxes(m, IO, CO2_conc, yearfraction, 1);
  box_compute_fluxes(m, DO, CO2_conc, yearfraction, 1);

  double currentflux = m->box[HL].atmosphere_flux + m->box[LL].atmosphere_flux;
  double solver_flux = c[C_OCEAN] - ocean_totalcpool(m);
  double adjustment = 0.0;
  if (currentflux) adjustment = (solver_flux - currentflux) / 2.0;
  m->box[HL].atmosphere_flux = m->box[HL].atmosphere_flux + adjustment;
  m->box[LL].atmosphere_flux = m->box[LL].atmosphere_flux + adjustment;
  box_separate_surface_fluxes(m, &m->box[HL]);
  box_separate_surface_fluxes(m, &m->box[LL]);

  double cflux_annualdiff = solver_flux / yearfraction - m->lastflux_annualized;
  if (cflux_annualdiff > OCEAN_TSR_TRIGGER1) {
    double r = m->max_timestep * OCEAN_TSR_FACTOR;
    m->max_timestep = OCEAN_MIN_TIMESTEP > r ? OCEAN_MIN_TIMESTEP : r; /* max(a, b) */
    m->reduced_timestep_timeout = OCEAN_TSR_TIMEOUT;
  } else if (!in_partial_year && m->reduced_timestep_timeout) {
    int d = m->reduced_timestep_timeout - 1;
    m->reduced_timestep_timeout = d > 0 ? d : 0;
    if (!m->reduced_timestep_timeout) {
      double r = m->max_timestep / OCEAN_TSR_FACTOR;
      m->max_timestep = r < OCEAN_MAX_TIMESTEP ? r : OCEAN_MAX_TIMESTEP; /* min(a, b) */
      if (m->max_timestep < OCEAN_MAX_TIMESTEP) m->reduced_timestep_timeout = OCEAN_TSR_TIMEOUT;
    }
  }

  double lastflux = m->box[LL].atmosphere_flux + m->box[HL].atmosphere_flux;
  m->annualflux_sumHL = m->annualflux_sumHL + m->box[HL].atmosphere_flux;
  m->annualflux_sumLL = m->annualflux_sumLL + m->box[LL].atmosphere_flux;
  m->annualflux_sum = m->annualflux_sum + lastflux;
  m->lastflux_annualized = lastflux / yearfraction;

  box_update_state(m, HL);
  box_update_state(m, LL);
  box_update_state(m, IO);
  box_update_state(m, DO);
  m->ocean_ODEstartdate = t;
}

/* M_DUMP_TO_DEEP_OCEAN: ocean_component.cpp:146-154 */
static void ocean_dump_to_deep(member_t *m, double carbon) {
  carbon = carbon + m->box[DO].carbon;
  /* set_carbon -> adjust_pool_to_val(C, allow_untracked = true): a positive difference enters
   * as source "untracked"; no sign check on the value (fluxpool.hpp:181-192) */
  box_t *b = &m->box[DO];
  const double diff = carbon - b->carbon;
  if (b->tracking && diff > 0) {
    tmap_t u;
    memset(&u, 0, sizeof u);
    tm_set_self(&u, TP_UNTRACKED);
    tmap_t adj = b->cmap;
    tm_add(m, b->carbon, &adj, FP(m, diff), &u);
    b->cmap = adj;
  }
  b->carbon = carbon;
}

/* ---------------------------------------------------------------------------------- */
/* SimpleNbox                                                                          */
static double snbox_CO2_conc(member_t *m) { /* simpleNbox.cpp:414-420 */
  return FP(m, m->atmos_c * PGC_TO_PPMVCO2);
}

/* sum_map over one per-biome field, in std::map (biome name) order: simpleNbox.cpp:428-438 */
#define SUM_MAP(m, field) sum_map_((m), offsetof(bio_t, field))
static double sum_map_(member_t *m, size_t off) {
  double sum = 0.0;
  for (int k = 0; k < m->nb; ++k)
    sum = FP(m, sum + *(const double *)((const char *)&m->bio[m->border[k]] + off));
  return sum;
}

static double snbox_npp(member_t *m, const bio_t *b) { /* simpleNbox-runtime.cpp:622-635 */
  double npp = FP(m, b->par.npp_flux0);
  npp = FP(m, npp * b->co2fert);
  npp = FP(m, npp * m->npp_luc_adjust);
  return npp;
}
static double snbox_rh_fda(member_t *m, const bio_t *b) { /* :653-665 */
  double dflux = FP(m, b->detritus_c * 0.25);
  return FP(m, dflux * b->tempfertd);
}
static double snbox_rh_fsa(member_t *m, const bio_t *b) { /* :671-683 */
  double soilflux = FP(m, b->soil_c * 0.02);
  return FP(m, soilflux * b->tempferts);
}
static double snbox_rh_ftpa_co2(member_t *m, const bio_t *b) { /* :689-701 */
  double tpfc = FP(m, b->thawed_permafrost_c * (1 - b->par.fpf_static));
  double tpflux = FP(m, tpfc * 0.02);
  double r = FP(m, tpflux * b->tempferts);
  return FP(m, r * (1.0 - b->par.rh_ch4_frac));
}
static double snbox_rh_ftpa_ch4(member_t *m, const bio_t *b) { /* :707-711 */
  double r = FP(m, snbox_rh_ftpa_co2(m, b) / (1.0 - b->par.rh_ch4_frac));
  return FP(m, r * b->par.rh_ch4_frac);
}
static double snbox_rh(member_t *m, const bio_t *b) { /* :717-721 */
  double r = FP(m, snbox_rh_fda(m, b) + snbox_rh_fsa(m, b));
  return FP(m, r + snbox_rh_ftpa_co2(m, b));
}

/* tseries::exists(year) && get(year) for a constraint series stored densely per model year
 * (NaN = no entry for that year) */
static int cn_get(const member_t *m, const double *series, int year, double *v) {
  if (!series) return 0;
  const int r = year - m->p->start_year;
  if (r < 0 || r >= m->nrow) return 0;
  if (isnan(series[r])) return 0;
  *v = series[r];
  return 1;
}

/* :744-772 */
static void snbox_compute_pf_thaw_refreeze(const bio_t *b, double rh_co2, double rh_ch4, double *x,
                                           double *y, double *z) {
  double biome_c_thaw = b->permafrost_c * b->f_new_thaw;
  double pf_refreeze_tp = 0.0, pf_refreeze_soil = 0.0;
  if (biome_c_thaw < 0) {
    const double pf_refreeze = -biome_c_thaw;
    biome_c_thaw = 0.0;
    const double thawed_remaining = b->thawed_permafrost_c - rh_co2 - rh_ch4;
    pf_refreeze_tp = pf_refreeze < thawed_remaining ? pf_refreeze : thawed_remaining; /* std::min */
  }
  *x = biome_c_thaw; *y = pf_refreeze_tp; *z = pf_refreeze_soil;
}

static void snbox_getCValues(member_t *m, double t, double c[]) { /* :247-258 */
  c[C_ATMOS] = m->atmos_c;
  c[C_VEG] = SUM_MAP(m, veg_c);
  c[C_DET] = SUM_MAP(m, detritus_c);
  c[C_SOIL] = SUM_MAP(m, soil_c);
  c[C_PERMAFROST] = SUM_MAP(m, permafrost_c);
  c[C_THAWEDP] = SUM_MAP(m, thawed_permafrost_c);
  c[C_OCEAN] = ocean_totalcpool(m);          /* ocean_component.cpp:587-591 */
  m->ocean_ODEstartdate = t;
  c[C_EARTH] = m->earth_c;
  m->snbox_ODEstartdate = t;
}

/* :781-934 */
static int snbox_calcderivs(member_t *m, double t, const double c[], double dcdt[]) {
  m->cnt.rhs_evals++;
  const int omodel_err = ocean_calcderivs(m, t, c, dcdt);
  const double ao_exchange = dcdt[C_OCEAN];
  double ocean_uptake = 0.0, ocean_release = 0.0;
  if (ao_exchange >= 0.0) ocean_uptake = FP(m, ao_exchange);
  else ocean_release = FP(m, -ao_exchange);

  double npp_current = 0.0, npp_fav = 0.0, npp_fad = 0.0, npp_fas = 0.0;
  double rh_fda_current = 0.0, rh_fsa_current = 0.0, rh_ftpa_co2_current = 0.0,
         rh_ftpa_ch4_current = 0.0;
  for (int ib = 0; ib < m->nb; ++ib) { /* biome_list order :809-821 */
    const bio_t *b = &m->bio[ib];
    const double npp_biome = snbox_npp(m, b);
    npp_current = FP(m, npp_current + npp_biome);
    npp_fav = FP(m, npp_fav + FP(m, npp_biome * b->par.f_nppv));
    npp_fad = FP(m, npp_fad + FP(m, npp_biome * b->par.f_nppd));
    npp_fas = FP(m, npp_fas + FP(m, npp_biome * (1 - b->par.f_nppv - b->par.f_nppd)));
    rh_fda_current = FP(m, rh_fda_current + snbox_rh_fda(m, b));
    rh_fsa_current = FP(m, rh_fsa_current + snbox_rh_fsa(m, b));
    rh_ftpa_co2_current = FP(m, rh_ftpa_co2_current + snbox_rh_ftpa_co2(m, b));
    rh_ftpa_ch4_current = FP(m, rh_ftpa_ch4_current + snbox_rh_ftpa_ch4(m, b));
  }
  double rh_current = FP(m, FP(m, rh_fda_current + rh_fsa_current) + rh_ftpa_co2_current);

  double litter_flux = 0.0, litter_fvd = 0.0, litter_fvs = 0.0, detsoil_flux = 0.0;
  for (int ib = 0; ib < m->nb; ++ib) { /* :826-839 */
    const bio_t *b = &m->bio[ib];
    const double v = FP(m, b->veg_c * 0.035);
    litter_flux = FP(m, litter_flux + v);
    litter_fvd = FP(m, litter_fvd + FP(m, v * b->par.f_litterd));
    litter_fvs = FP(m, litter_fvs + FP(m, v * (1 - b->par.f_litterd)));
  }
  for (int ib = 0; ib < m->nb; ++ib)
    detsoil_flux = FP(m, detsoil_flux + FP(m, m->bio[ib].detritus_c * 0.6));

  const double total = c[C_VEG] + c[C_DET] + c[C_SOIL];
  double luc_fva = FP(m, FP(m, m->current_luc_e * c[C_VEG]) / total);
  double luc_fda = FP(m, FP(m, m->current_luc_e * c[C_DET]) / total);
  double luc_fsa = FP(m, FP(m, m->current_luc_e * c[C_SOIL]) / total);
  double luc_fav = m->current_luc_u;
  double ch4ox_current = 0.0;

  double pf_thaw_c = 0.0, pf_refreeze_tp = 0.0, pf_refreeze_soil = 0.0;
  if (!m->in_spinup) {
    for (int ib = 0; ib < m->nb; ++ib) {
      const bio_t *b = &m->bio[ib];
      double x, y, z;
      snbox_compute_pf_thaw_refreeze(b, snbox_rh_ftpa_co2(m, b), snbox_rh_ftpa_ch4(m, b), &x, &y,
                                     &z);
      pf_thaw_c = FP(m, pf_thaw_c + FP(m, x));
      pf_refreeze_tp = FP(m, pf_refreeze_tp + FP(m, y));
      pf_refreeze_soil = FP(m, pf_refreeze_soil + FP(m, z));
    }
  }

  /* NBP constraint :871-898: NPP and RH (and their parts) are scaled so that their net
   * matches the user's NBP of year round(t) */
  {
    double nbp_c;
    if (!m->in_spinup && cn_get(m, m->cn ? m->cn->nbp : NULL, (int)round(t), &nbp_c)) {
      const double nbp = npp_current - rh_current - m->current_luc_e + m->current_luc_u;
      const double diff = nbp_c - nbp;
      const double npp_current_old = npp_current;
      npp_current = FP(m, npp_current + diff / 2.0);
      const double npp_ratio = npp_current / npp_current_old;
      npp_fav = FP(m, npp_fav * npp_ratio);
      npp_fad = FP(m, npp_fad * npp_ratio);
      npp_fas = FP(m, npp_fas * npp_ratio);
      const double rh_current_old = rh_current;
      rh_current = FP(m, rh_current - diff / 2.0);
      const double rh_ratio = rh_current / rh_current_old;
      rh_fda_current = FP(m, rh_fda_current * rh_ratio);
      rh_fsa_current = FP(m, rh_fsa_current * rh_ratio);
      rh_ftpa_co2_current = FP(m, rh_ftpa_co2_current * rh_ratio);
    }
  }

  dcdt[C_ATMOS] = m->current_ffi_e - m->current_daccs_u + m->current_luc_e - m->current_luc_u +
                  ch4ox_current - ocean_uptake + ocean_release - npp_current + rh_current;
  dcdt[C_VEG] = npp_fav - litter_flux - luc_fva + luc_fav;
  dcdt[C_DET] = npp_fad + litter_fvd - detsoil_flux - rh_fda_current - luc_fda;
  dcdt[C_SOIL] =
      npp_fas + litter_fvs + detsoil_flux - rh_fsa_current - pf_refreeze_soil - luc_fsa;
  dcdt[C_PERMAFROST] = -pf_thaw_c + pf_refreeze_soil + pf_refreeze_tp;
  dcdt[C_THAWEDP] = pf_thaw_c - pf_refreeze_tp - rh_ftpa_ch4_current - rh_ftpa_co2_current;
  dcdt[C_OCEAN] = ocean_uptake - ocean_release;
  dcdt[C_EARTH] = -m->current_ffi_e + m->current_daccs_u;
  return omodel_err;
}

/* lognormal cdf (Boost): erfc(-(ln x - mu)/(sigma*sqrt2))/2 */
static double lognormal_cdf(double mu, double sigma, double x) {
  if (x == 0) return 0;
  const double root_two = 1.41421356237309504880168872420969807856967187537694;
  double diff = (log(x) - mu) / (sigma * root_two);
  return erfc(-diff) / 2;
}

/* :945-1072; Tland = current D_LAND_TAS; row(t) indexes the scenario table */
static void snbox_slowparameval(member_t *m, double t, double Tland) {
  const ho_params *p = m->p;
  m->ocean_in_spinup = m->in_spinup; /* ocean_component.cpp:630-633 */
  if (m->in_spinup) {
    m->current_luc_e = m->current_luc_u = m->current_ffi_e = m->current_daccs_u = 0.0;
  } else {
    const double *row = m->raw + (size_t)((int)t - p->start_year) * HO_NRAW;
    m->current_luc_e = FP(m, row[HO_RAW_LUC_E]);
    m->current_luc_u = FP(m, row[HO_RAW_LUC_U]);
    m->current_ffi_e = FP(m, row[HO_RAW_FFI]);
    m->current_daccs_u = FP(m, row[HO_RAW_DACCS]);
  }
  m->npp_luc_adjust = (m->end_of_spinup_vegc - m->cum_luc_va) / m->end_of_spinup_vegc;

  for (int ib = 0; ib < m->nb; ++ib) {
    bio_t *b = &m->bio[ib];
    if (m->in_spinup) b->co2fert = 1.0;
    else b->co2fert = 1 + b->par.beta * log(snbox_CO2_conc(m) / p->C0); /* :614-616 */
  }

  for (int ib = 0; ib < m->nb; ++ib) {
    bio_t *b = &m->bio[ib];
    double tfs_last = 0.0;
    if (t > p->start_year && m->have_tempferts_last) tfs_last = b->tempferts_last_year;
    if (m->in_spinup) {
      b->tempfertd = 1.0;
      b->tempferts = 1.0;
      b->f_frozen = 1.0;
      b->f_new_thaw = 0.0;
      continue;
    }
    double wf = b->par.warmingfactor;
    const double Tland_biome = Tland * wf;
    b->tempfertd = pow(b->par.q10_rh, (Tland_biome / 10.0));
    b->f_new_thaw = 0.0;
    if (b->permafrost_c) {
      double f_frozen_current = 1.0;
      if (Tland_biome > 0)
        f_frozen_current = 1 - lognormal_cdf(b->par.pf_mu, b->par.pf_sigma, Tland_biome);
      b->f_new_thaw = b->f_frozen - f_frozen_current;
      b->f_frozen = f_frozen_current;
    }
    double Tland_rm = 0.0;
    if (t > p->start_year + 0) {
      for (int i = (int)(t - 0 - 200); i < t - 0; i++) {
        /* Tland_record.get(i): first key start+1; flat extrapolation below it */
        int k = i - p->start_year;
        if (k < 1) k = 1;
        Tland_rm += m->Tland_record[k] * wf;
      }
      Tland_rm /= 200;
    }
    b->tempferts = pow(b->par.q10_rh, (Tland_rm / 10.0));
    if (b->tempferts < tfs_last) b->tempferts = tfs_last;
  }
}

/* :270-609 */
static void snbox_stashCValues(member_t *m, double t, const double c[]) {
  const ho_params *p = m->p;
  const double yf = (t - m->snbox_ODEstartdate);
  if (!(yf >= 0 && yf <= 1)) fail_member(m, HO_ERR_YEARFRACTION);

  double ffi_flux = FP(m, m->current_ffi_e);   /* earth_c.flux_from_fluxpool */
  double ccs_flux = FP(m, m->current_daccs_u);
  /* Carbon tracking: a flux made by X.flux_from_fluxpool(..) carries a COPY of X's source map
   * as of that statement.  T = tracking on; *_0 = maps at the top of this stash. */
  const int T = m->tracking;
  const tmap_t atm_0 = m->tm[TP_ATMOS], earth_0 = m->tm[TP_EARTH], veg_0 = m->tm[TP_VEG],
               det_0 = m->tm[TP_DET], soil_0 = m->tm[TP_SOIL], perm_0 = m->tm[TP_PERMAFROST],
               thawed_0 = m->tm[TP_THAWEDP];

  ocean_stashCValues(m, t, c);
  tmap_t oa_map = m->box[LL].oamap, ao_map = m->box[LL].aomap;
  if (T) {
    tm_add(m, m->box[LL].oa_flux, &oa_map, m->box[HL].oa_flux, &m->box[HL].oamap);
    tm_add(m, m->box[LL].ao_flux, &ao_map, m->box[HL].ao_flux, &m->box[HL].aomap);
  }
  double oa_flux = FP(m, m->box[LL].oa_flux + m->box[HL].oa_flux); /* get_oaflux */
  double ao_flux = FP(m, m->box[LL].ao_flux + m->box[HL].ao_flux);

  double luc_e_untracked = m->current_luc_e, luc_u_untracked = m->current_luc_u;
  double npp_total = 0.0, rh_total = 0.0; /* sum_npp, sum_rh: biome_list order */
  for (int ib = 0; ib < m->nb; ++ib) npp_total = FP(m, npp_total + snbox_npp(m, &m->bio[ib]));
  for (int ib = 0; ib < m->nb; ++ib) rh_total = FP(m, rh_total + snbox_rh(m, &m->bio[ib]));
  const double permafrost_total = SUM_MAP(m, permafrost_c);

  double alf = npp_total - rh_total - luc_e_untracked + luc_u_untracked;
  double npp_rh_total = FP(m, npp_total + rh_total);

  double newatmos = FP(m, c[C_ATMOS]);
  double newveg = FP(m, c[C_VEG]);
  double newdet = FP(m, c[C_DET]);
  double newsoil = FP(m, c[C_SOIL]);
  double newpermafrost = FP(m, c[C_PERMAFROST]);
  double solver_tpf = c[C_THAWEDP];
  if (fabs(solver_tpf) < 1e-10) solver_tpf = 0.0;
  double newthawedpf = FP(m, solver_tpf);

  double rh_nbp_constraint_adjust = 1.0;
  /* NBP constraint :343-383 */
  {
    double nbp_c;
    const int rounded_t = (int)round(t);
    if (!m->in_spinup && cn_get(m, m->cn ? m->cn->nbp : NULL, rounded_t, &nbp_c)) {
      const double diff = nbp_c - alf;
      npp_total = FP(m, npp_total + diff / 2.0);
      rh_nbp_constraint_adjust = FP(m, rh_total - diff / 2.0) / rh_total;
      rh_total = FP(m, rh_total - diff / 2.0);
      const double pool_diff = diff * yf;
      const double total_land = c[C_DET] + c[C_VEG] + c[C_SOIL] + c[C_THAWEDP];
      newdet = FP(m, newdet + pool_diff * c[C_DET] / total_land);
      newveg = FP(m, newveg + pool_diff * c[C_VEG] / total_land);
      newsoil = FP(m, newsoil + pool_diff * c[C_SOIL] / total_land);
      newthawedpf = FP(m, newthawedpf + pool_diff * c[C_THAWEDP] / total_land);
      ocean_dump_to_deep(m, -pool_diff);
      alf = npp_total - rh_total - luc_e_untracked + luc_u_untracked;
    }
  }
  m->nbp = alf;

  const double total = c[C_VEG] + c[C_DET] + c[C_SOIL];
  const double luc_e = luc_e_untracked, luc_u = luc_u_untracked;
  m->cum_luc_va = m->cum_luc_va + ((luc_e - luc_u) * c[C_VEG] / total);

  for (int ib = 0; ib < m->nb; ++ib) { /* :399-531, biome_list order */
    bio_t *b = &m->bio[ib];
    const ho_biome *p = &b->par;
    const double wt = FP(m, snbox_npp(m, b) + snbox_rh(m, b)) / npp_rh_total;
    const double wt_pf = permafrost_total > 0 ? b->permafrost_c / permafrost_total : 0;

    const double veg_frac = b->veg_c / total;
    const double det_frac = b->detritus_c / total;
    const double soil_frac = b->soil_c / total;
    double luc_fva_biome_flux = FP(m, FP(m, FP(m, luc_e_untracked * veg_frac)) * yf);
    double luc_fda_biome_flux = FP(m, FP(m, FP(m, luc_e_untracked * det_frac)) * yf);
    double luc_fsa_biome_flux = FP(m, FP(m, FP(m, luc_e_untracked * soil_frac)) * yf);
    double luc_fav_biome_flux = FP(m, FP(m, luc_u_untracked) * yf);

    double npp_biome = FP(m, npp_total * wt);
    double npp_fav_biome_flux = FP(m, FP(m, FP(m, npp_biome * p->f_nppv)) * yf);
    double npp_fad_biome_flux = FP(m, FP(m, FP(m, npp_biome * p->f_nppd)) * yf);
    double npp_fas_biome_flux =
        FP(m, FP(m, FP(m, npp_biome * (1 - p->f_nppv - p->f_nppd))) * yf);

    double rh_fda_adj = FP(m, snbox_rh_fda(m, b) * rh_nbp_constraint_adjust);
    double rh_fsa_adj = FP(m, snbox_rh_fsa(m, b) * rh_nbp_constraint_adjust);
    double rh_ftpa_co2_adj = FP(m, snbox_rh_ftpa_co2(m, b) * rh_nbp_constraint_adjust);
    double rh_ftpa_ch4_adj = FP(m, snbox_rh_ftpa_ch4(m, b) * rh_nbp_constraint_adjust);
    /* final_npp = npp_biome (:429); final_rh = rh_fda_adj + rh_fsa_adj + rh_ftpa_co2_adj +
     * rh_ftpa_ch4_adj (:446-447) */
    b->final_npp = npp_biome;
    b->final_rh = FP(m, FP(m, FP(m, rh_fda_adj + rh_fsa_adj) + rh_ftpa_co2_adj) + rh_ftpa_ch4_adj);

    double rh_fda_flux = FP(m, FP(m, rh_fda_adj) * yf);
    double rh_fsa_flux = FP(m, FP(m, rh_fsa_adj) * yf);
    double rh_fpa_co2_flux = FP(m, FP(m, rh_ftpa_co2_adj) * yf);
    double rh_fpa_ch4_flux = FP(m, FP(m, rh_ftpa_ch4_adj) * yf);
    b->RH_ch4 = rh_fpa_ch4_flux;

    /* luc fluxes :458-462 */
    if (T) tm_add(m, m->atmos_c, &m->tm[TP_ATMOS], luc_fva_biome_flux, &veg_0);
    double a = FP(m, m->atmos_c + luc_fva_biome_flux);
    a = FP(m, a - luc_fav_biome_flux);
    if (T) tm_add(m, a, &m->tm[TP_ATMOS], luc_fda_biome_flux, &det_0);
    a = FP(m, a + luc_fda_biome_flux);
    if (T) tm_add(m, a, &m->tm[TP_ATMOS], luc_fsa_biome_flux, &soil_0);
    a = FP(m, a + luc_fsa_biome_flux);
    m->atmos_c = a;
    if (T) tm_add(m, b->veg_c, &m->tm[TP_VEG], luc_fav_biome_flux, &atm_0);
    double vg = FP(m, b->veg_c + luc_fav_biome_flux);
    vg = FP(m, vg - luc_fva_biome_flux);
    b->veg_c = vg;
    FP(m, b->detritus_c - luc_fda_biome_flux); /* :461, no effect except the throw */
    b->soil_c = FP(m, b->soil_c - luc_fsa_biome_flux);

    /* npp fluxes :465-469 */
    if (T) {
      tm_add(m, b->veg_c, &m->tm[TP_VEG], npp_fav_biome_flux, &atm_0);
      tm_add(m, b->detritus_c, &m->tm[TP_DET], npp_fad_biome_flux, &atm_0);
      tm_add(m, b->soil_c, &m->tm[TP_SOIL], npp_fas_biome_flux, &atm_0);
    }
    b->veg_c = FP(m, b->veg_c + npp_fav_biome_flux);
    b->detritus_c = FP(m, b->detritus_c + npp_fad_biome_flux);
    b->soil_c = FP(m, b->soil_c + npp_fas_biome_flux);
    a = FP(m, m->atmos_c - npp_fav_biome_flux);
    a = FP(m, a - npp_fad_biome_flux);
    a = FP(m, a - npp_fas_biome_flux);
    m->atmos_c = a;

    /* rh fluxes :472-481 */
    if (T) tm_add(m, m->atmos_c, &m->tm[TP_ATMOS], rh_fda_flux, &det_0);
    a = FP(m, m->atmos_c + rh_fda_flux);
    if (T) tm_add(m, a, &m->tm[TP_ATMOS], rh_fsa_flux, &soil_0);
    a = FP(m, a + rh_fsa_flux);
    if (T) tm_add(m, a, &m->tm[TP_ATMOS], rh_fpa_co2_flux, &thawed_0);
    a = FP(m, a + rh_fpa_co2_flux);
    m->atmos_c = a;
    b->detritus_c = FP(m, b->detritus_c - rh_fda_flux);
    b->soil_c = FP(m, b->soil_c - rh_fsa_flux);
    double tp = FP(m, b->thawed_permafrost_c - rh_fpa_co2_flux);
    tp = FP(m, tp - rh_fpa_ch4_flux);
    b->thawed_permafrost_c = tp;
    m->cumulative_pf_ch4 += rh_fpa_ch4_flux;

    if (!m->in_spinup) { /* :484-503 */
      double x, y, z;
      snbox_compute_pf_thaw_refreeze(b, rh_ftpa_co2_adj, rh_ftpa_ch4_adj, &x, &y, &z);
      double pf_thaw = FP(m, FP(m, FP(m, x)) * yf);
      double pf_refreeze_tp = FP(m, FP(m, FP(m, y)) * yf);
      double pf_refreeze_soil = FP(m, FP(m, FP(m, z)) * yf);
      /* pf_thaw carries permafrost's map, pf_refreeze_tp thawed permafrost's, pf_refreeze_soil
       * the soil's map as of now (after the luc and npp additions above) */
      const tmap_t soil_now = m->tm[TP_SOIL];
      double pc = FP(m, b->permafrost_c - pf_thaw);
      if (T) tm_add(m, pc, &m->tm[TP_PERMAFROST], pf_refreeze_tp, &thawed_0);
      pc = FP(m, pc + pf_refreeze_tp);
      if (T) tm_add(m, pc, &m->tm[TP_PERMAFROST], pf_refreeze_soil, &soil_now);
      pc = FP(m, pc + pf_refreeze_soil);
      b->permafrost_c = pc;
      if (T) tm_add(m, b->thawed_permafrost_c, &m->tm[TP_THAWEDP], pf_thaw, &perm_0);
      tp = FP(m, b->thawed_permafrost_c + pf_thaw);
      tp = FP(m, tp - pf_refreeze_tp);
      b->thawed_permafrost_c = tp;
      b->soil_c = FP(m, b->soil_c - pf_refreeze_soil);
    }

    /* litter :506-511 */
    double litter_flux = FP(m, b->veg_c * (0.035 * yf));
    double litter_fvd_flux = FP(m, litter_flux * p->f_litterd);
    double litter_fvs_flux = FP(m, litter_flux * (1 - p->f_litterd));
    if (T) {
      tm_add(m, b->detritus_c, &m->tm[TP_DET], litter_fvd_flux, &m->tm[TP_VEG]);
      tm_add(m, b->soil_c, &m->tm[TP_SOIL], litter_fvs_flux, &m->tm[TP_VEG]);
    }
    b->detritus_c = FP(m, b->detritus_c + litter_fvd_flux);
    b->soil_c = FP(m, b->soil_c + litter_fvs_flux);
    b->veg_c = FP(m, b->veg_c - litter_flux);

    /* detritus -> soil :514-521 */
    double detsoil_flux = FP(m, b->detritus_c * (0.6 * yf));
    if (T) tm_add(m, b->soil_c, &m->tm[TP_SOIL], detsoil_flux, &m->tm[TP_DET]);
    b->soil_c = FP(m, b->soil_c + detsoil_flux);
    b->detritus_c = FP(m, b->detritus_c - detsoil_flux);

    /* adjust to solver values (no sign check) :524-530 */
    b->veg_c = newveg * wt;
    b->detritus_c = newdet * wt;
    b->soil_c = newsoil * wt;
    b->permafrost_c = newpermafrost * wt_pf;
    b->thawed_permafrost_c = newthawedpf * wt_pf;
  }

  /* :534-541 */
  double e = FP(m, m->earth_c - ffi_flux);
  if (T) tm_add(m, e, &m->tm[TP_EARTH], ccs_flux, &atm_0);
  e = FP(m, e + ccs_flux);
  m->earth_c = e;
  if (T) tm_add(m, m->atmos_c, &m->tm[TP_ATMOS], ffi_flux, &earth_0);
  double a = FP(m, m->atmos_c + ffi_flux);
  a = FP(m, a - ccs_flux);
  if (T) tm_add(m, a, &m->tm[TP_ATMOS], oa_flux, &oa_map);
  a = FP(m, a + oa_flux);
  a = FP(m, a - ao_flux);
  m->atmos_c = a;
  m->earth_c = c[C_EARTH];
  m->atmos_c = newatmos;

  /* mass balance :546-564 */
  double sum = 0.0;
  for (int i = 0; i < NC; i++) sum += c[i];
  sum += m->cumulative_pf_ch4;
  const double diff = fabs(sum - m->masstot);
  if (m->masstot > 0.0 && diff > MB_EPSILON) fail_member(m, HO_ERR_MASS);
  m->masstot = sum;

  /* spin-up pinning / CO2 constraint :567-603 (CO2_constrain.exists(t): exact key only, so a
   * stash that ends inside a year is never constrained) */
  double co2_c = 0.0;
  const int have_co2_c = !m->in_spinup && t == floor(t) &&
                         cn_get(m, m->cn ? m->cn->co2 : NULL, (int)t, &co2_c);
  if (have_co2_c) {
    FP(m, co2_c); /* fluxpool atmppmv.set(...) */
    double atmos_cpool_to_match = FP(m, co2_c / PGC_TO_PPMVCO2);
    double Ca_residual = m->atmos_c - atmos_cpool_to_match;
    ocean_dump_to_deep(m, Ca_residual);
    m->atmos_c = FP(m, m->atmos_c - Ca_residual);
  }
  if (m->in_spinup) {
    double atmos_cpool_to_match = FP(m, p->C0 / PGC_TO_PPMVCO2);
    double Ca_residual = m->atmos_c - atmos_cpool_to_match;
    ocean_dump_to_deep(m, Ca_residual);
    m->atmos_c = FP(m, m->atmos_c - Ca_residual);
  }
  m->snbox_ODEstartdate = t;
}

/* record_state: simpleNbox.cpp:789-840 (only the part with side effects on the model) */
static void snbox_record_state(member_t *m) {
  for (int ib = 0; ib < m->nb; ++ib) {
    bio_t *b = &m->bio[ib];
    if (!m->in_spinup) {
      (void)snbox_npp(m, b);
      (void)snbox_rh_fda(m, b);
      (void)snbox_rh_fsa(m, b);
      (void)snbox_rh_ftpa_co2(m, b);
      b->RH_ch4 = snbox_rh_ftpa_ch4(m, b);
    } else {
      b->RH_ch4 = 0.0;
    }
    b->tempferts_last_year = b->tempferts;
  }
  m->have_tempferts_last = 1;
}

/* ---------------------------------------------------------------------------------- */
/* CarbonCycleSolver + odeint                                                          */
typedef struct {
  int first_call;
  double dxdt[NC];
} stepper_t;

/* returns 0 = success, 1 = step rejected, or a calcderivs status (CARBON_CYCLE_RETRY) */
static int try_step(member_t *m, stepper_t *st, double x[], double *t, double *dt) {
  const ho_params *p = m->p;
  const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
  const double b21 = 1.0 / 5.0;
  const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
  const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
  const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0,
               b54 = -212.0 / 729.0;
  const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0,
               b64 = 49.0 / 176.0, b65 = -5103.0 / 18656.0;
  const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0,
               c6 = 11.0 / 84.0;
  const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0,
               dc5 = c5 - (-92097.0 / 339200.0), dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;
  double tmp[NC], k2[NC], k3[NC], k4[NC], k5[NC], k6[NC], xnew[NC], dxdtnew[NC], xerr[NC];
  double f1, f2, f3, f4, f5, f6;
  int rc;
  const double tt = *t, h = *dt;
  if (st->first_call) {
    if ((rc = snbox_calcderivs(m, tt, x, st->dxdt))) return rc;
    st->first_call = 0;
  }
  const double *k1 = st->dxdt;
  f1 = h * b21;
  for (int i = 0; i < NC; ++i) tmp[i] = 1.0 * x[i] + f1 * k1[i];
  if ((rc = snbox_calcderivs(m, tt + h * a2, tmp, k2))) return rc;
  f1 = h * b31; f2 = h * b32;
  for (int i = 0; i < NC; ++i) tmp[i] = 1.0 * x[i] + f1 * k1[i] + f2 * k2[i];
  if ((rc = snbox_calcderivs(m, tt + h * a3, tmp, k3))) return rc;
  f1 = h * b41; f2 = h * b42; f3 = h * b43;
  for (int i = 0; i < NC; ++i) tmp[i] = 1.0 * x[i] + f1 * k1[i] + f2 * k2[i] + f3 * k3[i];
  if ((rc = snbox_calcderivs(m, tt + h * a4, tmp, k4))) return rc;
  f1 = h * b51; f2 = h * b52; f3 = h * b53; f4 = h * b54;
  for (int i = 0; i < NC; ++i)
    tmp[i] = 1.0 * x[i] + f1 * k1[i] + f2 * k2[i] + f3 * k3[i] + f4 * k4[i];
  if ((rc = snbox_calcderivs(m, tt + h * a5, tmp, k5))) return rc;
  f1 = h * b61; f2 = h * b62; f3 = h * b63; f4 = h * b64; f5 = h * b65;
  for (int i = 0; i < NC; ++i)
    tmp[i] = 1.0 * x[i] + f1 * k1[i] + f2 * k2[i] + f3 * k3[i] + f4 * k4[i] + f5 * k5[i];
  if ((rc = snbox_calcderivs(m, tt + h, tmp, k6))) return rc;
  f1 = h * c1; f2 = h * c3; f3 = h * c4; f4 = h * c5; f5 = h * c6;
  for (int i = 0; i < NC; ++i)
    xnew[i] = 1.0 * x[i] + f1 * k1[i] + f2 * k3[i] + f3 * k4[i] + f4 * k5[i] + f5 * k6[i];
  if ((rc = snbox_calcderivs(m, tt + h, xnew, dxdtnew))) return rc;
  f1 = h * dc1; f2 = h * dc3; f3 = h * dc4; f4 = h * dc5; f5 = h * dc6; f6 = h * dc7;
  for (int i = 0; i < NC; ++i)
    xerr[i] = f1 * k1[i] + f2 * k3[i] + f3 * k4[i] + f4 * k5[i] + f5 * k6[i] + f6 * dxdtnew[i];

  double max_rel_err = 0.0;
  const double a_dxdt = 1.0 * fabs(h);
  for (int i = 0; i < NC; ++i) {
    double e = fabs(xerr[i]) / (p->eps_abs + p->eps_rel * (1.0 * fabs(x[i]) + a_dxdt * fabs(k1[i])));
    max_rel_err = max_rel_err > e ? max_rel_err : e;
  }
  if (max_rel_err > 1.0) {
    double f = 9.0 / 10.0 * pow(max_rel_err, -1.0 / (4.0 - 1.0));
    *dt *= (f > 1.0 / 5.0 ? f : 1.0 / 5.0);
    m->cnt.steps_rejected++;
    return 1;
  }
  *t += h;
  if (max_rel_err < 0.5) {
    double error = pow(5.0, -5.0);
    error = error > max_rel_err ? error : max_rel_err;
    *dt *= 9.0 / 10.0 * pow(error, -1.0 / 5.0);
  }
  for (int i = 0; i < NC; ++i) { x[i] = xnew[i]; st->dxdt[i] = dxdtnew[i]; }
  m->cnt.steps_accepted++;
  return 0;
}

/* integrate_adaptive; the observer writes t into the solver (carbon-cycle-solver.cpp:196-200) */
static int integrate_adaptive(member_t *m, double x[], double t, double t_end, double dt) {
  stepper_t st;
  st.first_call = 1;
  m->cnt.integrate_calls++;
  while (t_end - t > DBL_EPSILON) {
    m->t = t;
    if ((t + dt) - t_end > DBL_EPSILON) dt = t_end - t;
    int rc, fails = 0;
    do {
      rc = try_step(m, &st, x, &t, &dt);
      if (rc > 1) return rc; /* bad_derivative_exception */
      if (rc == 1 && ++fails >= 500) fail_member(m, HO_ERR_STEPPER);
    } while (rc == 1);
  }
  m->t = t;
  return 0;
}

/* CarbonCycleSolver::run: carbon-cycle-solver.cpp:222-303 */
static void solver_run(member_t *m, const double tnew, double Tland) {
  double *c = m->c;
  snbox_getCValues(m, m->t, c);
  snbox_slowparameval(m, m->t, Tland);
  int retry = 0;
  while (m->t < tnew && retry < MAX_CARBON_MODEL_RETRIES) {
    double t_start = m->t;
    double t_target = tnew;
    while (m->t < t_target && retry < MAX_CARBON_MODEL_RETRIES) {
      int stat = integrate_adaptive(m, c, t_start, t_target, m->dt);
      if (stat == CARBON_CYCLE_RETRY) {
        ++retry;
        t_target = t_start + (t_target - t_start) / 2.0;
        m->t = t_start;
        m->dt = t_target - m->t;
        snbox_getCValues(m, m->t, c);
      }
    }
    if (retry < MAX_CARBON_MODEL_RETRIES) {
      retry = 0;
      snbox_stashCValues(m, m->t, c);
    }
  }
  if (!(m->t == tnew)) fail_member(m, HO_ERR_RETRIES);
  snbox_record_state(m);
}

/* ---------------------------------------------------------------------------------- */
/* member-independent gas series                                                       */
void ho_gas_series(const ho_params *p, const double *raw, double *n2o, double *halo_rf) {
  ho_gas_series_constrained(p, raw, NULL, n2o, halo_rf);
}

/* with optional N2O / halocarbon concentration constraints (n2o_component.cpp:136-160,
 * halocarbon_component.cpp:189-192).  Returns N0 as N2OComponent::prepareToRun leaves it. */
double ho_gas_series_constrained(const ho_params *p, const double *raw, const ho_constraints *cn,
                                 double *n2o, double *halo_rf) {
  const int nrow = p->end_year - p->start_year + 1;
  double N0 = p->N0;
  if (cn && cn->n2o && !isnan(cn->n2o[0])) N0 = cn->n2o[0]; /* n2o_component.cpp:141-146 */
  n2o[0] = N0; /* n2o_component.cpp:136-147 */
  for (int r = 1; r < nrow; ++r) { /* :150-191 */
    if (cn && cn->n2o && !isnan(cn->n2o[r])) {
      n2o[r] = cn->n2o[r];
      continue;
    }
    const double *row = raw + (size_t)r * HO_NRAW;
    double previous_n2o = n2o[r - 1];
    double tau = p->TN2O0 * (pow(previous_n2o / N0, -0.05));
    const double current_n2oem = row[HO_RAW_N2O_E] + row[HO_RAW_N2O_NAT];
    const double dN2O = current_n2oem / p->UC_N2O - previous_n2o / tau;
    n2o[r] = previous_n2o + dN2O;
  }
  for (int g = 0; g < HO_NHALO; ++g) { /* halocarbon_component.cpp:181-229 */
    double Ha = p->halo_H0[g];
    halo_rf[g] = 0.0; /* row 0 never read */
    const double tau = p->halo_tau[g];
    for (int r = 1; r < nrow; ++r) {
      const double *row = raw + (size_t)r * HO_NRAW;
      const double timestep = 1.0;
      const double alpha = 1 / tau;
      double emissMol = row[HO_RAW_HALO0 + g] / p->halo_molarMass[g] * timestep;
      double concDeltaEmiss = emissMol / (0.1 * 1.8);
      double expfac = exp(-alpha);
      Ha = Ha * expfac + concDeltaEmiss * tau * (1.0 - expfac);
      if (cn && cn->halo && !isnan(cn->halo[(size_t)r * HO_NHALO + g]))
        Ha = cn->halo[(size_t)r * HO_NHALO + g]; /* concentration-forced year */
      double rf_unadjusted = p->halo_rho[g] * Ha;
      double adjusted_rf = rf_unadjusted + p->halo_delta[g] * rf_unadjusted;
      halo_rf[(size_t)r * HO_NHALO + g] = adjusted_rf;
    }
  }
  return N0;
}

/* ---------------------------------------------------------------------------------- */
/* DOECLIM                                                                             */
static const double dc_dt = 1, dc_ak = 0.31, dc_bk = 1.59, dc_csw = 0.13,
                    dc_earth_area = 5100656E8, dc_rlam = 1.43, dc_zbot = 4000.0, dc_bsi = 1.3,
                    dc_cal = 0.52, dc_cas = 7.80, dc_flnd = 0.29, dc_fso = 0.95;
#define DC_SECS_PER_YEAR (60.0 * 60.0 * 24.0 * 365.2422)

static void doeclim_kernel(double taubot, int ns, double *Ker) {
  /* temperature_component.cpp:303-371 */
  const double dt = dc_dt;
  for (int i = 0; i < ns; ++i) {
    double KT0, KTA1, KTB1, KTA2, KTB2, KTA3, KTB3;
    if (i == ns - 1) {
      KT0 = 4.0 - 2.0 * pow(2.0, 0.5);
      KTA1 = -8.0 * exp(-taubot / dt) + 4.0 * pow(2.0, 0.5) * exp(-0.5 * taubot / dt);
      KTB1 = 4.0 * pow((M_PI * taubot / dt), 0.5) *
             (1.0 + erf(pow(0.5 * taubot / dt, 0.5)) - 2.0 * erf(pow(taubot / dt, 0.5)));
      KTA2 = 8.0 * exp(-4.0 * taubot / dt) - 4.0 * pow(2.0, 0.5) * exp(-2.0 * taubot / dt);
      KTB2 = -8.0 * pow((M_PI * taubot / dt), 0.5) *
             (1.0 + erf(pow((2.0 * taubot / dt), 0.5)) - 2.0 * erf(2.0 * pow((taubot / dt), 0.5)));
      KTA3 = -8.0 * exp(-9.0 * taubot / dt) + 4.0 * pow(2.0, 0.5) * exp(-4.5 * taubot / dt);
      KTB3 = 12.0 * pow((M_PI * taubot / dt), 0.5) *
             (1.0 + erf(pow((4.5 * taubot / dt), 0.5)) - 2.0 * erf(3.0 * pow((taubot / dt), 0.5)));
    } else {
      KT0 = 4.0 * pow((double)(ns - i), 0.5) - 2.0 * pow((double)(ns + 1 - i), 0.5) -
            2.0 * pow((double)(ns - 1 - i), 0.5);
      KTA1 = -8.0 * pow((double)(ns - i), 0.5) * exp(-taubot / dt / (double)(ns - i)) +
             4.0 * pow((double)(ns + 1 - i), 0.5) * exp(-taubot / dt / (double)(ns + 1 - i)) +
             4.0 * pow((double)(ns - 1 - i), 0.5) * exp(-taubot / dt / (double)(ns - 1 - i));
      KTB1 = 4.0 * pow((M_PI * taubot / dt), 0.5) *
             (erf(pow((taubot / dt / (double)(ns - 1 - i)), 0.5)) +
              erf(pow((taubot / dt / (double)(ns + 1 - i)), 0.5)) -
              2.0 * erf(pow((taubot / dt / (double)(ns - i)), 0.5)));
      KTA2 = 8.0 * pow((double)(ns - i), 0.5) * exp(-4.0 * taubot / dt / (double)(ns - i)) -
             4.0 * pow((double)(ns + 1 - i), 0.5) * exp(-4.0 * taubot / dt / (double)(ns + 1 - i)) -
             4.0 * pow((double)(ns - 1 - i), 0.5) * exp(-4.0 * taubot / dt / (double)(ns - 1 - i));
      KTB2 = -8.0 * pow((M_PI * taubot / dt), 0.5) *
             (erf(2.0 * pow((taubot / dt / (double)(ns - 1 - i)), 0.5)) +
              erf(2.0 * pow((taubot / dt / (double)(ns + 1 - i)), 0.5)) -
              2.0 * erf(2.0 * pow((taubot / dt / (double)(ns - i)), 0.5)));
      KTA3 = -8.0 * pow((double)(ns - i), 0.5) * exp(-9.0 * taubot / dt / (double)(ns - i)) +
             4.0 * pow((double)(ns + 1 - i), 0.5) * exp(-9.0 * taubot / dt / (double)(ns + 1 - i)) +
             4.0 * pow((double)(ns - 1 - i), 0.5) * exp(-9.0 * taubot / dt / (double)(ns - 1 - i));
      KTB3 = 12.0 * pow((M_PI * taubot / dt), 0.5) *
             (erf(3.0 * pow((taubot / dt / (double)(ns - 1 - i)), 0.5)) +
              erf(3.0 * pow((taubot / dt / (double)(ns + 1 - i)), 0.5)) -
              2.0 * erf(3.0 * pow((taubot / dt / (double)(ns - i)), 0.5)));
    }
    Ker[i] = KT0 + KTA1 + KTB1 + KTA2 + KTB2 + KTA3 + KTB3;
  }
}

static double doeclim_taubot(double diff) {
  double kcon = DC_SECS_PER_YEAR / 10000;
  double keff = kcon * diff;
  return pow(dc_zbot, 2) / keff;
}

void ho_doeclim_kernel(double diff, int ns, double *ker) {
  doeclim_kernel(doeclim_taubot(diff), ns, ker);
}

/* temperature_component.cpp:196-413 */
static void doeclim_prepareToRun(member_t *m) {
  const ho_params *p = m->p;
  const double dt = dc_dt, ak = dc_ak, bk = dc_bk, csw = dc_csw, rlam = dc_rlam, bsi = dc_bsi,
               cal = dc_cal, cas = dc_cas, flnd = dc_flnd, fso = dc_fso;
  const double S = p->S, qco2 = p->qco2, diff = p->diff;
  int ns = p->end_year - p->start_year + 1;
  m->ns = ns;
  double kcon = DC_SECS_PER_YEAR / 10000;
  double ocean_area = (1.0 - flnd) * dc_earth_area;
  double cnum = rlam * flnd + bsi * (1.0 - flnd);
  double cden = rlam * flnd - ak * (rlam - bsi);
  double cfl = flnd * cnum / cden * qco2 / S - bk * (rlam - bsi) / cden;
  double cfs = (rlam * flnd - ak / (1.0 - flnd) * (rlam - bsi)) * cnum / cden * qco2 / S +
               rlam * flnd / (1.0 - flnd) * bk * (rlam - bsi) / cden;
  double kls = bk * rlam * flnd / cden - ak * flnd * cnum / cden * qco2 / S;
  double keff = kcon * diff;
  m->powtoheat = ocean_area * DC_SECS_PER_YEAR / pow(10.0, 22);
  double taubot = pow(dc_zbot, 2) / keff;
  double taucfs = cas / cfs;
  double taucfl = cal / cfl;
  double taudif = pow(cas, 2) / pow(csw, 2) * M_PI / keff;
  double tauksl = (1.0 - flnd) * cas / kls;
  double taukls = flnd * cal / kls;
  m->taucfl = taucfl; m->taukls = taukls; m->taucfs = taucfs; m->tauksl = tauksl;
  m->taudif = taudif;

  doeclim_kernel(taubot, ns, m->Ker);
  double *C = m->Cc, *A = m->A, *B = m->B;
  C[0] = 1.0 / pow(taucfl, 2.0) + 1.0 / pow(taukls, 2.0) + 2.0 / taucfl / taukls +
         bsi / taukls / tauksl;
  C[1] = -1 * bsi / pow(taukls, 2.0) - bsi / taucfl / taukls - bsi / taucfs / taukls -
         pow(bsi, 2.0) / taukls / tauksl;
  C[2] = -1 * bsi / pow(tauksl, 2.0) - 1.0 / taucfs / tauksl - 1.0 / taucfl / tauksl -
         1.0 / taukls / tauksl;
  C[3] = 1.0 / pow(taucfs, 2.0) + pow(bsi, 2.0) / pow(tauksl, 2.0) + 2.0 * bsi / taucfs / tauksl +
         bsi / taukls / tauksl;
  for (int i = 0; i < 4; i++) C[i] = C[i] * (pow(dt, 2.0) / 12.0);
  B[0] = 1.0 + dt / (2.0 * taucfl) + dt / (2.0 * taukls);
  B[1] = -dt / (2.0 * taukls) * bsi;
  B[2] = -dt / (2.0 * tauksl);
  B[3] = 1.0 + dt / (2.0 * taucfs) + dt / (2.0 * tauksl) * bsi + 2.0 * fso * pow((dt / taudif), 0.5);
  A[0] = 1.0 - dt / (2.0 * taucfl) - dt / (2.0 * taukls);
  A[1] = dt / (2.0 * taukls) * bsi;
  A[2] = dt / (2.0 * tauksl);
  A[3] = 1.0 - dt / (2.0 * taucfs) - dt / (2.0 * tauksl) * bsi +
         m->Ker[ns - 1] * fso * pow((dt / taudif), 0.5);
  for (int i = 0; i < 4; i++) {
    B[i] = B[i] + C[i];
    A[i] = A[i] + C[i];
  }
  /* invert_1d_2x2_matrix: temperature_component.cpp:81-94 */
  double temp_d = (B[0] * B[3] - B[1] * B[2]);
  double temp = 1 / temp_d;
  m->IB[0] = temp * B[3];
  m->IB[1] = temp * -1 * B[1];
  m->IB[2] = temp * -1 * B[2];
  m->IB[3] = temp * B[0];
  m->tas = m->tas_land = m->sst = m->heatflux = 0.0;
}

/* temperature_component.cpp:417-557 */
static void doeclim_run(member_t *m, int tstep, double rf_tot) {
  const double dt = dc_dt, bsi = dc_bsi, cal = dc_cal, cas = dc_cas, flnd = dc_flnd, fso = dc_fso;
  const double taucfl = m->taucfl, taukls = m->taukls, taucfs = m->taucfs, tauksl = m->tauksl,
               taudif = m->taudif;
  const int ns = m->ns;
  double *forcing = m->forcing, *temp_sst = m->temp_sst, *temp_landair = m->temp_landair;
  const double *Ker = m->Ker, *A = m->A, *IB = m->IB;
  forcing[tstep] = rf_tot;
  double DQ1 = 0.0, DQ2 = 0.0, QC1 = 0.0, QC2 = 0.0, DelQL = 0.0, DelQO = 0.0, DPAST1 = 0.0,
         DPAST2 = 0.0, DTEAUX1 = 0.0, DTEAUX2 = 0.0;
  m->temp[tstep] = 0.0;
  temp_landair[tstep] = 0.0;
  temp_sst[tstep] = 0.0;
  m->heat_mixed[tstep] = 0.0;
  m->heat_interior[tstep] = 0.0;
  m->heatflux_mixed[tstep] = 0.0;
  m->heatflux_interior[tstep] = 0.0;
  const double *QL = forcing, *QO = forcing;
  if (tstep > 0) {
    DelQL = QL[tstep] - QL[tstep - 1];
    DelQO = QO[tstep] - QO[tstep - 1];
    QC1 = (DelQL / cal * (1.0 / taucfl + 1.0 / taukls) - bsi * DelQO / cas / taukls);
    QC2 = (DelQO / cas * (1.0 / taucfs + bsi / tauksl) - DelQL / cal / tauksl);
    QC1 = QC1 * pow(dt, 2.0) / 12.0;
    QC2 = QC2 * pow(dt, 2.0) / 12.0;
    DQ1 = 0.5 * dt / cal * (QL[tstep] + QL[tstep - 1]);
    DQ2 = 0.5 * dt / cas * (QO[tstep] + QO[tstep - 1]);
    DQ1 = DQ1 + QC1;
    DQ2 = DQ2 + QC2;
    for (int i = 0; i <= tstep; i++) DPAST2 = DPAST2 + temp_sst[i] * Ker[ns - tstep + i - 1];
    DPAST2 = DPAST2 * fso * pow((dt / taudif), 0.5);
    DTEAUX1 = A[0] * temp_landair[tstep - 1] + A[1] * temp_sst[tstep - 1];
    DTEAUX2 = A[2] * temp_landair[tstep - 1] + A[3] * temp_sst[tstep - 1];
    temp_landair[tstep] = IB[0] * (DQ1 + DPAST1 + DTEAUX1) + IB[1] * (DQ2 + DPAST2 + DTEAUX2);
    temp_sst[tstep] = IB[2] * (DQ1 + DPAST1 + DTEAUX1) + IB[3] * (DQ2 + DPAST2 + DTEAUX2);
  } else {
    temp_landair[0] = 0.0;
    temp_sst[0] = 0.0;
  }
  m->temp[tstep] = flnd * temp_landair[tstep] + (1.0 - flnd) * bsi * temp_sst[tstep];
  /* user-supplied global temperature (:510-525): overwrite, then back-calculate land and
   * sea-surface values */
  if (m->cn && m->cn->tas && tstep >= m->cn->tas_first_row && tstep <= m->cn->tas_last_row) {
    m->temp[tstep] = m->cn->tas[tstep];
    temp_landair[tstep] = (m->temp[tstep] - (1.0 - flnd) * bsi * temp_sst[tstep]) / flnd;
    temp_sst[tstep] = (m->temp[tstep] - flnd * temp_landair[tstep]) / ((1.0 - flnd) * bsi);
  }
  if (tstep > 0) {
    m->heatflux_mixed[tstep] = cas * (temp_sst[tstep] - temp_sst[tstep - 1]);
    for (int i = 0; i < tstep; i++)
      m->heatflux_interior[tstep] = m->heatflux_interior[tstep] + temp_sst[i] * Ker[ns - tstep + i];
    m->heatflux_interior[tstep] =
        cas * fso / pow((taudif * dt), 0.5) * (2.0 * temp_sst[tstep] - m->heatflux_interior[tstep]);
    m->heat_mixed[tstep] = m->heat_mixed[tstep - 1] + m->heatflux_mixed[tstep] * (m->powtoheat * dt);
    m->heat_interior[tstep] =
        m->heat_interior[tstep - 1] + m->heatflux_interior[tstep] * (fso * m->powtoheat * dt);
  }
  /* setoutputs: :706-746 */
  m->heatflux = m->heatflux_mixed[tstep] + fso * m->heatflux_interior[tstep];
  m->tas = m->temp[tstep];
  m->tas_land = temp_landair[tstep];
  m->sst = temp_sst[tstep];
  m->tas_ocean = bsi * temp_sst[tstep];
  /* user-provided land-ocean warming ratio (:722-739): what getData(land_tas / ocean_tas / sst)
   * hands to the other components and to callers (:586-622); DOECLIM's own arrays stay */
  if (m->p->lo_warming_ratio != 0) {
    const double r = m->p->lo_warming_ratio;
    const double temp_oceanair_constrain = m->temp[tstep] / ((r * flnd) + (1 - flnd));
    const double temp_landair_constrain = temp_oceanair_constrain * r;
    const double temp_sst_constrain = temp_oceanair_constrain / bsi;
    m->tas_land = temp_landair_constrain;
    m->sst = temp_sst_constrain;
    m->tas_ocean = temp_oceanair_constrain;
  }
}

/* ---------------------------------------------------------------------------------- */
/* forcing: forcing_component.cpp:300-532.  Returns absolute total; *fco2 etc. absolute. */
static double forcing_total(member_t *m, int r, double CO2_conc, double Ma, double Na,
                            double ozone, double *fco2_out, double *fch4_out, double *fn2o_out) {
  const ho_params *p = m->p;
  const double a1 = -2.4785e-7, b1 = 7.5906e-4, c1 = -2.1492e-3, d1 = 5.2488;
  const double a2 = -3.4197e-4, b2 = 2.5455e-4, c2 = -2.4357e-4, d2 = 0.12173;
  const double a3 = -8.9603e-5, b3 = -1.2462e-4, d3 = 0.045194;
  const double aci_beta = 2.279759, s_BCOC = 111.05064063,
               s_SO2 = (260.34644166 * 1000) * (32.065 / 64.066);
  const double *row = m->raw + (size_t)r * HO_NRAW;
  const double *hrf = m->halo_rf + (size_t)r * HO_NHALO;
  const double C0 = p->C0, M0 = m->M0_ch4, N0 = m->N0_n2o;

  double C_alpha_max = C0 - (b1 / (2 * a1));
  double n2o_alpha = c1 * sqrt(Na);
  double alpha_prime = 0;
  if (CO2_conc > C_alpha_max) alpha_prime = d1 - (pow(b1, 2) / (4 * a1));
  else if (C0 < CO2_conc && CO2_conc < C_alpha_max)
    alpha_prime = d1 + a1 * pow((CO2_conc - C0), 2) + b1 * (CO2_conc - C0);
  else if (CO2_conc <= C0) alpha_prime = d1;
  else fail_member(m, HO_ERR_CO2SARF);
  double sarf_co2 = (alpha_prime + n2o_alpha) * log(CO2_conc / C0);
  double fco2 = (sarf_co2 * p->delta_co2) + sarf_co2;
  double sarf_n2o = (a2 * sqrt(CO2_conc) + b2 * sqrt(Na) + c2 * sqrt(Ma) + d2) * (sqrt(Na) - sqrt(N0));
  double fn2o = (p->delta_n2o * sarf_n2o) + sarf_n2o;
  double sarf_ch4 = (a3 * sqrt(Ma) + b3 * sqrt(Na) + d3) * (sqrt(Ma) - sqrt(M0));
  double fch4 = (p->delta_ch4 * sarf_ch4) + sarf_ch4;
  const double Ma_base = 1831, stratH2O_base = 0.0485;
  const double fh2o_strat = stratH2O_base * ((Ma - M0) / (Ma_base - M0));
  const double fo3_trop = 0.042 * ozone;
  double E_BC = row[HO_RAW_BC], E_OC = row[HO_RAW_OC], E_SO2 = row[HO_RAW_SO2], E_NH3 = row[HO_RAW_NH3];
  double alpha = p->aero_scalar;
  double fbc = alpha * p->rho_bc * E_BC;
  double foc = alpha * p->rho_oc * E_OC;
  double fso2 = alpha * p->rho_so2 * E_SO2;
  double fnh3 = alpha * p->rho_nh3 * E_NH3;
  double aci_rf = alpha * (-1 * aci_beta * log(1 + (E_SO2 / s_SO2) + ((E_BC + E_OC) / s_BCOC)));
  double falbedo = row[HO_RAW_ALBEDO];
  double fvol = p->vol_scalar * row[HO_RAW_SV];
  double fmisc = row[HO_RAW_MISC];

  /* halo index by name, order of forcing_component.cpp:402-409 */
  enum { CF4, C2F6, HFC23, HFC32, HFC4310, HFC125, HFC134a, HFC143a, HFC227ea, HFC245fa, SF6,
         CFC11, CFC12, CFC113, CFC114, CFC115, CCl4, CH3CCl3, HCFC22, HCFC141b, HCFC142b,
         halon1211, halon1301, halon2402, CH3Cl, CH3Br };
  /* std::map<string, unitval> iteration = byte-wise key order (forcing_component.cpp:492-495) */
  double Ftot = 0.0;
  Ftot = Ftot + fbc;             /* RF_BC */
  Ftot = Ftot + hrf[C2F6];       /* RF_C2F6 */
  Ftot = Ftot + hrf[CCl4];       /* RF_CCl4 */
  Ftot = Ftot + hrf[CF4];        /* RF_CF4 */
  Ftot = Ftot + hrf[CFC11];      /* RF_CFC11 */
  Ftot = Ftot + hrf[CFC113];     /* RF_CFC113 */
  Ftot = Ftot + hrf[CFC114];     /* RF_CFC114 */
  Ftot = Ftot + hrf[CFC115];     /* RF_CFC115 */
  Ftot = Ftot + hrf[CFC12];      /* RF_CFC12 */
  Ftot = Ftot + hrf[CH3Br];      /* RF_CH3Br */
  Ftot = Ftot + hrf[CH3CCl3];    /* RF_CH3CCl3 */
  Ftot = Ftot + hrf[CH3Cl];      /* RF_CH3Cl */
  Ftot = Ftot + fch4;            /* RF_CH4 */
  Ftot = Ftot + fco2;            /* RF_CO2 */
  Ftot = Ftot + fh2o_strat;      /* RF_H2O_strat */
  Ftot = Ftot + hrf[HCFC141b];   /* RF_HCFC141b */
  Ftot = Ftot + hrf[HCFC142b];   /* RF_HCFC142b */
  Ftot = Ftot + hrf[HCFC22];     /* RF_HCFC22 */
  Ftot = Ftot + hrf[HFC125];     /* RF_HFC125 */
  Ftot = Ftot + hrf[HFC134a];    /* RF_HFC134a */
  Ftot = Ftot + hrf[HFC143a];    /* RF_HFC143a */
  Ftot = Ftot + hrf[HFC227ea];   /* RF_HFC227ea */
  Ftot = Ftot + hrf[HFC23];      /* RF_HFC23 */
  Ftot = Ftot + hrf[HFC245fa];   /* RF_HFC245fa */
  Ftot = Ftot + hrf[HFC32];      /* RF_HFC32 */
  Ftot = Ftot + hrf[HFC4310];    /* RF_HFC4310 */
  Ftot = Ftot + fn2o;            /* RF_N2O */
  Ftot = Ftot + fnh3;            /* RF_NH3 */
  Ftot = Ftot + fo3_trop;        /* RF_O3_trop */
  Ftot = Ftot + foc;             /* RF_OC */
  Ftot = Ftot + hrf[SF6];        /* RF_SF6 */
  Ftot = Ftot + fso2;            /* RF_SO2 */
  Ftot = Ftot + aci_rf;          /* RF_aci */
  Ftot = Ftot + falbedo;         /* RF_albedo */
  Ftot = Ftot + hrf[halon1211];  /* RF_halon1211 */
  Ftot = Ftot + hrf[halon1301];  /* RF_halon1301 */
  Ftot = Ftot + hrf[halon2402];  /* RF_halon2402 */
  Ftot = Ftot + fmisc;           /* RF_misc */
  Ftot = Ftot + fvol;            /* RF_vol */
  *fco2_out = fco2; *fch4_out = fch4; *fn2o_out = fn2o;
  return Ftot;
}

/* ---------------------------------------------------------------------------------- */
void ho_default_params(ho_params *p) {
  memset(p, 0, sizeof *p);
  p->start_year = 1745; p->end_year = 2300; p->do_spinup = 1; p->max_spinup = 2000;
  p->S = 3.0; p->diff = 1.042; p->qco2 = 3.75;
  p->beta = 0.65; p->q10_rh = 1.2; p->f_nppv = 0.35; p->f_nppd = 0.60; p->f_litterd = 0.98;
  p->npp_flux0 = 56.2; p->C0 = 277.15;
  p->veg_c = 550; p->detritus_c = 55; p->soil_c = 917; p->permafrost_c = 865;
  p->warmingfactor = 1.0; p->rh_ch4_frac = 0.023; p->pf_mu = 1.67; p->pf_sigma = 0.986;
  p->fpf_static = 0.74;
  p->tt = 72000000; p->tu = 49000000; p->twi = 12500000; p->tid = 200000000;
  p->preind_C_surface = 900; p->preind_C_ID = 37100; p->spinup_chem = 0;
  p->eps_abs = 1.0e-6; p->eps_rel = 1.0e-6; p->dt = 0.25; p->eps_spinup = 0.001;
  p->baseyear = 1750; p->aero_scalar = 1.0; p->vol_scalar = 1.0;
  p->delta_co2 = 0.05; p->delta_ch4 = -.14; p->delta_n2o = 0.07;
  p->rho_bc = 0.06386286; p->rho_oc = -0.006407143; p->rho_so2 = -7.469841e-06;
  p->rho_nh3 = -0.002146032;
  p->M0 = 731.41; p->Tsoil = 120; p->Tstrat = 150; p->UC_CH4 = 2.78;
  p->TOH0 = 9.6; p->CNOX = 8.4e-3; p->CCO = -1.575e-4; p->CNMVOC = -4.725e-4; p->CCH4 = -0.32;
  p->PO3 = 30.0;
  p->N0 = 273.87; p->UC_N2O = 4.8; p->TN2O0 = 132;
  static const double tau[HO_NHALO] = {50000.0, 10000.0, 228.0, 5.4, 17.0, 30.0, 14.0, 51.0, 36.0,
                                       7.9, 3200.0, 52.0, 102.0, 93.0, 189, 540, 32.0, 5.0, 11.9,
                                       9.4, 18.0, 16.0, 72.0, 28.0, 0.9, 0.8};
  static const double rho[HO_NHALO] = {0.000099, 0.000261, 0.000191, 0.000111, 0.000357, 0.000234,
                                       0.000167, 0.000168, 0.000273, 0.000245, 0.000567, 0.000259,
                                       0.00032, 0.000301, 0.000314, 0.000246, 0.000166, 0.000065,
                                       0.000214, 0.000161, 0.000193, 0.00003, 0.000299, 0.000312,
                                       0.000005, 0.000004};
  static const double mm[HO_NHALO] = {88.0043, 138.01, 70.0, 52.0, 252.0, 120.02, 102.02, 84.04,
                                      170.03, 134.0, 146.06, 137.35, 120.9, 187.35, 170.9, 154.45,
                                      153.8, 133.35, 86.45, 116.9, 100.45, 165.35, 148.9, 259.8,
                                      50.45, 50.45};
  for (int g = 0; g < HO_NHALO; ++g) {
    p->halo_tau[g] = tau[g]; p->halo_rho[g] = rho[g]; p->halo_molarMass[g] = mm[g];
    p->halo_delta[g] = 0.0; p->halo_H0[g] = 0.0;
  }
  p->halo_delta[11] = 0.13; /* CFC11 */
  p->halo_delta[12] = 0.13; /* CFC12 */
  p->halo_H0[0] = 35.0;     /* CF4 */
  p->halo_H0[24] = 504.0;   /* CH3Cl */
  p->halo_H0[25] = 5.8;     /* CH3Br */
}

/* test hook for the fluxpool KATs: (a, fa, mask_a) + (b, fb, mask_b) -> fa, mask_a; returns the
 * member status (HO_ERR_TRACKING if the private-constructor checks fail) */
int ho_tm_add(double a, double *fa, uint32_t *mask_a, double b, const double *fb, uint32_t mask_b) {
  member_t *m = (member_t *)calloc(1, sizeof(member_t));
  tmap_t A, B;
  memcpy(A.f, fa, sizeof A.f); A.mask = *mask_a;
  memcpy(B.f, fb, sizeof B.f); B.mask = mask_b;
  int status = HO_OK;
  if (setjmp(m->fail)) status = m->status;
  else tm_add(m, a, &A, b, &B);
  memcpy(fa, A.f, sizeof A.f);
  *mask_a = A.mask;
  free(m);
  return status;
}

static double *dalloc(int n) { return (double *)calloc((size_t)n, sizeof(double)); }

int ho_run_member(const ho_params *p, const double *raw, int run_to, double *out, int nyears_cap,
                  int *fail_year, ho_counters *counters, ho_spinup_state *spin) {
  return ho_run_member_ex(p, raw, NULL, run_to, out, nyears_cap, fail_year, counters, spin, 9999,
                          NULL, NULL);
}

/* per-biome outputs of the next run on this thread (ho_run_member_biomes) */
static __thread double *g_bio_out = NULL;

int ho_run_member_biomes(const ho_params *p, const double *raw, const ho_constraints *cn, int run_to,
                         double *out, int nyears_cap, int *fail_year, double *bio_out) {
  g_bio_out = bio_out;
  const int st = ho_run_member_ex(p, raw, cn, run_to, out, nyears_cap, fail_year, NULL, NULL, 9999,
                                  NULL, NULL);
  g_bio_out = NULL;
  return st;
}

int ho_run_member_tracked(const ho_params *p, const double *raw, int run_to, double *out,
                          int nyears_cap, int *fail_year, ho_counters *counters,
                          ho_spinup_state *spin, int tracking_date, double *track_frac,
                          uint32_t *track_mask) {
  return ho_run_member_ex(p, raw, NULL, run_to, out, nyears_cap, fail_year, counters, spin,
                          tracking_date, track_frac, track_mask);
}

int ho_run_member_ex(const ho_params *p, const double *raw, const ho_constraints *cn, int run_to,
                     double *out, int nyears_cap, int *fail_year, ho_counters *counters,
                     ho_spinup_state *spin, int tracking_date, double *track_frac,
                     uint32_t *track_mask) {
  member_t *m = (member_t *)calloc(1, sizeof(member_t));
  m->cn = cn;
  m->tracking_date = tracking_date; /* core.cpp:60: default 9999 = never */
  if (track_frac)
    for (size_t i = 0; i < (size_t)nyears_cap * HO_NPOOL * HO_NSRC; ++i) track_frac[i] = NAN;
  if (track_mask) memset(track_mask, 0, (size_t)nyears_cap * HO_NPOOL * sizeof(uint32_t));
  const int nrow = p->end_year - p->start_year + 1;
  m->p = p; m->raw = raw; m->nrow = nrow;
  m->Tland_record = dalloc(nrow + 1);
  m->CH4 = dalloc(nrow); m->O3 = dalloc(nrow); m->N2O = dalloc(nrow);
  m->halo_rf = dalloc(nrow * HO_NHALO);
  m->rf_tot = dalloc(nrow); m->rf_co2 = dalloc(nrow); m->rf_ch4 = dalloc(nrow);
  m->rf_n2o = dalloc(nrow); m->co2_ts = dalloc(nrow);
  m->Ker = dalloc(nrow); m->temp = dalloc(nrow); m->temp_landair = dalloc(nrow);
  m->temp_sst = dalloc(nrow); m->heatflux_mixed = dalloc(nrow); m->heatflux_interior = dalloc(nrow);
  m->heat_mixed = dalloc(nrow); m->heat_interior = dalloc(nrow); m->forcing = dalloc(nrow);
  if (out)
    for (int i = 0; i < HO_NOUT * nyears_cap; ++i) out[i] = NAN;
  volatile int cur_year = p->start_year;
  int status = HO_OK;

  if (setjmp(m->fail)) {
    status = m->status;
    if (fail_year) *fail_year = cur_year;
    goto done;
  }

  /* ---- prepareToRun of every component ---- */
  m->N0_n2o = ho_gas_series_constrained(p, raw, cn, m->N2O, m->halo_rf);
  /* OHComponent::prepareToRun reads CH4's M0 first (dependency order), THEN
   * CH4Component::prepareToRun overwrites its M0 with a start-date constraint
   * (oh_component.cpp:131, ch4_component.cpp:137-147) */
  m->M0_ch4 = p->M0;
  if (cn && cn->ch4 && !isnan(cn->ch4[0])) m->M0_ch4 = cn->ch4[0];
  m->CH4[0] = m->M0_ch4;
  m->O3[0] = p->PO3;     /* o3_component.cpp:118-123 */
  m->tau_oh = p->TOH0;
  ocean_prepareToRun(m);
  /* SimpleNbox: simpleNbox.cpp:45-81, simpleNbox-runtime.cpp:61-197 */
  m->nb = p->n_biomes > 1 ? p->n_biomes : 1;
  if (m->nb > HO_MAX_BIOMES || (m->nb > 1 && tracking_date < 9999)) { /* not restated */
    status = HO_ERR_UNSUPPORTED;
    goto done;
  }
  for (int ib = 0; ib < m->nb; ++ib) {
    bio_t *b = &m->bio[ib];
    if (p->n_biomes > 1) {
      b->par = p->biome[ib];
      m->border[ib] = p->biome_order[ib];
    } else { /* the "global" biome of the scalar fields */
      ho_biome g = {p->veg_c, p->detritus_c, p->soil_c, p->permafrost_c, p->npp_flux0, p->beta,
                    p->q10_rh, p->warmingfactor, p->f_nppv, p->f_nppd, p->f_litterd,
                    p->rh_ch4_frac, p->pf_mu, p->pf_sigma, p->fpf_static};
      b->par = g;
      m->border[ib] = ib;
    }
    b->veg_c = FP(m, b->par.veg_c);
    b->detritus_c = FP(m, b->par.detritus_c);
    b->soil_c = FP(m, b->par.soil_c);
    b->permafrost_c = FP(m, b->par.permafrost_c);
    b->thawed_permafrost_c = 0.0;
    b->co2fert = b->tempfertd = b->tempferts = b->f_frozen = 1.0;
    b->f_new_thaw = 0.0;
    b->RH_ch4 = 0.0;
    b->final_npp = b->final_rh = 0.0;
  }
  m->earth_c = 5500;
  m->cum_luc_va = 0.0;
  m->npp_luc_adjust = 1.0;
  m->end_of_spinup_vegc = SUM_MAP(m, veg_c);
  m->cumulative_pf_ch4 = 0.0;
  m->has_been_run_before = 0;
  m->atmos_c = FP(m, p->C0 * PPMVCO2_TO_PGC);
  for (int i = 0; i < 7; ++i) tm_init(&m->tm[i], i); /* every pool starts as {own name: 1} */
  m->atmosphere_cpool = m->tm[TP_ATMOS];            /* simpleNbox-runtime.cpp:195 */
  m->atmosphere_cpool_tracking = 0;
  m->masstot = 0.0;
  m->t = p->start_year; /* solver prepareToRun, carbon-cycle-solver.cpp:126 */
  m->dt = p->dt;
  doeclim_prepareToRun(m);

  /* ---- spin-up: core.cpp:394-420, carbon-cycle-solver.cpp:313-370 ---- */
  if (p->do_spinup) {
    m->in_spinup = 1;
    int spunup = 0, step = 0, solver_in_spinup = 0;
    while (!spunup && ++step < p->max_spinup) {
      /* ocean run_spinup -> run(step): CO2 from atmos_c_ts (flat), SST = 0 */
      ocean_run(m, (double)step, snbox_CO2_conc(m), m->sst);
      if (!solver_in_spinup) {
        solver_in_spinup = 1;
        m->t = step - 1;
      }
      double c_old[NC], c_new[NC];
      snbox_getCValues(m, m->t, c_old);
      solver_run(m, step, m->tas_land);
      snbox_getCValues(m, step, c_new);
      double max_dcdt = 0.0;
      for (int i = 0; i < NC; i++) {
        double d = fabs(c_new[i] - c_old[i]);
        if (d > max_dcdt) max_dcdt = d;
      }
      spunup = (max_dcdt < p->eps_spinup);
      if (spunup) m->t = p->start_year;
    }
    m->cnt.spinup_steps = (uint64_t)step;
    m->in_spinup = 0;
    m->t = p->start_year; /* core.cpp: even if not spun up the run starts at startDate */
  }
  m->have_tempferts_last = 0; /* tempferts_tv[t] is only consulted for t > startDate */
  if (spin) {
    spin->atmos = m->atmos_c; spin->veg = SUM_MAP(m, veg_c); spin->det = SUM_MAP(m, detritus_c);
    spin->soil = SUM_MAP(m, soil_c); spin->permafrost = SUM_MAP(m, permafrost_c);
    spin->thawed = SUM_MAP(m, thawed_permafrost_c); spin->earth = m->earth_c;
    for (int i = 0; i < 4; ++i) spin->ocean[i] = m->box[i].carbon;
    spin->spinup_steps = (int)m->cnt.spinup_steps;
    spin->alk_HL = spin->alk_LL = 0;
  }
  m->co2_ts[0] = m->atmos_c;

  /* ---- main loop: core.cpp:483-504; component order SURVEY.md section 1 ---- */
  if (run_to < 0 || run_to > p->end_year) run_to = p->end_year;
  for (int y = p->start_year + 1; y <= run_to; ++y) {
    cur_year = y;
    const int r = y - p->start_year;
    const double *row = raw + (size_t)r * HO_NRAW;
    const double *row0 = raw;

    /* OH: oh_component.cpp:137-174 */
    {
      const double previous_ch4 = m->CH4[r - 1];
      double toh = 0.0;
      if (previous_ch4 != p->M0) {
        const double a = p->CCH4 * ((1.0 * log(previous_ch4)) - log(p->M0));
        const double b = p->CNOX * ((1.0 * row[HO_RAW_NOX]) - row0[HO_RAW_NOX]);
        const double c = p->CCO * ((1.0 * +row[HO_RAW_CO]) - row0[HO_RAW_CO]);
        const double d = p->CNMVOC * ((1.0 * +row[HO_RAW_NMVOC]) - row0[HO_RAW_NMVOC]);
        toh = a + b + c + d;
      }
      m->tau_oh = p->TOH0 * exp(-toh);
    }
    /* CH4: ch4_component.cpp:152-199 */
    double ch4_c;
    if (cn_get(m, cn ? cn->ch4 : NULL, y, &ch4_c)) {
      m->CH4[r] = ch4_c;
    } else {
      const double current_ch4em = row[HO_RAW_CH4_E];
      const double current_toh = m->tau_oh;
      const double rh_ch4 = SUM_MAP(m, RH_ch4) * (1000.0 * 16.04 / 12.01);
      const double ch4n = row[HO_RAW_CH4N];
      const double emisTocon = (current_ch4em + rh_ch4 + ch4n) / p->UC_CH4;
      const double previous_ch4 = m->CH4[r - 1];
      const double soil_sink = previous_ch4 / p->Tsoil;
      const double strat_sink = previous_ch4 / p->Tstrat;
      const double oh_sink = previous_ch4 / current_toh;
      const double dCH4 = emisTocon - soil_sink - strat_sink - oh_sink;
      m->CH4[r] = previous_ch4 + dCH4;
    }
    /* O3: o3_component.cpp:126-146 */
    m->O3[r] = (5 * log(m->CH4[r])) + (0.125 * row[HO_RAW_NOX]) + (0.0011 * row[HO_RAW_CO]) +
               (0.0033 * row[HO_RAW_NMVOC]);

    /* ocean: CO2 = atmos_c_ts.get(y) = flat extrapolation of year y-1; SST current */
    ocean_run(m, (double)y, FP(m, m->co2_ts[r - 1] * PGC_TO_PPMVCO2), m->sst);
    if (spin && y == p->start_year + 1) {
      spin->alk_HL = m->box[HL].chem.alk;
      spin->alk_LL = m->box[LL].chem.alk;
    }
    /* SimpleNbox::run: simpleNbox-runtime.cpp:206-227 */
    if (!m->has_been_run_before) {
      m->end_of_spinup_vegc = SUM_MAP(m, veg_c);
      m->has_been_run_before = 1;
    }
    /* tracking start (:215-220), then tell the ocean what the atmosphere is made of (:225) */
    if ((double)y == (double)m->tracking_date) m->tracking = 1;
    m->Tland_record[r] = m->tas_land;
    m->atmosphere_cpool = m->tm[TP_ATMOS];
    m->atmosphere_cpool_tracking = m->tracking;
    /* solver */
    solver_run(m, (double)y, m->tas_land);
    m->co2_ts[r] = m->atmos_c;
    const double CO2_conc = FP(m, m->atmos_c * PGC_TO_PPMVCO2);

    /* forcing */
    double rf_tot_rel = 0.0, rf_co2_rel = 0.0, rf_ch4_rel = 0.0, rf_n2o_rel = 0.0;
    if (!((double)y < p->baseyear)) {
      double fco2, fch4, fn2o;
      double Ftot = forcing_total(m, r, CO2_conc, m->CH4[r], m->N2O[r], m->O3[r], &fco2, &fch4, &fn2o);
      /* user-supplied total forcing (forcing_component.cpp:498-505): every year up to the
       * series' last date, Ftot_constrain.get() interpolating / extrapolating flat */
      if (cn && cn->rf_tot && r <= cn->rf_tot_last_row) Ftot = cn->rf_tot[r];
      if ((double)y == p->baseyear) {
        m->base_tot = Ftot; m->base_co2 = fco2; m->base_ch4 = fch4; m->base_n2o = fn2o;
      }
      rf_tot_rel = Ftot - m->base_tot;
      rf_co2_rel = fco2 - m->base_co2;
      rf_ch4_rel = fch4 - m->base_ch4;
      rf_n2o_rel = fn2o - m->base_n2o;
    }
    m->rf_tot[r] = rf_tot_rel;
    /* temperature */
    doeclim_run(m, r, rf_tot_rel);

    if (out && r - 1 < nyears_cap) {
      const int i = r - 1;
#define OUT(k, v) out[(size_t)(k) * nyears_cap + i] = (v)
      if (g_bio_out) /* getData("<biome>.<name>"), simpleNbox.cpp:533-697 */
        for (int ib = 0; ib < m->nb; ++ib) {
          const bio_t *b = &m->bio[ib];
          const double v[HO_NBIOME_OUT] = {b->veg_c, b->detritus_c, b->soil_c, b->permafrost_c,
                                           b->thawed_permafrost_c, b->final_npp, b->final_rh};
          for (int k = 0; k < HO_NBIOME_OUT; ++k)
            g_bio_out[((size_t)ib * HO_NBIOME_OUT + k) * nyears_cap + i] = v[k];
        }
      OUT(HO_OUT_CO2, CO2_conc);
      OUT(HO_OUT_TAS, m->tas);
      OUT(HO_OUT_RF_TOT, rf_tot_rel);
      OUT(HO_OUT_RF_CO2, rf_co2_rel);
      OUT(HO_OUT_HEATFLUX, m->heatflux);
      OUT(HO_OUT_OCEAN_C, m->box[DO].carbon + m->box[IO].carbon + m->box[LL].carbon + m->box[HL].carbon);
      OUT(HO_OUT_HL_PH, m->box[HL].chem.pH);
      OUT(HO_OUT_ATMOS_C, m->atmos_c);
      OUT(HO_OUT_SST, m->sst);
      OUT(HO_OUT_PERMAFROST_C, SUM_MAP(m, permafrost_c));
      OUT(HO_OUT_CH4, m->CH4[r]);
      OUT(HO_OUT_N2O, m->N2O[r]);
      OUT(HO_OUT_O3, m->O3[r]);
      OUT(HO_OUT_LAND_TAS, m->tas_land);
      OUT(HO_OUT_VEG_C, SUM_MAP(m, veg_c));
      OUT(HO_OUT_DETRITUS_C, SUM_MAP(m, detritus_c));
      OUT(HO_OUT_SOIL_C, SUM_MAP(m, soil_c));
      OUT(HO_OUT_THAWEDP_C, SUM_MAP(m, thawed_permafrost_c));
      OUT(HO_OUT_EARTH_C, m->earth_c);
      OUT(HO_OUT_NBP, m->nbp);
      OUT(HO_OUT_OCEAN_UPTAKE, m->annualflux_sum);
      OUT(HO_OUT_LL_PH, m->box[LL].chem.pH);
      OUT(HO_OUT_PCO2_HL, m->box[HL].chem.PCO2o);
      OUT(HO_OUT_PCO2_LL, m->box[LL].chem.PCO2o);
      OUT(HO_OUT_CARBON_HL, m->box[HL].carbon);
      OUT(HO_OUT_CARBON_LL, m->box[LL].carbon);
      OUT(HO_OUT_CARBON_IO, m->box[IO].carbon);
      OUT(HO_OUT_CARBON_DO, m->box[DO].carbon);
      OUT(HO_OUT_RF_CH4, rf_ch4_rel);
      OUT(HO_OUT_RF_N2O, rf_n2o_rel);
      OUT(HO_OUT_RH_CH4, SUM_MAP(m, RH_ch4));
      OUT(HO_OUT_NPP, SUM_MAP(m, final_npp));
      OUT(HO_OUT_RH, SUM_MAP(m, final_rh));
      OUT(HO_OUT_GMST, dc_flnd * m->temp_landair[r] + (1.0 - dc_flnd) * m->temp_sst[r]);
      OUT(HO_OUT_OCEAN_TAS, m->tas_ocean);
      OUT(HO_OUT_FLUX_MIXED, m->heatflux_mixed[r]);
      OUT(HO_OUT_FLUX_INTERIOR, m->heatflux_interior[r]);
      OUT(HO_OUT_TIMESTEPS, (double)m->timesteps);
#undef OUT
    }
    /* what the CSVFluxPoolVisitor would print for this year (csv_tracking_visitor.cpp:80-137) */
    if (m->tracking && r - 1 < nyears_cap) {
      const tmap_t *maps[HO_NPOOL] = {&m->tm[0], &m->tm[1], &m->tm[2], &m->tm[3], &m->tm[4],
                                      &m->tm[5], &m->tm[6], &m->box[HL].cmap, &m->box[LL].cmap,
                                      &m->box[IO].cmap, &m->box[DO].cmap};
      for (int k = 0; k < HO_NPOOL; ++k) {
        if (track_frac)
          memcpy(track_frac + ((size_t)(r - 1) * HO_NPOOL + k) * HO_NSRC, maps[k]->f,
                 sizeof(double) * HO_NSRC);
        if (track_mask) track_mask[(size_t)(r - 1) * HO_NPOOL + k] = maps[k]->mask;
      }
    }
  }

done:
  if (counters) *counters = m->cnt;
  free(m->Tland_record); free(m->CH4); free(m->O3); free(m->N2O); free(m->halo_rf);
  free(m->rf_tot); free(m->rf_co2); free(m->rf_ch4); free(m->rf_n2o); free(m->co2_ts);
  free(m->Ker); free(m->temp); free(m->temp_landair); free(m->temp_sst);
  free(m->heatflux_mixed); free(m->heatflux_interior); free(m->heat_mixed);
  free(m->heat_interior); free(m->forcing);
  free(m);
  return status;
}
