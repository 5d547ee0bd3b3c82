"""tests/golden/make_golden.py -- regenerates the committed fixtures.  Runs ONLY in the build
container (needs /root/reference and oracle/_ref/libhector_ref.so); the fixtures it writes are
what travels to the GPU box.

  scenarios.npz       dense raw scenario tables [556][44] for the 8 shipped SSP inis, read from
                      /root/reference/inst/input/*.ini + tables/*.csv with the reference's
                      tseries semantics (exact key / linear interpolation / single value).
  hector_comp.npz     the reference's own golden trajectories
                      (/root/reference/tests/testthat/compdata/hector_comp.csv, 10 vars x 556 yr).
  ref_runs.npz        known-answer trajectories produced by the UNMODIFIED reference
                      (oracle/_ref) for a set of parameter vectors / scenarios, incl. the
                      expected-failure member.
  ref_tracking.npz    carbon-tracking known answers from the UNMODIFIED reference: source
                      fractions + key masks of the 11 tracked pools at selected years (read at
                      full precision through oracle/ref_driver.cpp), and the reference's own
                      6-digit getTrackingData() CSV for one short run.
  ref_outputs_extra.npz  NPP, RH, gmst, ocean_tas, heatflux_mixed, heatflux_interior of three runs
                      of the UNMODIFIED reference (outputs added after ref_runs.npz was frozen).
  ref_constraints.npz known answers of the UNMODIFIED reference under user constraints (CO2, NBP,
                      tas, RF_tot, CH4, N2O, halocarbon concentrations), incl. two expected
                      failures.
"""
import csv
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

HALOS = ["CF4", "C2F6", "HFC23", "HFC32", "HFC4310", "HFC125", "HFC134a", "HFC143a", "HFC227ea",
         "HFC245fa", "SF6", "CFC11", "CFC12", "CFC113", "CFC114", "CFC115", "CCl4", "CH3CCl3",
         "HCFC22", "HCFC141b", "HCFC142b", "halon1211", "halon1301", "halon2402", "CH3Cl",
         "CH3Br"]
# (section, variable) in HO_RAW_* order (oracle/hector_oracle.h)
RAW = [("simpleNbox", "ffi_emissions"), ("simpleNbox", "daccs_uptake"),
       ("simpleNbox", "luc_emissions"), ("simpleNbox", "luc_uptake"),
       ("CH4", "CH4_emissions"), ("CH4", "CH4N"),
       ("OH", "NOX_emissions"), ("OH", "CO_emissions"), ("OH", "NMVOC_emissions"),
       ("bc", "BC_emissions"), ("oc", "OC_emissions"), ("so2", "SO2_emissions"),
       ("nh3", "NH3_emissions"), ("so2", "SV"), ("simpleNbox", "RF_albedo"),
       ("forcing", "RF_misc"), ("N2O", "N2O_emissions"), ("N2O", "N2O_natural_emissions")]
RAW += [("%s_halocarbon" % h, "%s_emissions" % h) for h in HALOS]
RAW_NAMES = [v for _, v in RAW]
SCENARIOS = ["ssp119", "ssp126", "ssp245", "ssp370", "ssp434", "ssp460", "ssp534-over", "ssp585"]


def parse_ini(path):
    """{section: {key: value-string}} with ';' comments stripped (inih semantics, src/ini.c)."""
    out, sec = {}, None
    for line in open(path):
        line = line.split(";")[0].strip()
        if not line:
            continue
        if line.startswith("["):
            sec = line[1:line.index("]")]
            out.setdefault(sec, {})
        elif "=" in line:
            k, v = line.split("=", 1)
            out[sec][k.strip()] = v.strip()
    return out


def read_csv_column(path, var):
    """csv_table_reader.cpp:115-196 -> {date: value}"""
    rows = {}
    header = None
    for line in open(path, newline=""):
        line = line.rstrip("\r\n")
        if not line or line.startswith(";"):
            continue
        cells = [c.strip() for c in line.split(",")]
        if header is None:
            header = cells
            col = header.index(var, 1)
            continue
        if cells[0] == "UNITS" or cells[col] == "":
            continue
        rows[float(cells[0])] = float(cells[col])
    return rows


def tseries_get(series, t):
    """tseries.hpp:317-334 + h_interpolator.cpp:109-125 (linear, flat ends)"""
    if len(series) == 1:
        return next(iter(series.values()))
    if t in series:
        return series[t]
    xs = sorted(series)
    if t < xs[0]:
        return series[xs[0]]
    if t >= xs[-1]:
        return series[xs[-1]]
    i = max(k for k, x in enumerate(xs) if x <= t)
    x0, x1 = xs[i], xs[i + 1]
    y0, y1 = series[x0], series[x1]
    return y0 + (t - x0) * (y1 - y0) / (x1 - x0)


def scenario_table(scn, ini_path=None):
    ini_path = ini_path or os.path.join(REF, "inst/input/hector_%s.ini" % scn)
    ini = parse_ini(ini_path)
    start, end = int(ini["core"]["startDate"]), int(ini["core"]["endDate"])
    tab = np.zeros((end - start + 1, len(RAW)))
    for j, (sec, var) in enumerate(RAW):
        series = {}
        for k, v in ini[sec].items():
            if k == var and v.startswith("csv:"):
                series = read_csv_column(os.path.join(os.path.dirname(ini_path), v[4:]), var)
            elif k == var:
                series = {0.0: float(v)}
            elif k.startswith(var + "["):
                series[float(k[len(var) + 1:-1])] = float(v)
        assert series, (sec, var)
        for r in range(end - start + 1):
            tab[r, j] = tseries_get(series, float(start + r))
    return tab, ini


REF_VARS = ["CO2_concentration", "global_tas", "RF_tot", "RF_CO2", "heatflux", "ocean_c", "HL_pH",
            "atmos_co2", "sst", "permafrost_c", "CH4_concentration", "N2O_concentration",
            "O3_concentration", "land_tas", "veg_c", "detritus_c", "soil_c", "thawedp_c",
            "earth_c", "NBP", "ocean_uptake", "LL_pH", "HL_PCO2", "LL_PCO2", "HL_ocean_c",
            "LL_ocean_c", "IO_ocean_c", "DO_ocean_c", "RF_CH4", "RF_N2O", "rh_ch4"]


def lhs(M, seed=20241017):
    """SURVEY.md section 8(d) sampler: (S, q10_rh, beta, diff) Latin hypercube."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo = np.array([2.0, 1.0, 0.2, 0.5])
    hi = np.array([5.0, 2.6, 0.9, 2.5])
    cols = []
    for j in range(4):
        u = (rng.permutation(M) + rng.random(M)) / M
        cols.append(lo[j] + u * (hi[j] - lo[j]))
    return np.stack(cols, axis=1)


def main():
    from oracle import ref
    # 1. scenarios
    tabs = {}
    for s in SCENARIOS:
        tabs[s], _ = scenario_table(s)
    np.savez_compressed(os.path.join(OUT, "scenarios.npz"), names=np.array(RAW_NAMES),
                        **{s.replace("-", "_"): t for s, t in tabs.items()})
    # 2. the reference's own golden file
    gold = {}
    for r in csv.DictReader(open(os.path.join(REF, "tests/testthat/compdata/hector_comp.csv"))):
        gold.setdefault(r["variable"], {})[int(r["year"])] = float(r["value"])
    gv = sorted(gold)
    np.savez_compressed(os.path.join(OUT, "hector_comp.npz"), variables=np.array(gv),
                        years=np.arange(1745, 2301),
                        values=np.array([[gold[v][y] for y in range(1745, 2301)] for v in gv]))
    # 3. reference KAT runs
    cases = []  # (name, scenario, params)
    cases.append(("default_ssp245", "ssp245", {}))
    cases.append(("default_ssp585", "ssp585", {}))
    cases.append(("default_ssp119", "ssp119", {}))
    cases.append(("default_ssp534-over", "ssp534-over", {}))
    cases.append(("corner_lo_ssp245", "ssp245", dict(S=1.5, q10_rh=1.0, beta=0.1, diff=0.3)))
    cases.append(("corner_hi_ssp245", "ssp245", dict(S=6.0, q10_rh=3.5, beta=1.0, diff=3.0)))
    cases.append(("fail_ssp585", "ssp585", dict(S=8.0, q10_rh=4.0, beta=0.05, diff=0.2)))
    cases.append(("aero_vol_ssp370", "ssp370", dict(aero_scalar=1.4, vol_scalar=0.85, S=3.7)))
    cases.append(("nppv_ssp245", "ssp245", dict(f_nppv=0.3)))
    X = lhs(16)
    for i in range(16):
        cases.append(("lhs16_%02d" % i, "ssp245",
                      dict(S=X[i, 0], q10_rh=X[i, 1], beta=X[i, 2], diff=X[i, 3])))
    # full variable list for a few cases, a short list for the rest (keeps the fixture small);
    # rows that were not requested are NaN
    FULL = {"default_ssp245", "default_ssp585", "corner_hi_ssp245", "aero_vol_ssp370"}
    SHORT = ["CO2_concentration", "global_tas", "RF_tot", "ocean_c", "HL_pH", "permafrost_c",
             "CH4_concentration", "rh_ch4"]
    TINY = ["CO2_concentration", "global_tas"]
    names, scns, pnames, pvals, outs, oks, errs = [], [], [], [], [], [], []
    for name, scn, params in cases:
        want = REF_VARS if name in FULL else (TINY if name.startswith("lhs") else SHORT)
        ok, err, o, _ = ref.run_member(os.path.join(REF, "inst/input/hector_%s.ini" % scn),
                                       params, want)
        out = np.full((len(REF_VARS) + 1, o.shape[1]), np.nan)
        for k, v in enumerate(want):
            out[REF_VARS.index(v)] = o[k]
        out[-1] = o[-1]
        print(name, ok, err[:80])
        names.append(name); scns.append(scn); oks.append(ok); errs.append(err)
        pnames.append(",".join(params.keys()))
        pvals.append(np.array(list(params.values()) + [np.nan] * (8 - len(params))))
        outs.append(out)
    np.savez_compressed(os.path.join(OUT, "ref_runs.npz"), names=np.array(names),
                        scenarios=np.array(scns), param_names=np.array(pnames),
                        param_values=np.array(pvals), ok=np.array(oks), errors=np.array(errs),
                        variables=np.array(REF_VARS + ["ocean_timesteps"]),
                        values=np.array(outs))


TRACK_CASES = [  # (name, scenario, trackingDate, params)
    ("ssp245_1750", "ssp245", 1750, {}),
    ("ssp585_1850_pert", "ssp585", 1850, dict(S=4.2, q10_rh=2.0, beta=0.4, diff=1.8)),
    ("ssp119_2000", "ssp119", 2000, {}),
    ("ssp245_1750_corner_hi", "ssp245", 1750, dict(S=6.0, q10_rh=3.5, beta=1.0, diff=3.0)),
]


def tracking_years(tdate):
    ys = set(range(tdate, tdate + 4)) | set(range(tdate + 10 - tdate % 10, 2301, 10)) | {2300}
    return sorted(y for y in ys if y <= 2300)


def make_tracking():
    from oracle import ref
    names, scns, tdates, pnames, pvals = [], [], [], [], []
    years, fracs, masks, values = [], [], [], []
    for name, scn, tdate, params in TRACK_CASES:
        c = ref.RefCore(os.path.join(REF, "inst/input/hector_%s.ini" % scn))
        c.setdata("core", "trackingDate", tdate)
        for k, v in params.items():
            c.setdata(ref.PARAM_COMPONENT[k], k, v)
        c.prepare()
        ys = tracking_years(tdate)
        Y = np.full(64, -1); F = np.full((64, 11, 12), np.nan); K = np.zeros((64, 11), np.uint32)
        V = np.full((64, 11), np.nan)
        for i, y in enumerate(ys):
            c.run(y)
            on, v, f, pres = c.tracking_state()
            assert on
            Y[i] = y; F[i] = f; V[i] = v
            K[i] = (pres.astype(np.uint32) << np.arange(12, dtype=np.uint32)).sum(1)
        c.close()
        print(name, len(ys), "years")
        names.append(name); scns.append(scn); tdates.append(tdate)
        pnames.append(",".join(params.keys()))
        pvals.append(np.array(list(params.values()) + [np.nan] * (8 - len(params))))
        years.append(Y); fracs.append(F); masks.append(K); values.append(V)
    # the reference's own output format, one short run
    c = ref.RefCore(os.path.join(REF, "inst/input/hector_ssp245.ini"))
    c.setdata("core", "trackingDate", 1750)
    c.prepare()
    c.run(1755)
    csv_text = c.tracking_csv()
    c.close()
    np.savez_compressed(os.path.join(OUT, "ref_tracking.npz"), names=np.array(names),
                        scenarios=np.array(scns), tracking_dates=np.array(tdates),
                        param_names=np.array(pnames), param_values=np.array(pvals),
                        years=np.array(years), frac=np.array(fracs), mask=np.array(masks),
                        pool_values=np.array(values), pools=np.array(ref.TRACK_POOLS),
                        sources=np.array(ref.TRACK_SOURCES), csv_1750_1755=np.array(csv_text))


CONSTRAINT_UNITS = {"CH4_constrain": "ppbv CH4", "N2O_constrain": "ppbv N2O",
                    "HFC23_constrain": "pptv", "CFC11_constrain": "pptv",
                    "RF_tot_constrain": "W/m2", "tas_constrain": "degC",
                    "CO2_constrain": "ppmv CO2", "NBP_constrain": "Pg C/yr"}
CONSTRAINT_VARS = ["CO2_concentration", "global_tas", "RF_tot", "CH4_concentration",
                   "N2O_concentration", "NBP", "ocean_c", "veg_c", "soil_c", "sst", "land_tas",
                   "heatflux", "RF_CH4", "RF_N2O", "DO_ocean_c", "thawedp_c"]


def constraint_cases():
    """name -> {constraint: {year: value}}; built from the emission-driven SSP2-4.5 run of the
    unmodified reference, like tests/testthat/test_constraints.R does"""
    from oracle import ref
    ok, err, o, _ = ref.run_member(os.path.join(REF, "inst/input/hector_ssp245.ini"), {},
                                   ["CO2_concentration", "CH4_concentration",
                                    "N2O_concentration", "NBP"])
    assert ok, err
    co2, ch4, n2o, nbp = o[0], o[1], o[2], o[3]
    yrs = range(1746, 2301)
    return {
        "ch4_all": {"CH4_constrain": dict([(1745, 731.41 * 1.2)] +
                                          [(y, ch4[y - 1746] * 1.2) for y in yrs])},
        "ch4_part": {"CH4_constrain": {y: ch4[y - 1746] * 0.9 for y in range(1900, 2051)}},
        "n2o_all": {"N2O_constrain": dict([(1745, 273.87 * 1.1)] +
                                          [(y, n2o[y - 1746] * 1.1) for y in yrs])},
        "n2o_part": {"N2O_constrain": {y: 300.0 + 0.1 * (y - 1950) for y in range(1950, 2101)}},
        "halo": {"HFC23_constrain": {y: 5.0 + 0.2 * (y - 1990) for y in range(1990, 2201)},
                 "CFC11_constrain": {y: 100.0 for y in range(1800, 1900)}},
        "rf_tot": {"RF_tot_constrain": {y: 0.01 * (y - 1850) for y in range(1850, 2001)}},
        "rf_tot_sparse": {"RF_tot_constrain": {1800: 0.1, 1900: 0.5, 2000: 2.0, 2050: 3.0}},
        "tas": {"tas_constrain": {y: 0.005 * (y - 1900) for y in range(1900, 2051)}},
        "tas_sparse": {"tas_constrain": {1850: 0.0, 1950: 0.4, 2100: 2.5}},
        "co2": {"CO2_constrain": {y: co2[y - 1746] * 1.02 for y in range(1746, 2101)}},
        "co2_part": {"CO2_constrain": {y: 300.0 + 1.0 * (y - 1950) for y in range(1950, 2021)}},
        "nbp_near": {"NBP_constrain": {y: nbp[y - 1746] + 0.2 for y in range(1950, 2051)}},
        "nbp_fail_mass": {"NBP_constrain": {y: 1.0 for y in range(1900, 2101)}},
        "nbp_fail_negative": {"NBP_constrain": {y: nbp[y - 1746] for y in range(1850, 2301)}},
        "combo": {"CO2_constrain": {y: co2[y - 1746] * 0.99 for y in range(1850, 2015)},
                  "tas_constrain": {y: 0.004 * (y - 1850) for y in range(1850, 2015)},
                  "CH4_constrain": {y: ch4[y - 1746] for y in range(1850, 2015)}},
    }


def make_constraints():
    """ref_constraints.npz: the unmodified reference run with user constraints set through
    sendMessage(M_SETDATA) before prepareToRun; NaN from the failing year on."""
    from oracle import ref
    cases = constraint_cases()
    names, specs, vals, fails = [], [], [], []
    for name, spec in cases.items():
        c = ref.RefCore(os.path.join(REF, "inst/input/hector_ssp245.ini"))
        for var, d in spec.items():
            for y, v in d.items():
                c.setvar(var, float(v), CONSTRAINT_UNITS[var], float(y))
        c.prepare()
        out = np.full((len(CONSTRAINT_VARS), 555), np.nan)
        fail = 0
        for y in range(1746, 2301):
            try:
                c.run(y)
            except ref.RefError as e:
                fail = y
                print("  ", name, "fails in", y, str(e)[:70])
                break
            for k, v in enumerate(CONSTRAINT_VARS):
                out[k, y - 1746] = c.fetch(v, y)
        c.close()
        print(name, "ok" if not fail else "failed")
        names.append(name); vals.append(out); fails.append(fail)
        # spec as a flat table: rows (constraint index, year, value)
        rows = []
        for var, d in spec.items():
            for y, v in d.items():
                rows.append((list(CONSTRAINT_UNITS).index(var), y, v))
        specs.append(np.array(rows, dtype=np.float64))
    width = max(len(r) for r in specs)
    S = np.full((len(specs), width, 3), np.nan)
    for i, r in enumerate(specs):
        S[i, :len(r)] = r
    np.savez_compressed(os.path.join(OUT, "ref_constraints.npz"), names=np.array(names),
                        constraint_names=np.array(list(CONSTRAINT_UNITS)), spec=S,
                        variables=np.array(CONSTRAINT_VARS), values=np.array(vals),
                        fail_year=np.array(fails))


EXTRA_VARS = ["NPP", "RH", "gmst", "ocean_tas", "heatflux_mixed", "heatflux_interior",
              # what a land-ocean warming ratio moves (the lo_ratio_* cases)
              "CO2_concentration", "global_tas", "land_tas", "sst", "veg_c", "soil_c", "HL_pH",
              "heatflux", "permafrost_c",
              # forcing agents and halocarbons the engine derives at fetch time
              "RF_BC", "RF_OC", "RF_SO2", "RF_NH3", "RF_aci", "RF_vol", "RF_albedo", "RF_misc",
              "RF_O3_trop", "RF_H2O_strat", "RF_CF4", "FadjCF4", "RF_CFC11", "FadjCFC11",
              "RF_CH3Br", "FadjHFC134a", "CF4_concentration", "CFC12_concentration"]
EXTRA_CASES = [("default_ssp245", "ssp245", {}),
               ("pert_ssp585", "ssp585", dict(S=4.2, q10_rh=2.0, beta=0.4, diff=1.8)),
               ("aero_vol_ssp370", "ssp370", dict(aero_scalar=1.4, vol_scalar=0.85, S=3.7)),
               ("corner_lo_ssp119", "ssp119", dict(S=1.5, q10_rh=1.0, beta=0.1, diff=0.3)),
               # user-provided land-ocean warming ratio (temperature_component.cpp:586-622, 722-739)
               ("lo_ratio_ssp245", "ssp245", dict(lo_warming_ratio=1.6)),
               ("lo_ratio_ssp585", "ssp585", dict(lo_warming_ratio=1.3, S=4.0, q10_rh=2.2)),
               ("lo_ratio_low_ssp126", "ssp126", dict(lo_warming_ratio=0.9, beta=0.2, diff=2.5))]


def make_extra_outputs():
    """ref_outputs_extra.npz: outputs added after ref_runs.npz was frozen (NPP, RH, gmst,
    ocean_tas, heatflux_mixed, heatflux_interior), from the UNMODIFIED reference."""
    from oracle import ref
    names, scns, pnames, pvals, vals = [], [], [], [], []
    for name, scn, params in EXTRA_CASES:
        ok, err, o, _ = ref.run_member(os.path.join(REF, "inst/input/hector_%s.ini" % scn), params,
                                       EXTRA_VARS)
        assert ok, err
        names.append(name); scns.append(scn); pnames.append(",".join(params))
        pvals.append(np.array(list(params.values()) + [np.nan] * (8 - len(params))))
        vals.append(o[:len(EXTRA_VARS)])
        print(name, "ok")
    np.savez_compressed(os.path.join(OUT, "ref_outputs_extra.npz"), names=np.array(names),
                        scenarios=np.array(scns), param_names=np.array(pnames),
                        param_values=np.array(pvals), variables=np.array(EXTRA_VARS),
                        values=np.array(vals))


# the rest of R's ALL_VARS() (tests/testthat/test_set_get_data.R "Can fetch all variables"): what
# the outputstream visitor prints beyond the variables above
MORE_VARS = ["HL_sst", "LL_sst", "HL_DIC", "LL_DIC", "DIC", "pH", "PCO2", "ML_ocean_c", "TAU_OH",
             "f_frozen", "HL_CO3", "LL_CO3", "CO3", "HL_ocean_uptake", "LL_ocean_uptake", "rh_det",
             "rh_soil",
             # what the engine derives them from
             "sst", "land_tas", "permafrost_c", "HL_ocean_c", "LL_ocean_c", "HL_pH", "LL_pH",
             "HL_PCO2", "LL_PCO2", "CH4_concentration"]


def make_more_outputs():
    """ref_outputs_more.npz: the remaining outputstream variables from the UNMODIFIED reference
    (three of EXTRA_CASES)"""
    from oracle import ref
    names, scns, pnames, pvals, vals = [], [], [], [], []
    for name, scn, params in [EXTRA_CASES[0], EXTRA_CASES[1], EXTRA_CASES[5]]:
        ok, err, o, _ = ref.run_member(os.path.join(REF, "inst/input/hector_%s.ini" % scn), params,
                                       MORE_VARS)
        assert ok, err
        names.append(name); scns.append(scn); pnames.append(",".join(params))
        pvals.append(np.array(list(params.values()) + [np.nan] * (8 - len(params))))
        vals.append(o[:len(MORE_VARS)])
        print(name, "ok")
    np.savez_compressed(os.path.join(OUT, "ref_outputs_more.npz"), names=np.array(names),
                        scenarios=np.array(scns), param_names=np.array(pnames),
                        param_values=np.array(pvals), variables=np.array(MORE_VARS),
                        values=np.array(vals))


BIOME_VARS = ["CO2_concentration", "global_tas", "veg_c", "detritus_c", "soil_c", "permafrost_c",
              "thawedp_c", "NPP", "RH", "NBP", "CH4_concentration", "RF_tot", "HL_pH", "ocean_c",
              "land_tas", "atmos_co2", "earth_c"]
BIOME_OWN_VARS = ["veg_c", "detritus_c", "soil_c", "permafrost_c", "thawedp_c", "NPP", "RH"]
BIOME_GLOBAL_KEYS = ["npp_flux0", "veg_c", "detritus_c", "soil_c", "permafrost_c", "f_nppv", "f_nppd",
                     "f_litterd", "beta", "q10_rh"]


def _split(fracs, **over):
    """biomes that split the default global pools and NPP by `fracs` (like R/biome.R split_biome
    does), with per-biome overrides"""
    glob = dict(npp_flux0=56.2, veg_c=550.0, detritus_c=55.0, soil_c=917.0, permafrost_c=865.0,
                f_nppv=0.35, f_nppd=0.60, f_litterd=0.98, beta=0.65, q10_rh=1.2)
    out = {}
    for name, fr in fracs.items():
        b = dict(glob)
        for k in ("npp_flux0", "veg_c", "detritus_c", "soil_c", "permafrost_c"):
            b[k] = glob[k] * fr
        b.update(over.get(name, {}))
        out[name] = b
    return out


# name -> (scenario, {biome: {field: value}} in creation order, {other parameter: value})
BIOME_CASES = {
    "two_ssp245": ("ssp245", _split(
        {"boreal": 0.3, "tropical": 0.7},
        boreal=dict(beta=0.5, q10_rh=2.2, warmingfactor=1.8, permafrost_c=865.0, f_nppv=0.30),
        tropical=dict(beta=0.7, q10_rh=1.6, warmingfactor=0.9, permafrost_c=0.0, f_litterd=0.95)), {}),
    # creation order differs from the name order the reference's std::map sums in
    "three_ssp585": ("ssp585", _split(
        {"tundra": 0.2, "amazon": 0.45, "midlat": 0.35},
        tundra=dict(q10_rh=2.6, warmingfactor=2.1, permafrost_c=700.0, pf_mu=1.9, fpf_static=0.6,
                    rh_ch4_frac=0.05),
        amazon=dict(beta=0.8, warmingfactor=0.8, permafrost_c=0.0, f_nppd=0.55),
        midlat=dict(beta=0.4, q10_rh=1.9, permafrost_c=165.0, pf_sigma=1.1)), dict(S=3.9, diff=1.6)),
    "even_ssp126": ("ssp126", _split({"north": 0.5, "south": 0.5}), dict(q10_rh_all=1.8)),
    "four_ssp370": ("ssp370", _split(
        {"d": 0.1, "c": 0.2, "b": 0.3, "a": 0.4},
        d=dict(warmingfactor=1.6, q10_rh=1.9), c=dict(beta=0.3), b=dict(f_nppv=0.4, f_nppd=0.5),
        a=dict(permafrost_c=0.0)), dict(lo_warming_ratio=1.5)),
    # the smallest biome's detritus pool goes negative: "Flux and pool values may not be negative"
    "four_fail_ssp370": ("ssp370", _split(
        {"d": 0.1, "c": 0.2, "b": 0.3, "a": 0.4},
        d=dict(warmingfactor=2.5, q10_rh=3.0), c=dict(beta=0.1), b=dict(f_nppv=0.5, f_nppd=0.4),
        a=dict(permafrost_c=0.0)), dict(lo_warming_ratio=1.5)),
}


# user constraints on top of biomes: {case: {constraint: {year: value}}}
BIOME_CONSTRAINTS = {
    "two_nbp_co2_ssp245": {"NBP_constrain": {y: 0.4 + 0.02 * (y - 1990) for y in range(1990, 2011)},
                           "CO2_constrain": {y: 470.0 + 2.0 * (y - 2050) for y in range(2050, 2061)}},
}
BIOME_CASES["two_nbp_co2_ssp245"] = BIOME_CASES["two_ssp245"]


def biome_ini(scn, biomes, path):
    """the shipped ini with the global [simpleNbox] pool / parameter lines replaced by
    <biome>.<name>= lines (simpleNbox.cpp:201-236)"""
    src = os.path.join(REF, "inst/input/hector_%s.ini" % scn)
    lines = []
    for line in open(src).read().splitlines():
        key = line.split("=")[0].strip()
        if key in BIOME_GLOBAL_KEYS:
            continue
        lines.append(line.replace("csv:tables/", "csv:%s/inst/input/tables/" % REF))
        if line.strip() == "[simpleNbox]":
            for b, vals in biomes.items():
                for k, v in vals.items():
                    lines.append("%s.%s=%r" % (b, k, float(v)))
    open(path, "w").write("\n".join(lines) + "\n")


def _run_constrained(ref, ini, params, constraints, variables):
    """like ref.run_member, with dated constraint entries set before prepareToRun"""
    c = ref.RefCore(ini)
    for k, v in params.items():
        c.setdata(ref.PARAM_COMPONENT[k], k, v)
    for var, d in constraints.items():
        for y, v in d.items():
            c.setvar(var, float(v), CONSTRAINT_UNITS[var], float(y))
    c.prepare()
    o = np.full((len(variables) + 1, 555), np.nan)
    ok, err = True, ""
    for y in range(1746, 2301):
        try:
            c.run(y)
        except ref.RefError as e:
            ok, err = False, str(e)
            break
        for k, v in enumerate(variables):
            o[k, y - 1746] = c.fetch(v, y)
        o[-1, y - 1746] = c.fetch_component("ocean", "ocean_timesteps")
    c.close()
    return ok, err, o


def make_biomes():
    """ref_biomes.npz: multi-biome runs of the UNMODIFIED reference (biome-split pools,
    simpleNbox-runtime.cpp:399-531)"""
    import tempfile
    from oracle import ref
    tmp = tempfile.mkdtemp()
    names, vals, fails, owns = [], [], [], []
    for name, (scn, biomes, params) in BIOME_CASES.items():
        params = dict(params)
        q10_all = params.pop("q10_rh_all", None)
        if q10_all is not None:
            for b in biomes.values():
                b["q10_rh"] = q10_all
        ini = os.path.join(tmp, name + ".ini")
        biome_ini(scn, biomes, ini)
        # the global datum (sums over biomes), then every biome's own "<biome>.<name>"
        per_biome = ["%s.%s" % (b, v) for b in biomes for v in BIOME_OWN_VARS]
        if name in BIOME_CONSTRAINTS:
            ok, err, o = _run_constrained(ref, ini, params, BIOME_CONSTRAINTS[name],
                                          BIOME_VARS + per_biome)
        else:
            ok, err, o, _ = ref.run_member(ini, params, BIOME_VARS + per_biome)
        # a failing run leaves NaN from the failing year on (the driver steps year by year)
        fail = 0 if ok else 1746 + int(np.argmax(np.isnan(o[0])))
        if not ok:
            o[:, fail - 1746:] = np.nan
        own = np.full((4, len(BIOME_OWN_VARS), 555), np.nan)
        own[:len(biomes)] = o[len(BIOME_VARS):-1].reshape(len(biomes), len(BIOME_OWN_VARS), 555)
        o = np.vstack([o[:len(BIOME_VARS)], o[-1:]])
        names.append(name); vals.append(o); fails.append(fail); owns.append(own)
        print(name, "ok" if ok else "fails in %d: %s" % (fail, err[:80]))
    import json
    spec = {n: dict(scenario=c[0], biomes=c[1], params={k: v for k, v in c[2].items()
                                                         if k != "q10_rh_all"},
                    constraints={k: {str(y): v for y, v in d.items()}
                                 for k, d in BIOME_CONSTRAINTS.get(n, {}).items()})
            for n, c in BIOME_CASES.items()}
    np.savez_compressed(os.path.join(OUT, "ref_biomes.npz"), names=np.array(names),
                        variables=np.array(BIOME_VARS + ["ocean_timesteps"]),
                        values=np.array(vals), fail_year=np.array(fails),
                        biome_variables=np.array(BIOME_OWN_VARS), biome_values=np.array(owns),
                        spec=np.array(json.dumps(spec)))


def make_biomes_stash():
    """ref_biomes_stash.npz: the per-stash outputs (HL_ocean_uptake, LL_ocean_uptake, rh_det,
    rh_soil -- the latter two summed over the biomes) of the first two multi-biome cases of
    ref_biomes.npz, from the UNMODIFIED reference"""
    import tempfile
    from oracle import ref
    tmp = tempfile.mkdtemp()
    V = ["HL_ocean_uptake", "LL_ocean_uptake", "rh_det", "rh_soil", "RH", "ocean_uptake"]
    names, vals = [], []
    for name in list(BIOME_CASES)[:2]:
        scn, biomes, params = BIOME_CASES[name]
        assert "q10_rh_all" not in params and name not in BIOME_CONSTRAINTS
        ini = os.path.join(tmp, name + ".ini")
        biome_ini(scn, biomes, ini)
        ok, err, o, _ = ref.run_member(ini, dict(params), V)
        assert ok, err
        names.append(name); vals.append(o[:len(V)])
    np.savez_compressed(os.path.join(OUT, "ref_biomes_stash.npz"), names=np.array(names),
                        variables=np.array(V), values=np.array(vals))
    print("biomes_stash", names)


def make_biomes_frozen():
    """ref_biomes_frozen.npz: f_frozen (the weighted mean of simpleNbox.cpp:492-513) and each
    biome's own <biome>.f_frozen of the first two multi-biome cases of ref_biomes.npz, from the
    UNMODIFIED reference"""
    import tempfile
    from oracle import ref
    tmp = tempfile.mkdtemp()
    names, variables, vals = [], [], []
    for name in list(BIOME_CASES)[:2]:
        scn, biomes, params = BIOME_CASES[name]
        ini = os.path.join(tmp, name + ".ini")
        biome_ini(scn, biomes, ini)
        V = ["f_frozen"] + ["%s.f_frozen" % b for b in biomes]
        ok, err, o, _ = ref.run_member(ini, dict(params), V)
        assert ok, err
        names.append(name); variables.append(",".join(V)); vals.append(np.asarray(o[:len(V)]))
    np.savez_compressed(os.path.join(OUT, "ref_biomes_frozen.npz"), names=np.array(names),
                        variables=np.array(variables), **{"values_%d" % k: v for k, v in enumerate(vals)})
    print("biomes_frozen", names, [v.shape for v in vals])


ALLPARAM_VARS = ["CO2_concentration", "global_tas", "RF_tot", "HL_pH", "CH4_concentration",
                 "N2O_concentration", "O3_concentration", "veg_c", "soil_c", "permafrost_c",
                 "ocean_c", "heatflux"]


def make_allparams():
    """ref_allparams.npz: every scalar parameter (and tau / rho / delta of four halocarbons)
    perturbed at once, three random draws on three SSPs, from the UNMODIFIED reference.  The
    draws are those of tools/sweep_params_vs_ref.py (same table of ranges, seed 4)."""
    import json
    from oracle import ref, port
    tool = os.path.join(os.path.dirname(os.path.dirname(OUT)), "tools", "sweep_params_vs_ref.py")
    src = open(tool).read().split("SCN = [")[0]   # the table of ranges only
    ns = {"__file__": tool}
    argv, sys.argv = sys.argv, [tool]
    exec(compile(src, tool, "exec"), ns)
    sys.argv = argv
    P = ns["P"]
    rng = np.random.default_rng(4)
    d = port.default_params()
    names, scns, specs, vals = [], [], [], []
    for case, scn in enumerate(["ssp119", "ssp370", "ssp585"]):
        over_ref, over_or, halo = {}, {}, {}
        for name, (comp, field, lo, hi) in P.items():
            v = float(getattr(d, field)) * rng.uniform(lo, hi)
            over_ref[(comp, name)] = v
            over_or[field] = v
        for g in rng.choice(len(port.HALOS), 4, replace=False):
            gas = port.HALOS[int(g)]
            for fld, ini in (("halo_tau", "tau"), ("halo_rho", "rho_" + gas),
                             ("halo_delta", "delta_" + gas)):
                cur = getattr(d, fld)[int(g)]
                v = float(cur * rng.uniform(0.7, 1.3)) if cur != 0 else float(rng.uniform(-0.1, 0.1))
                halo["%s[%d]" % (fld, int(g))] = v
                over_ref[(gas + "_halocarbon", ini)] = v
        ok, err, o, _ = ref.run_member(os.path.join(REF, "inst/input/hector_%s.ini" % scn),
                                       over_ref, ALLPARAM_VARS)
        assert ok, err
        names.append("allparams_%d_%s" % (case, scn)); scns.append(scn)
        specs.append(dict(params=over_or, halo=halo)); vals.append(o)
        print(names[-1], "ok")
    np.savez_compressed(os.path.join(OUT, "ref_allparams.npz"), names=np.array(names),
                        scenarios=np.array(scns), spec=np.array(json.dumps(specs)),
                        variables=np.array(ALLPARAM_VARS + ["ocean_timesteps"]),
                        values=np.array(vals))


LUC_PULSE_INI = os.path.join(REF, "tests/testthat/input/luc_pulse.ini")
# what luc_pulse.ini sets differently from the shipped scenario inis (oracle field: value)
LUC_PULSE_PARAMS = dict(end_year=1850, beta=0.0, q10_rh=1.0, npp_flux0=55.9, permafrost_c=0.0,
                        soil_c=1782.0, S=2.7, diff=2.4, rho_bc=0.0508, rho_oc=-.00621,
                        rho_so2=-.00000724, rho_nh3=-.00208)


def make_luc_pulse():
    """ref_luc_pulse.npz: the reference's own LUC-pulse case (tests/testthat/test_pulse.R with
    tests/testthat/input/luc_pulse.ini: no emissions, beta = 0, Q10 = 1, no permafrost, a run
    that ends in 1850, one land-use pulse in 1800) run by the UNMODIFIED reference, with the
    dense input table an independent reading of its csv yields."""
    from oracle import ref
    tab, ini = scenario_table(None, LUC_PULSE_INI)
    for k, v in LUC_PULSE_PARAMS.items():   # the table above is what the ini says
        sec = {"end_year": None, "S": "temperature", "diff": "temperature"}.get(k, "simpleNbox")
        if k.startswith("rho_"):
            sec = "forcing"
        if sec:
            assert float(ini[sec][k]) == v, (k, ini[sec][k], v)
    ny = int(ini["core"]["endDate"]) - int(ini["core"]["startDate"])
    variables = REF_VARS + ["luc_emissions", "NPP", "RH"]
    ok, err, o, _ = ref.run_member(LUC_PULSE_INI, {}, variables, nyears=ny)
    assert ok, err
    assert not np.isnan(o).any()
    np.savez_compressed(os.path.join(OUT, "ref_luc_pulse.npz"), table=tab,
                        param_names=np.array(list(LUC_PULSE_PARAMS)),
                        param_values=np.array(list(LUC_PULSE_PARAMS.values()), dtype=np.float64),
                        variables=np.array(variables + ["ocean_timesteps"]), values=o)
    print("luc_pulse: veg_c 1799 %.9f 1801 %.9f" % (o[variables.index("veg_c")][53], o[variables.index("veg_c")][55]))


PICONTROL_PARAMS = dict(soil_c=1782.0, permafrost_c=0.0, rho_bc=0.0508, rho_oc=-.00621,
                        rho_so2=-.00000724, rho_nh3=-.00208)


def make_picontrol():
    """ref_picontrol.npz: inst/input/hector_picontrol.ini (every series a single `name[1745]=v`
    entry, a one-entry CO2 constraint in the start year, no permafrost) run by the UNMODIFIED
    reference (tests/testthat/test_inis.R: every shipped ini makes a core)."""
    from oracle import ref
    path = os.path.join(REF, "inst/input/hector_picontrol.ini")
    tab, ini = scenario_table(None, path)
    assert ini["simpleNbox"]["CO2_constrain[1745]"] == "277.15"
    ok, err, o, _ = ref.run_member(path, {}, REF_VARS)
    assert ok, err
    np.savez_compressed(os.path.join(OUT, "ref_picontrol.npz"), table=tab,
                        param_names=np.array(list(PICONTROL_PARAMS)),
                        param_values=np.array(list(PICONTROL_PARAMS.values()), dtype=np.float64),
                        variables=np.array(REF_VARS + ["ocean_timesteps"]), values=o)
    print("picontrol: CO2 2300 %.12f Tgav 2300 %.3e" % (o[0][-1], o[1][-1]))


# (component, ini name, engine name, value) of the perturbed start-date case
STARTDATE_PERTURBED = [("simpleNbox", "C0", "C0", 285.0), ("simpleNbox", "veg_c", "veg_c", 600.0),
                       ("simpleNbox", "detritus_c", "detritus_c", 50.0),
                       ("simpleNbox", "soil_c", "soil_c", 1000.0),
                       ("simpleNbox", "permafrost_c", "permafrost_c", 700.0),
                       ("simpleNbox", "npp_flux0", "npp_flux0", 50.0),
                       ("simpleNbox", "beta", "beta", 0.4), ("simpleNbox", "q10_rh", "q10_rh", 2.0),
                       ("ocean", "preind_surface_c", "preind_surface_c", 950.0),
                       ("ocean", "preind_interdeep_c", "preind_interdeep_c", 36000.0),
                       ("CH4", "M0", "M0", 700.0), ("N2O", "N0", "N0", 270.0),
                       ("ozone", "PO3", "PO3", 28.0), ("temperature", "S", "S", 4.0)]


def make_startdate():
    """ref_startdate.npz: what the UNMODIFIED reference answers for the START date (R's
    fetchvars keeps dates >= startdate, R/messages.R:66; tests/testthat/test_parameters.R:46
    fetches the CO2 concentration of 1745): the post-spin-up pools, the preindustrial
    concentrations, zeros for temperatures / forcings / fluxes, the spin-up's last NPP and RH --
    and which variables have no entry there (NaN in the fixture: NBP).  Two cases: SSP2-4.5
    defaults and a perturbed member; the core ran to 1760 before the fetch."""
    from oracle import ref
    names, vals = [], []
    variables = REF_VARS + ["NPP", "RH", "gmst", "ocean_tas", "heatflux_mixed", "heatflux_interior"]
    for case, over in (("default", []), ("perturbed", STARTDATE_PERTURBED)):
        with ref.RefCore(os.path.join(REF, "inst/input/hector_ssp245.ini")) as core:
            for comp, ini_name, _, v in over:
                core.setdata(comp, ini_name, v)
            core.prepare()
            core.run(1760.0)
            row = []
            for v in variables:
                try:
                    row.append(core.fetch(v, 1745.0))
                except ref.RefError:
                    row.append(np.nan)
            names.append(case); vals.append(row)
            print(case, dict(zip(variables, row)))
    np.savez_compressed(os.path.join(OUT, "ref_startdate.npz"), names=np.array(names),
                        variables=np.array(variables), values=np.array(vals),
                        perturbed_names=np.array([e for _, _, e, _ in STARTDATE_PERTURBED]),
                        perturbed_values=np.array([v for _, _, _, v in STARTDATE_PERTURBED]))


OUTPUTSTREAM_YEARS = [1746, 1747, 1749, 1750, 1751, 1800, 1900, 2000, 2100]


def make_outputstream():
    """ref_outputstream_ssp245.txt: the rows the UNMODIFIED reference's CSVOutputStreamVisitor
    writes for SSP2-4.5 (src/main.cpp:91-106 through oracle/ref_driver.cpp ref_outputstream), the
    model years OUTPUTSTREAM_YEARS only (the whole file has 118 rows a year) -- row order,
    component and variable names, units text and the four significant digits every value is
    printed with (the forcing visitor sets precision(4) and returns early before the base year
    without restoring it, src/csv_outputstream_visitor.cpp:143-147)."""
    import ctypes as C
    import tempfile
    from oracle import ref
    L = ref.lib()
    L.ref_outputstream.argtypes = [C.c_char_p, C.c_double, C.c_char_p]
    tmp = os.path.join(tempfile.mkdtemp(), "os.csv")
    assert L.ref_outputstream(os.path.join(REF, "inst/input/hector_ssp245.ini").encode(), 2100.0,
                              tmp.encode()) == 0, L.ref_last_error()
    keep = []
    for line in open(tmp).read().splitlines()[2:]:
        c = line.split(",")
        if c[2] == "0" and int(c[0]) in OUTPUTSTREAM_YEARS:
            keep.append(line)
    open(os.path.join(OUT, "ref_outputstream_ssp245.txt"), "w").write("\n".join(keep) + "\n")
    print("outputstream rows kept:", len(keep))


if __name__ == "__main__":
    if "biomes_frozen" in sys.argv[1:]:
        make_biomes_frozen()
    elif "biomes_stash" in sys.argv[1:]:
        make_biomes_stash()
        make_biomes_frozen()
    elif "outputstream" in sys.argv[1:]:
        make_outputstream()
    elif "more" in sys.argv[1:]:
        make_more_outputs()
    elif "startdate" in sys.argv[1:]:
        make_startdate()
    elif "luc_pulse" in sys.argv[1:]:
        make_luc_pulse()
    elif "picontrol" in sys.argv[1:]:
        make_picontrol()
    elif "allparams" in sys.argv[1:]:
        make_allparams()
    elif "biomes" in sys.argv[1:]:
        make_biomes()
    elif "extra" in sys.argv[1:]:
        make_extra_outputs()
    elif "tracking" in sys.argv[1:]:
        make_tracking()
    elif "constraints" in sys.argv[1:]:
        make_constraints()
    else:
        main()
        make_tracking()
        make_constraints()
        make_extra_outputs()
        make_biomes()
        make_allparams()
        make_luc_pulse()
        make_picontrol()
        make_startdate()
        make_more_outputs()
        make_outputstream()
        make_biomes_stash()
        make_biomes_frozen()
