"""GPU parity of user constraints (tests/testthat/test_constraints.R
territory): CO2 / tas / RF_tot / CH4 / N2O / halocarbon concentration series, through the C ABI,
against the unmodified reference's committed known answers (tests/golden/ref_constraints.npz)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

TOL = 1e-10
YEARS = np.arange(1746, 2301, dtype=np.float64)
CASES = util.ref_constraints()


def _apply(ens, spec, scenario=0):
    for name, d in spec.items():
        ys = sorted(d)
        ens.setvar_series(name, ys, [d[y] for y in ys], scenario=scenario)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_constraints_vs_reference_golden(case):
    import hector_b200 as hb
    variables = list(case["values"])
    ens = hb.Ensemble(2, util.scenarios()["ssp245"], outputs=variables)
    _apply(ens, case["spec"])
    ens.run()
    st, fy = ens.status()
    n = 555
    if case["fail_year"]:
        # the reference aborts this run ("Mass not conserved" / negative pool): same verdict
        # class, same year, NaN from that year on
        assert (st != 0).all() and (fy == case["fail_year"]).all(), (st, fy)
        n = case["fail_year"] - 1746
    else:
        assert (st == 0).all()
    got = ens.fetchvars(YEARS)
    bad = {}
    for v in variables:
        e = util.parity_err(got[v][0][:n], case["values"][v][:n], v)
        if e > TOL:
            bad[v] = e
        assert np.array_equal(got[v][0], got[v][1], equal_nan=True)
        assert np.isnan(got[v][0][n:]).all()
    assert not bad, bad
    ens.close()


def test_constraint_roundtrip_like_the_reference_test():
    """test_constraints.R: feed a run's own CH4 / N2O / CO2 back as constraints (setvar after
    the run, then reset) and get the same climate; scale the CH4 constraint and see it come
    back verbatim and warm the planet."""
    import hector_b200 as hb
    outs = ["global_tas", "CH4_concentration", "N2O_concentration", "CO2_concentration", "RF_CH4"]
    ens = hb.Ensemble(4, util.scenarios()["ssp245"], outputs=outs)
    S = np.array([2.5, 3.0, 3.5, 4.0])
    ens.setvar("S", S)
    ens.run()
    free = ens.fetchvars(YEARS)
    yrs = YEARS.astype(int)
    # all members share the CH4 / N2O trajectories only if rh_ch4 is negligible; use member 1
    ens.setvar_series("N2O_constrain", yrs, free["N2O_concentration"][1])
    ens.reset()
    ens.run()
    con = ens.fetchvars(YEARS)
    for v in outs:
        assert util.parity_err(con[v][1], free[v][1], v) < TOL, v
    # now a perturbed CH4 constraint
    new_ch4 = free["CH4_concentration"][1] * 3
    ens.setvar_series("CH4_constrain", yrs, new_ch4)
    ens.reset()
    ens.run()
    pert = ens.fetchvars(YEARS)
    assert np.array_equal(pert["CH4_concentration"][0], new_ch4)
    assert np.array_equal(pert["CH4_concentration"][3], new_ch4)
    sel = (YEARS >= 2000) & (YEARS <= 2100)
    assert (pert["global_tas"][:, sel] >= con["global_tas"][:, sel]).all()
    assert (pert["RF_CH4"][:, sel] > con["RF_CH4"][:, sel]).all()
    ens.close()


def test_constraints_per_scenario_and_perturbed_members_vs_oracle():
    """two scenarios in one engine, only one of them constrained; perturbed members against
    the oracle"""
    from oracle import port
    import hector_b200 as hb
    tabs = util.scenarios()
    case = [c for c in util.ref_constraints() if c["name"] == "combo"][0]
    M = 12
    X = util.lhs(M, seed=7)
    scen = np.arange(M) % 2
    ens = hb.Ensemble(M, [tabs["ssp245"], tabs["ssp245"]], member_scenario=scen,
                      outputs=["CO2_concentration", "global_tas", "ocean_timesteps"])
    for j, n in enumerate(["S", "q10_rh", "beta", "diff"]):
        ens.setvar(n, X[:, j])
    _apply(ens, case["spec"], scenario=1)
    ens.run()
    st, _ = ens.status()
    assert (st == 0).all()
    got = ens.fetchvars(YEARS)
    for i in (0, 1, 6, 11):
        kw = dict(S=X[i, 0], q10_rh=X[i, 1], beta=X[i, 2], diff=X[i, 3])
        if scen[i] == 1:
            ost, _, out = port.run_member_constrained(tabs["ssp245"], case["spec"], **kw)
        else:
            ost, _, out, _, _ = port.run_member(tabs["ssp245"], **kw)
        assert ost == 0
        assert np.array_equal(got["ocean_timesteps"][i], out[-1]), i
        assert util.parity_err(got["CO2_concentration"][i], out[0], "CO2_concentration") < TOL
        assert util.parity_err(got["global_tas"][i], out[1], "global_tas") < TOL
    ens.close()


def test_co2_constraint_with_tracking_sends_untracked_carbon_to_the_deep_ocean():
    """CO2 constraint + carbon tracking: the residual dumped into the deep box shows up as
    source "untracked" (fluxpool.hpp:181-192 via ocean_component.cpp:146-154)"""
    from oracle import port
    import hector_b200 as hb
    tab = util.scenarios()["ssp245"]
    case = [c for c in util.ref_constraints() if c["name"] == "co2_part"][0]
    ens = hb.Ensemble(1, tab, outputs=["CO2_concentration"], tracking_date=1900, track_every=0)
    _apply(ens, case["spec"])
    ens.run()
    frac, mask = ens.fetch_tracking(2300)
    deep = hb.TRACK_POOLS.index("deep")
    unt = hb.TRACK_SOURCES.index("untracked")
    assert mask[0, deep] >> unt & 1
    assert frac[0, deep, unt] > 0
    assert abs(frac[0].sum(axis=1) - 1).max() < 1e-12
    got = ens.fetch("CO2_concentration", YEARS)
    assert util.parity_err(got[0], case["values"]["CO2_concentration"], "CO2_concentration") < TOL
    ens.close()


def test_nbp_constraint_perturbed_members_vs_oracle():
    """NBP constraint (simpleNbox-runtime.cpp:343-383, 871-898): round(t) flips inside every
    yearly step, so the rescaled NPP / RH -- and with them the thawed-permafrost derivative --
    change between Runge-Kutta stages"""
    from oracle import port
    import hector_b200 as hb
    tab = util.scenarios()["ssp245"]
    case = [c for c in util.ref_constraints() if c["name"] == "nbp_near"][0]
    M = 8
    X = util.lhs(M, seed=11)
    outs = ["CO2_concentration", "global_tas", "NBP", "veg_c", "soil_c", "thawedp_c",
            "permafrost_c", "ocean_timesteps"]
    ens = hb.Ensemble(M, tab, outputs=outs, tracking_date=1900, track_every=0)
    for j, n in enumerate(["S", "q10_rh", "beta", "diff"]):
        ens.setvar(n, X[:, j])
    # negative adjustments too: pull NBP down late in the run, so that the pool difference that
    # goes to the deep ocean is positive and shows up as "untracked" carbon
    spec = {"NBP_constrain": dict(case["spec"]["NBP_constrain"])}
    for y in range(2060, 2081):
        spec["NBP_constrain"][y] = -0.3
    _apply(ens, spec)
    ens.run()
    st, fy = ens.status()
    got = ens.fetchvars(YEARS)
    frac, mask = ens.fetch_tracking(2300)
    n_ok = 0
    for i in range(M):
        ost, ofy, out, ofrac, omask = port.run_member_constrained(
            tab, spec, tracking_date=1900, S=X[i, 0], q10_rh=X[i, 1], beta=X[i, 2], diff=X[i, 3])
        assert (ost != 0) == (st[i] != 0), (i, ost, st[i])
        if ost:
            assert ofy == fy[i]
            continue
        n_ok += 1
        assert np.array_equal(got["ocean_timesteps"][i], out[-1]), i
        for v in outs[:-1]:
            assert util.parity_err(got[v][i], out[port.OUT_NAMES.index(v)], v) < TOL, (i, v)
        assert np.array_equal(mask[i], omask[-1]), i
        assert np.abs(frac[i] - ofrac[-1]).max() < 1e-12, i
    assert n_ok >= M // 2
    ens.close()


def test_constraints_from_ini_file(tmp_path):
    """newcore(ini) with tas_constrain=csv:tables/tas_historical.csv (the table the reference
    ships) and a dated CO2_constrain entry == the same constraints set through the API, and ==
    the oracle"""
    import csv
    import os
    from oracle import port
    import hector_b200 as hb
    from tests.test_ini_reader_cpu import constrained_ini, input_dir
    ini = constrained_ini(tmp_path)
    tas = {}
    for row in csv.reader(open(os.path.join(input_dir(), "tables", "tas_historical.csv"))):
        if len(row) == 2 and row[0].strip().isdigit():
            tas[int(row[0])] = float(row[1])
    spec = {"tas_constrain": tas, "CO2_constrain": {1900: 296.0}}
    a = hb.Ensemble.from_ini(ini, 2, outputs=["CO2_concentration", "global_tas", "sst"])
    b = hb.Ensemble(2, util.scenarios()["ssp245"], outputs=["CO2_concentration", "global_tas", "sst"])
    _apply(b, spec)
    for e in (a, b):
        e.setvar("S", np.array([2.5, 4.0]))
        e.run()
    ya, yb = a.fetchvars(YEARS), b.fetchvars(YEARS)
    for v in ya:
        assert np.array_equal(ya[v], yb[v]), v
    ost, _, out = port.run_member_constrained(util.scenarios()["ssp245"], spec, S=4.0)
    assert ost == 0
    assert util.parity_err(ya["CO2_concentration"][1], out[0], "CO2_concentration") < TOL
    assert util.parity_err(ya["global_tas"][1], out[1], "global_tas") < TOL
    yrs = np.array(sorted(tas))
    assert np.allclose(ya["global_tas"][1][yrs - 1746], [tas[y] for y in yrs], rtol=0, atol=1e-15)
    a.close()
    b.close()


def test_lo_warming_ratio_with_tas_and_co2_constraints_vs_oracle():
    """a land-ocean warming ratio on top of a tas and a CO2 constraint (the oracle is bit-identical
    to the reference for this combination): per-member ratios, the light constraint build"""
    from oracle import port
    import hector_b200 as hb
    tab = util.scenarios()["ssp245"]
    spec = {"tas_constrain": {y: 0.5 + 0.01 * (y - 1950) for y in range(1950, 2021)},
            "CO2_constrain": {y: 400.0 for y in range(2030, 2036)}}
    M = 5
    ratio = np.array([0.0, 1.2, 1.7, 0.9, 1.45])   # member 0: ratio off
    S = np.array([3.0, 2.2, 4.1, 3.6, 2.9])
    outs = ["CO2_concentration", "global_tas", "land_tas", "sst", "ocean_tas", "gmst", "veg_c",
            "heatflux", "ocean_timesteps"]
    ens = hb.Ensemble(M, tab, outputs=outs)
    ens.setvar("lo_warming_ratio", ratio)
    ens.setvar("S", S)
    _apply(ens, spec)
    ens.run()
    assert (ens.status()[0] == 0).all()
    got = ens.fetchvars(YEARS)
    for i in range(M):
        ost, _, out = port.run_member_constrained(tab, spec, S=S[i], lo_warming_ratio=ratio[i])
        assert ost == 0
        for v in outs:
            ref = out[port.OUT_NAMES.index(v)]
            if v == "ocean_timesteps":
                assert np.array_equal(got[v][i], ref), i
            else:
                assert util.parity_err(got[v][i], ref, v) < TOL, (i, v)
    ens.close()


def test_co2_tgav_only_builds_match_the_all_output_builds():
    """constraint and biome runs that record only CO2 / Tgav take kernel builds without the other
    outputs' code (like the plain runs): bit for bit what the all-output builds record"""
    import hector_b200 as hb
    tabs = util.scenarios()
    M = 160
    X = util.lhs(M, seed=11)
    minimal = ["CO2_concentration", "global_tas"]

    def run(outputs, make):
        ens = make(outputs)
        for j, n in enumerate(["S", "q10_rh", "beta", "diff"]):
            if make is biomes and n in ("q10_rh", "beta"):  # biome-specific there
                ens.setvar("boreal." + n, X[:, j])
                ens.setvar("tropical." + n, X[::-1, j].copy())
            else:
                ens.setvar(n, X[:, j])
        ens.run()
        st, fy = ens.status()
        got = ens.fetchvars(YEARS, minimal)
        ens.close()
        return st, fy, got

    # constraints without NBP (the "combo" case holds one; drop it) + lo_warming_ratio
    spec = {k: v for k, v in [c for c in util.ref_constraints() if c["name"] == "combo"][0]["spec"].items()
            if k != "NBP_constrain"}
    assert spec

    def constrained(outputs):
        ens = hb.Ensemble(M, tabs["ssp245"], outputs=outputs)
        _apply(ens, spec)
        ens.setvar("lo_warming_ratio", 1.6)
        return ens

    def biomes(outputs):
        case = util.ref_biomes()[0]
        ens = hb.Ensemble(M, tabs[case["scenario"]], outputs=outputs, biomes=list(case["biomes"]))
        for b, vals in case["biomes"].items():
            ens.set_biome(b, **vals)
        return ens

    for make in (constrained, biomes):
        a = run(minimal, make)
        b = run(minimal + ["RF_tot", "veg_c", "ocean_timesteps"], make)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        for v in minimal:
            assert np.array_equal(a[2][v], b[2][v], equal_nan=True), (make.__name__, v)
