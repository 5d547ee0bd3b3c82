"""The CPU oracle (oracle/hector_oracle.c) pinned against (a) the reference's own golden file
and (b) trajectories produced by the unmodified reference (oracle/_ref), both committed as
fixtures under tests/golden/ by tests/golden/make_golden.py.  CPU only."""
import numpy as np
import pytest

from oracle import port
from tests import util


def test_forcing_key_order():
    """forcing_component.cpp:492-495 sums a std::map<string,...> in byte-wise key order; the
    oracle hard-codes that order -- re-derive it here."""
    halos = ["CF4", "C2F6", "HFC23", "HFC32", "HFC4310", "HFC125", "HFC134a", "HFC143a",
             "HFC227ea", "HFC245fa", "SF6", "CFC11", "CFC12", "CFC113", "CFC114", "CFC115", "CCl4",
             "CH3CCl3", "HCFC22", "HCFC141b", "HCFC142b", "halon1211", "halon1301", "halon2402",
             "CH3Cl", "CH3Br"]
    keys = ["RF_" + h for h in halos] + ["RF_" + k for k in (
        "BC", "OC", "NH3", "SO2", "aci", "vol", "misc", "albedo", "CO2", "N2O", "CH4",
        "H2O_strat", "O3_trop")]
    order = sorted(keys, key=lambda s: s.encode())
    expect = ("BC C2F6 CCl4 CF4 CFC11 CFC113 CFC114 CFC115 CFC12 CH3Br CH3CCl3 CH3Cl CH4 CO2 "
              "H2O_strat HCFC141b HCFC142b HCFC22 HFC125 HFC134a HFC143a HFC227ea HFC23 HFC245fa "
              "HFC32 HFC4310 N2O NH3 O3_trop OC SF6 SO2 aci albedo halon1211 halon1301 halon2402 "
              "misc vol").split()
    assert [k[3:] for k in order] == expect
    assert len(order) == 39


def test_golden_file_ssp245():
    """tests/testthat/test_old-new.R:8-36 uses 20 years and tolerance 1e-10 (mean relative
    difference); we use all 555 years and also the pointwise metric."""
    gold, years = util.hector_comp()
    st, fy, out, cnt, sp = port.run_member(util.scenarios()["ssp245"])
    assert st == 0
    for v, g in gold.items():
        x = out[port.OUT_NAMES.index(v)]
        r = g[1:]
        mean_rel = np.abs(x - r).sum() / np.abs(r).sum()
        assert mean_rel < 1e-11, (v, mean_rel)
        assert util.parity_err(x, r, v) < 1e-10, v
    # work counters measured on the unmodified reference
    assert cnt["spinup_steps"] == 498
    assert cnt["newton_iterations"] == 122958
    assert cnt["steps_accepted"] == 2887 and cnt["steps_rejected"] == 0
    assert cnt["integrate_calls"] == 2339


def test_post_spinup_state():
    """SURVEY.md appendix C (unmodified reference, post-spin-up pools at 1745)."""
    st, fy, out, cnt, sp = port.run_member(util.scenarios()["ssp245"], run_to=1746)
    assert st == 0
    assert sp["veg"] == 561.99999976352547
    assert sp["det"] == 62.348941166518344
    assert sp["soil"] == 2030.5895679096495
    assert sp["atmos"] == 590.32949999999994
    assert sp["ocean"] == [146.61027582946105, 818.13107318705886, 8784.0600248177307,
                           27118.260117326001]
    assert 2100e-6 < sp["alk_HL"] < 2750e-6 and 2100e-6 < sp["alk_LL"] < 2750e-6


@pytest.mark.parametrize("case", util.ref_runs(), ids=lambda c: c["name"])
def test_against_reference_runs(case):
    """every committed reference trajectory: same failure verdict, same per-year ocean sub-step
    counts, values equal to ~1 ulp accumulated (observed: bit-identical)."""
    raw = util.scenarios()[case["scenario"]]
    st, fy, out, cnt, sp = port.run_member(raw, **case["params"])
    if not case["ok"]:
        assert st == 1 and "may not be negative" in case["error"]
        return
    assert st == 0
    for v, ref in case["values"].items():
        if np.isnan(ref).all():
            continue
        x = out[port.OUT_NAMES.index(v)]
        if v == "ocean_timesteps":
            assert np.array_equal(x, ref)
        else:
            assert np.allclose(x, ref, rtol=1e-12, atol=1e-13), (v, np.abs(x - ref).max())


def test_failed_member_reports_year():
    case = [c for c in util.ref_runs() if not c["ok"]][0]
    st, fy, out, cnt, sp = port.run_member(util.scenarios()[case["scenario"]], **case["params"])
    assert st == 1
    ref_years_done = int(np.sum(~np.isnan(case["values"]["CO2_concentration"])))
    assert fy == 1746 + ref_years_done
    assert np.isnan(out[0][ref_years_done:]).all() and not np.isnan(out[0][:ref_years_done]).any()


def test_chemistry_spot_values():
    """SURVEY.md appendix C chemistry spot values (alk = 2300e-6)."""
    area = 3.6e14
    h, o, it = port.csys(1.6, 146.61027582946105, 2300e-6, area * 0.15 * 100)
    assert abs(h - 1.0836826583516344e-8) < 1e-20 and it == 35
    assert abs(o[0] - 7.965097876380755) < 1e-13
    h, o, it = port.csys(20.9, 818.13107318705886, 2300e-6, area * (1 - 0.15) * 100)
    assert abs(h - 1.7116774980069958e-8) < 1e-20 and it == 35
    assert abs(o[0] - 7.766578058601485) < 1e-13


def test_reference_lib_if_present():
    """when oracle/_ref was built (build container / travelled to the GPU box) the restatement
    must be bit-identical to it on a fresh run."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    params = dict(S=4.2, q10_rh=1.7, beta=0.4, diff=2.1)
    ok, err, o, secs = ref.run_member(ref.ini_path("ssp126"), params,
                                      ["CO2_concentration", "global_tas", "HL_pH"])
    assert ok, err
    st, fy, out, cnt, sp = port.run_member(util.scenarios()["ssp126"], **params)
    assert st == 0
    for k, v in enumerate(["CO2_concentration", "global_tas", "HL_pH"]):
        assert np.array_equal(out[port.OUT_NAMES.index(v)], o[k]), v
    assert np.array_equal(out[-1], o[-1])


def test_extension_to_2500_kat():
    """SURVEY.md appendix C (unmodified reference): SSP5-8.5 with every series held at its 2300
    value to 2500 (BASELINE.json config 5 inputs, without tracking)."""
    raw = util.scenarios()["ssp585"]
    ext = np.vstack([raw, np.repeat(raw[-1:], 200, axis=0)])
    st, fy, out, cnt, sp = port.run_member(ext, port.default_params(end_year=2500))
    assert st == 0
    assert abs(out[0][2300 - 1746] - 1272.74969152107) < 1e-9
    assert abs(out[0][-1] - 1102.05842133113) < 1e-9
    assert abs(out[1][-1] - 6.60192887012203) < 1e-11


@pytest.mark.parametrize("case", util.ref_tracking(), ids=lambda c: c["name"])
def test_tracking_against_reference(case):
    """carbon tracking (fluxpool.hpp:197-257 through stashCValues / oceanbox): source fractions
    and key sets of the 11 tracked pools equal the unmodified reference's at every committed
    year (observed: bit-identical), nothing is reported before trackingDate, and switching
    tracking on leaves the trajectories bit-identical."""
    raw = util.scenarios()[case["scenario"]]
    st, fy, out, frac, mask = port.run_member_tracked(raw, case["tracking_date"],
                                                      **case["params"])
    assert st == 0
    st0, _, out0, _, _ = port.run_member(raw, **case["params"])
    assert st0 == 0 and np.array_equal(out, out0)
    first = case["tracking_date"] - 1746
    assert np.isnan(frac[:first]).all() and not mask[:first].any()
    idx = case["years"] - 1746
    assert np.array_equal(mask[idx], case["mask"])
    assert np.abs(frac[idx] - case["frac"]).max() <= 1e-15
    assert np.abs(frac[first:].sum(axis=2) - 1.0).max() < 1e-12
    # pool totals that go with the fractions
    names = ["atmos_co2", "earth_c", "veg_c", "detritus_c", "soil_c", "permafrost_c", "thawedp_c",
             "HL_ocean_c", "LL_ocean_c", "IO_ocean_c", "DO_ocean_c"]
    for k, v in enumerate(names):
        assert np.allclose(out[port.OUT_NAMES.index(v)][idx], case["pool_values"][:, k],
                           rtol=1e-13, atol=1e-13), v


def test_tracking_fluxpool_kats():
    """src/unit-testing/test_tracking.cpp:65-315 restated on the oracle's map algebra: adding a
    flux mixes source fractions by mass; subtraction and scaling keep them; a zero total splits
    evenly over the keys."""
    import ctypes as C
    L = port.lib()
    L.ho_tm_add.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_double,
                            C.POINTER(C.c_double), C.c_uint32]
    L.ho_tm_add.restype = C.c_int

    def add(a, fa, ma, b, fb, mb):
        fa = np.array(fa + [0.0] * (12 - len(fa)))
        fb = np.array(fb + [0.0] * (12 - len(fb)))
        m = C.c_uint32(ma)
        rc = L.ho_tm_add(a, fa.ctypes.data_as(C.POINTER(C.c_double)), C.byref(m), b,
                         fb.ctypes.data_as(C.POINTER(C.c_double)), mb)
        return rc, fa, m.value

    # test_tracking.cpp: c1 = 1 Pg from "a", c2 = 2 Pg from "b" -> 1/3, 2/3
    rc, f, m = add(1.0, [1.0], 0b01, 2.0, [0.0, 1.0], 0b10)
    assert rc == 0 and m == 0b11 and f[0] == 1.0 / 3.0 and f[1] == 2.0 / 3.0
    # adding more of the same source keeps fractions
    rc, f, m = add(3.0, [1.0 / 3.0, 2.0 / 3.0], 0b11, 3.0, [1.0 / 3.0, 2.0 / 3.0], 0b11)
    assert rc == 0 and abs(f[0] - 1.0 / 3.0) < 1e-16 and abs(f[1] - 2.0 / 3.0) < 1e-16
    # zero total: 1/n over the union of keys (fluxpool.hpp:248-250)
    rc, f, m = add(0.0, [1.0], 0b001, 0.0, [0.0, 0.0, 1.0], 0b100)
    assert rc == 0 and m == 0b101 and f[0] == 0.5 and f[2] == 0.5 and f[1] == 0.0
    # a key with fraction 0 stays a key
    rc, f, m = add(5.0, [1.0, 0.0], 0b11, 0.0, [0.0, 0.0, 1.0], 0b100)
    assert rc == 0 and m == 0b111 and f[0] == 1.0 and f[1] == 0.0 and f[2] == 0.0
    # fractions outside [0, 1] are rejected like the private constructor does (:105-112)
    rc, f, m = add(1.0, [1.5], 0b1, 1.0, [1.0], 0b1)
    assert rc != 0


@pytest.mark.parametrize("case", util.ref_constraints(), ids=lambda c: c["name"])
def test_constraints_against_reference(case):
    """user constraints (tests/testthat/test_constraints.R territory): CO2 / NBP / tas / RF_tot /
    CH4 / N2O / halocarbon concentration series applied exactly where the reference applies them
    (simpleNbox-runtime.cpp:343-383, 567-603, 871-898; temperature_component.cpp:510-525;
    forcing_component.cpp:498-505; ch4/n2o/halocarbon run()).  Observed: bit-identical,
    including the year and kind of the two expected failures."""
    raw = util.scenarios()["ssp245"]
    st, fy, out = port.run_member_constrained(raw, case["spec"])
    if case["fail_year"]:
        assert st in (1, 2) and fy == case["fail_year"]
    else:
        assert st == 0
    n = 555 if not case["fail_year"] else case["fail_year"] - 1746
    for v, ref in case["values"].items():
        x = out[port.OUT_NAMES.index(v)]
        assert np.allclose(x[:n], ref[:n], rtol=1e-12, atol=1e-13), (v, np.abs(x[:n] - ref[:n]).max())
        assert np.isnan(ref[n:]).all()


@pytest.mark.parametrize("case", util.ref_outputs_extra(), ids=lambda c: c["name"])
def test_extra_outputs_against_reference(case):
    """NPP, RH (final_npp / final_rh of the year's last stash), gmst, ocean_tas and the two heat
    flux components: bit-identical to the unmodified reference"""
    st, fy, out, cnt, sp = port.run_member(util.scenarios()[case["scenario"]], **case["params"])
    assert st == 0
    for v, ref in case["values"].items():
        if v not in port.OUT_NAMES:
            continue  # per-agent forcings: the engine derives them at fetch time, checked there
        assert np.array_equal(out[port.OUT_NAMES.index(v)], ref), v


@pytest.mark.parametrize("case", util.ref_biomes(), ids=lambda c: c["name"])
def test_biomes_against_reference(case):
    """biome-split pools (simpleNbox-runtime.cpp:399-531, per-biome slow parameters :965-1062):
    two, three (creation order != name order) and four biomes, one of them failing in the
    reference with a negative pool -- bit-identical up to the failing year, same year, same reason"""
    p = port.default_params()
    p.set_biomes(case["biomes"])
    st, fy, out, bio = port.run_member_biomes(util.scenarios()[case["scenario"]], p,
                                              spec=case["constraints"], **case["params"])
    n = 555
    if case["fail_year"]:
        assert (st, fy) == (1, case["fail_year"])  # HO_ERR_NEGATIVE
        n = case["fail_year"] - 1746
    else:
        assert st == 0
    for v, ref in case["values"].items():
        if v in port.OUT_NAMES:
            assert np.array_equal(out[port.OUT_NAMES.index(v)][:n], ref[:n]), v
    # every biome's own pools and final fluxes: getData("<biome>.<name>", date)
    for ib, b in enumerate(case["biomes"]):
        for k, v in enumerate(port.BIOME_OUT_NAMES):
            assert np.array_equal(bio[ib, k][:n], case["biome_values"]["%s.%s" % (b, v)][:n]), (b, v)


def test_biomes_with_tracking_are_refused():
    p = port.default_params()
    p.set_biomes(util.ref_biomes()[0]["biomes"])
    st = port.run_member_tracked(util.scenarios()["ssp245"], 1750, params=p)[0]
    assert st == 9  # HO_ERR_UNSUPPORTED


@pytest.mark.parametrize("case", util.ref_allparams(), ids=lambda c: c["name"])
def test_all_parameters_at_once_against_reference(case):
    """all 52 scalar parameters (temperature, land, ocean, solver, forcing, CH4 / OH / O3 / N2O)
    and tau / rho / delta of four halocarbons perturbed in one run: every input is wired the way
    the reference wires it -- bit-identical on 12 variables and the sub-step counts"""
    p = port.default_params()
    for key, v in case["halo"].items():
        fld, idx = key[:-1].split("[")
        getattr(p, fld)[int(idx)] = v
    st, fy, out, cnt, sp = port.run_member(util.scenarios()[case["scenario"]], params=p,
                                           **case["params"])
    assert st == 0
    for v, ref in case["values"].items():
        assert np.array_equal(out[port.OUT_NAMES.index(v)], ref), v


def test_luc_pulse_case_of_the_reference():
    """tests/testthat/test_pulse.R on input/luc_pulse.ini -- no emissions, beta = 0, Q10 = 1, no
    permafrost, end date 1850, one land-use pulse in 1800: the restatement is bit-identical to
    the unmodified reference on every variable, and the reference test's own assertions hold"""
    from oracle import port
    case = util.ref_luc_pulse()
    st, fy, out, _, _ = port.run_member(case["table"], port.default_params(**case["params"]))
    assert st == 0 and out.shape[1] == 105
    for v, ref in case["values"].items():
        if v in port.OUT_NAMES:
            assert np.array_equal(out[port.OUT_NAMES.index(v)], ref), v
    veg = out[port.OUT_NAMES.index("veg_c")]
    assert (np.diff(veg[1750 - 1746:1800 - 1746]) < 1e-6).all()   # flat after the spin-up
    assert (np.diff(veg[1801 - 1746:1851 - 1746]) < 1e-6).all()   # and after the pulse
    assert veg[1799 - 1746] - veg[1801 - 1746] > 2.0              # the pulse itself
    luc_in = case["table"][1:, 2]                                  # luc_emissions, 1746..1850
    assert np.array_equal(case["values"]["luc_emissions"], luc_in) and luc_in.sum() > 0


def test_picontrol_ini_of_the_reference():
    """inst/input/hector_picontrol.ini (tests/testthat/test_inis.R): constant inputs given as
    single entries, a one-entry CO2 constraint in the start year (it never binds: the first
    simulated year is 1746), no permafrost -- bit-identical to the unmodified reference"""
    from oracle import port
    case = util.ref_picontrol()
    st, fy, out = port.run_member_constrained(case["table"], case["constraints"], **case["params"])
    assert st == 0
    for v, ref in case["values"].items():
        assert np.array_equal(out[port.OUT_NAMES.index(v)], ref), v
    co2 = case["values"]["CO2_concentration"]
    assert np.abs(co2 - 277.15).max() < 0.01 and np.abs(case["values"]["global_tas"]).max() < 1e-3
