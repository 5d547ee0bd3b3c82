"""The library's own ini/csv reader (hx_ini.cpp) against the fixture tables, which were produced
from the same reference input files by an independent Python reader (tests/golden/make_golden.py).
Needs the reference input data: /root/reference (build container) or oracle/_ref/input (travels
to the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest

from hector_b200 import _capi
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUT_DIRS = ["/root/reference/inst/input", os.path.join(ROOT, "oracle", "_ref", "input")]


def input_dir():
    for d in INPUT_DIRS:
        if os.path.exists(os.path.join(d, "hector_ssp245.ini")):
            return d
    pytest.skip("reference input data not available")


def need_lib():
    if not os.path.exists(_capi.lib_path()):
        pytest.skip("libhector_b200.so not built")
    return _capi.lib()


@pytest.mark.parametrize("scn", ["ssp119", "ssp126", "ssp245", "ssp370", "ssp434", "ssp460",
                                 "ssp534-over", "ssp585"])
def test_tables_match_fixtures(scn):
    L = need_lib()
    ini = os.path.join(input_dir(), "hector_%s.ini" % scn).encode()
    s, e = C.c_int32(), C.c_int32()
    assert L.hx_ini_read(ini, C.byref(s), C.byref(e), None, 0) == 0, L.hx_last_error(None)
    assert (s.value, e.value) == (1745, 2300)
    tab = np.empty((556, 44))
    assert L.hx_ini_read(ini, None, None, tab.ctypes.data_as(C.POINTER(C.c_double)), 556) == 0
    assert np.array_equal(tab, util.scenarios()[scn])


def test_scalars_match_ini_defaults():
    L = need_lib()
    ini = os.path.join(input_dir(), "hector_ssp245.ini").encode()
    expect = {"S": 3.0, "diff": 1.042, "beta": 0.65, "q10_rh": 1.2, "C0": 277.15, "baseyear": 1750,
              "rho_so2": -7.469841e-06, "delta_ch4": -.14, "CF4.tau": 50000.0, "CH3Br.H0": 5.8,
              "CFC11.delta": 0.13, "HFC125.rho": 0.000234, "TN2O0": 132, "dt": 0.25,
              "max_spinup": 2000, "preind_interdeep_c": 37100}
    for k, v in expect.items():
        out = C.c_double()
        assert L.hx_ini_scalar(ini, k.encode(), C.byref(out)) == 0, k
        assert out.value == v, (k, out.value)


def test_unsupported_inputs_are_reported(tmp_path):
    L = need_lib()
    src = open(os.path.join(input_dir(), "hector_ssp245.ini")).read()
    d = input_dir()
    bad = tmp_path / "biome.ini"  # global and biome-specific data: simpleNbox-runtime.cpp:66-69
    bad.write_text(src.replace("csv:tables/", "csv:%s/tables/" % d).replace(
        "[simpleNbox]", "[simpleNbox]\nboreal.veg_c=100"))
    assert L.hx_ini_read(str(bad).encode(), None, None, None, 0) == -1
    assert b"both global and biome-specific" in L.hx_last_error(None)
    bad4 = tmp_path / "biome_unknown.ini"
    bad4.write_text(src.replace("csv:tables/", "csv:%s/tables/" % d).replace(
        "[simpleNbox]", "[simpleNbox]\nboreal.not_a_pool=100"))
    assert L.hx_ini_read(str(bad4).encode(), None, None, None, 0) == -1
    assert b"Unknown variable" in L.hx_last_error(None)
    chem = tmp_path / "spinup_chem.ini"
    chem.write_text(src.replace("csv:tables/", "csv:%s/tables/" % d).replace(
        "[ocean]", "[ocean]\nspinup_chem=1"))
    assert L.hx_ini_read(str(chem).encode(), None, None, None, 0) == -4
    bad3 = tmp_path / "lo.ini"
    bad3.write_text(src.replace("csv:tables/", "csv:%s/tables/" % d).replace(
        "[temperature]", "[temperature]\nlo_warming_ratio=1.6"))
    assert L.hx_ini_read(str(bad3).encode(), None, None, None, 0) == 0  # an input since r1
    out = C.c_double()
    assert L.hx_ini_scalar(str(bad3).encode(), b"lo_warming_ratio", C.byref(out)) == 0
    assert out.value == 1.6
    bad2 = tmp_path / "unknown.ini"
    bad2.write_text(src.replace("csv:tables/", "csv:%s/tables/" % d).replace(
        "[temperature]", "[temperature]\nnot_a_variable=1"))
    assert L.hx_ini_read(str(bad2).encode(), None, None, None, 0) == -1
    assert b"Unknown variable" in L.hx_last_error(None)


def constrained_ini(tmp_path):
    """hector_ssp245.ini + the shipped HadCRUT temperature constraint table + a dated CO2 entry"""
    d = input_dir()
    src = open(os.path.join(d, "hector_ssp245.ini")).read()
    p = tmp_path / "ssp245_tas.ini"
    p.write_text(src.replace("csv:tables/", "csv:%s/tables/" % d).replace(
        "[temperature]", "[temperature]\ntas_constrain=csv:%s/tables/tas_historical.csv" % d).replace(
        "[simpleNbox]", "[simpleNbox]\nCO2_constrain[1900]=296.0"))
    return str(p)


def test_constraint_inputs_are_read(tmp_path):
    """tas_constrain=csv:... and dated CO2_constrain[...] entries parse (host-only check)"""
    L = need_lib()
    y0, y1 = C.c_int32(), C.c_int32()
    assert L.hx_ini_read(constrained_ini(tmp_path).encode(), C.byref(y0), C.byref(y1), None, 0) == 0
    assert (y0.value, y1.value) == (1745, 2300)


def biome_ini(tmp_path, case):
    """the shipped ini of the case's scenario with the global land inputs replaced by
    <biome>.<name> lines -- what tests/golden/make_golden.py fed the reference"""
    d = input_dir()
    drop = {"npp_flux0", "veg_c", "detritus_c", "soil_c", "permafrost_c", "f_nppv", "f_nppd",
            "f_litterd", "beta", "q10_rh"}
    lines = []
    for line in open(os.path.join(d, "hector_%s.ini" % case["scenario"])).read().splitlines():
        if line.split("=")[0].strip() in drop:
            continue
        lines.append(line.replace("csv:tables/", "csv:%s/tables/" % d))
        if line.strip() == "[simpleNbox]":
            for b, vals in case["biomes"].items():
                lines += ["%s.%s=%r" % (b, k, float(v)) for k, v in vals.items()]
    p = tmp_path / (case["name"] + ".ini")
    p.write_text("\n".join(lines) + "\n")
    return str(p)


def test_biome_ini_parses(tmp_path):
    L = need_lib()
    y0, y1 = C.c_int32(), C.c_int32()
    ini = biome_ini(tmp_path, util.ref_biomes()[1])
    assert L.hx_ini_read(ini.encode(), C.byref(y0), C.byref(y1), None, 0) == 0, L.hx_last_error(None)
    assert (y0.value, y1.value) == (1745, 2300)


def test_luc_pulse_ini_of_the_reference():
    """tests/testthat/input/luc_pulse.ini: a run that ends in 1850, natural emissions given as
    `CH4N[1750]=223` entries instead of csv columns, a csv with constraint columns nobody asks
    for -- the table equals the independent reading in tests/golden/ref_luc_pulse.npz"""
    L = need_lib()
    ini = os.path.join(input_dir(), "testthat", "luc_pulse.ini")
    if not os.path.exists(ini):
        ini = "/root/reference/tests/testthat/input/luc_pulse.ini"
    if not os.path.exists(ini):
        pytest.skip("luc_pulse.ini not available")
    case = util.ref_luc_pulse()
    s, e = C.c_int32(), C.c_int32()
    assert L.hx_ini_read(ini.encode(), C.byref(s), C.byref(e), None, 0) == 0, L.hx_last_error(None)
    assert (s.value, e.value) == (1745, 1850)
    tab = np.empty((106, 44))
    assert L.hx_ini_read(ini.encode(), None, None, tab.ctypes.data_as(C.POINTER(C.c_double)), 106) == 0
    assert np.array_equal(tab, case["table"])
    for k, v in case["params"].items():
        if k == "end_year":
            continue
        out = C.c_double()
        assert L.hx_ini_scalar(ini.encode(), k.encode(), C.byref(out)) == 0, k
        assert out.value == v, (k, out.value, v)


def test_picontrol_ini_of_the_reference():
    """every series of hector_picontrol.ini is a single `name[1745]=value` entry"""
    L = need_lib()
    ini = os.path.join(input_dir(), "hector_picontrol.ini").encode()
    case = util.ref_picontrol()
    tab = np.empty((556, 44))
    assert L.hx_ini_read(ini, None, None, tab.ctypes.data_as(C.POINTER(C.c_double)), 556) == 0
    assert np.array_equal(tab, case["table"])
