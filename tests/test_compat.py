"""The drop-in boundary, literally: the reference's own callers compiled
UNMODIFIED against include/compat + libhector_b200.so --

  src/main.cpp          the CLI (`hector <ini>`): runs an ini file, writes outputstream_<run>.csv
  src/rcpp_hector.cpp   the R glue, with a stub Rcpp.h, driven the way R/hector.R and R/messages.R
                        drive it (tests/cpp/test_rcpp_glue.cpp)

CPU: both compile and link, and without a GPU they fail the way the reference reports errors.
GPU: their numbers against the CPU oracle.  The binaries are built by tests/cpp/build_compat.py
from the reference sources where they lie (build container only; the GPU box gets the binaries)."""
import csv
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cpp"))
import build_compat  # noqa: E402

INPUT_DIRS = ["/root/reference/inst/input", os.path.join(ROOT, "oracle", "_ref", "input")]


def ini_path():
    for d in INPUT_DIRS:
        p = os.path.join(d, "hector_ssp245.ini")
        if os.path.exists(p):
            return p
    pytest.skip("reference input data not available")


@pytest.fixture(scope="module")
def programs():
    if not os.path.exists(os.path.join(ROOT, "hector_b200", "libhector_b200.so")):
        pytest.skip("libhector_b200.so not built")
    if build_compat.reference_present():
        try:
            return build_compat.build()
        except subprocess.CalledProcessError as e:
            pytest.fail("unmodified reference callers do not compile against include/compat:\n" + e.stderr)
    out = {n: os.path.join(build_compat.OUT, n) for n in ("hector_cli", "rcpp_glue")}
    if not all(os.path.exists(p) for p in out.values()):
        pytest.skip("reference sources absent and no prebuilt binaries (run __graft_entry__.build())")
    return out


def test_unmodified_callers_compile_and_link(programs):
    for p in programs.values():
        assert os.access(p, os.X_OK)


def test_cli_without_gpu_reports_like_the_reference(programs, tmp_path):
    """no CPU fallback behind the CLI either: main.cpp's catch block prints the h_exception"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([programs["hector_cli"], ini_path()], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1, (r.returncode, r.stderr)
    assert "* Program exception:" in r.stderr and "CUDA" in r.stderr
    r = subprocess.run([programs["hector_cli"]], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "Usage: <program> <config file name>" in r.stderr


def _kv(stdout):
    return dict(ln.split("=", 1) for ln in stdout.splitlines() if "=" in ln)


@pytest.mark.gpu
def test_cli_runs_ssp245_and_matches_the_oracle(programs, tmp_path):
    from oracle import port
    r = subprocess.run([programs["hector_cli"], ini_path()], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    path = tmp_path / "output" / "outputstream_ssp245.csv"
    lines = open(path).read().splitlines()
    assert lines[0].startswith("# Output from hector version")
    assert lines[1] == "year,run_name,spinup,component,variable,value,units"
    rows = list(csv.reader(lines[2:]))
    st, _, out, _, _ = port.run_member(util.scenarios()["ssp245"])
    assert st == 0
    seen = {}
    for y, run, spin, comp, var, val, units in rows:
        assert run == "ssp245" and spin == "0"
        seen.setdefault(var, {})[int(y)] = (float(val), comp, units)
    assert set(seen["CO2_concentration"]) == set(range(1746, 2301))
    # four significant digits everywhere, like the reference's own file (see the visitor's header)
    for var, digits in (("CO2_concentration", 4), ("global_tas", 4), ("RF_tot", 4), ("HL_pH", 4),
                        ("veg_c", 4), ("ocean_uptake", 4)):
        ref = out[port.OUT_NAMES.index(var)]
        for y in (1746, 1850, 2000, 2100, 2300):
            if var == "RF_tot" and y < 1750:    # forcings are printed from the base year on
                assert y not in seen[var]
                continue
            got = seen[var][y][0]
            want = float("%.*g" % (digits, ref[y - 1746]))   # the stream's significant digits
            assert abs(got - want) <= 1.01 * 10.0 ** (np.floor(np.log10(max(abs(want), 1e-300))) - digits + 1) \
                or abs(got - want) < 1e-12, (var, y, got, want)
    assert seen["global_tas"][2100][1:] == ("temperature", "degC")
    assert seen["RF_tot"][2100][1:] == ("forcing", "W/m2")
    # tracking is off in the shipped ini: the tracking file exists and is empty
    assert os.path.getsize(tmp_path / "output" / "tracking_ssp245.csv") == 0
    # Row for row against the UNMODIFIED reference's own outputstream (tests/golden/
    # ref_outputstream_ssp245.txt, nine model years): the same rows in the same order with the
    # same component, variable and units text, values equal to the printed four digits (one unit
    # of the last digit where 1e-10 of difference crosses a rounding boundary) -- but for the
    # rows the engine has no number for
    skipped = {"HL_downwelling", "HL_Revelle", "LL_Revelle", "atmos_c_residual", "slr", "slr_no_ice",
               "sl_rc", "sl_rc_no_ice"}
    ref_rows = [r for r in csv.reader(open(os.path.join(os.path.dirname(__file__), "golden",
                                                        "ref_outputstream_ssp245.txt")))
                if r[4] not in skipped]
    years = sorted({int(r[0]) for r in ref_rows})
    mine = [r for r in rows if int(r[0]) in years]
    assert [r[:5] + r[6:] for r in mine] == [r[:5] + r[6:] for r in ref_rows]
    worst = 0.0
    for a, b in zip(mine, ref_rows):
        got, want = float(a[5]), float(b[5])
        ulp4 = 10.0 ** (np.floor(np.log10(max(abs(want), 1e-300))) - 3)
        assert abs(got - want) <= 1.01 * ulp4 or abs(got - want) < 1e-12, (a, b)
        worst = max(worst, abs(got - want) / ulp4 if want else 0.0)
    same = sum(a[5] == b[5] for a, b in zip(mine, ref_rows))
    print("outputstream rows compared: %d, textually identical values: %d" % (len(mine), same))
    assert same >= 0.99 * len(mine)


@pytest.mark.gpu
def test_r_glue_matches_the_oracle(programs):
    from oracle import port
    r = subprocess.run([programs["rcpp_glue"], ini_path()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    kv = _kv(r.stdout)
    assert kv["DONE"] == "1"
    assert (kv["STRT"], kv["END"], kv["TRACK"], kv["DATE"]) == ("1745", "2300", "9999", "2100")
    raw = util.scenarios()["ssp245"]
    co2, tas = port.OUT_NAMES.index("CO2_concentration"), port.OUT_NAMES.index("global_tas")

    def close(key, ref, floor=1e-3):
        got = float(kv[key])
        assert abs(got - ref) / max(abs(ref), floor) < 1e-10, (key, got, ref)

    st, _, out, _, _ = port.run_member(raw)
    close("TAS_2000", out[tas][2000 - 1746], 0.01)
    close("TAS_2100", out[tas][2100 - 1746], 0.01)
    close("CO2_2100", out[co2][2100 - 1746])
    assert kv["TAS_UNITS"] == "degC"
    st, _, out, _, _ = port.run_member(raw, S=4.5)
    close("S45_TAS_2300", out[tas][-1], 0.01)
    close("S45_CO2_2300", out[co2][-1])
    # dated setvar: eleven years of fossil emissions set to zero, S still 4.5
    edited = raw.copy()
    edited[2030 - 1745:2041 - 1745, 0] = 0.0
    st, _, out, _, _ = port.run_member(edited, S=4.5)
    close("FFI0_CO2_2100", out[co2][2100 - 1746])
    assert kv["BAD_UNIT_REFUSED"] == "1" and kv["STILL_VALID"] == "1"
    assert kv["VALID_AFTER_SHUTDOWN"] == "0"
    # biomes through rename_biome / create_biome + setvar
    assert kv["BIOMES0"] == "global" and kv["BIOMES"] == "boreal,tropical"
    assert kv["BIOMES_AFTER_DELETE"] == "1"
    d = port.default_params()
    glob = dict(veg_c=d.veg_c, detritus_c=d.detritus_c, soil_c=d.soil_c, permafrost_c=d.permafrost_c,
                npp_flux0=d.npp_flux0)
    common = dict(beta=d.beta, q10_rh=d.q10_rh, f_nppv=d.f_nppv, f_nppd=d.f_nppd, f_litterd=d.f_litterd)
    boreal = dict(common, warmingfactor=1.8, **{k: 0.4 * v for k, v in glob.items()})
    tropical = dict(common, beta=0.5, q10_rh=2.2, f_nppv=0.35, f_nppd=0.60, f_litterd=0.98,
                    **{k: 0.6 * v for k, v in glob.items()})
    p = port.default_params().set_biomes({"boreal": boreal, "tropical": tropical})
    st, _, out, bio = port.run_member_biomes(raw, p)
    assert st == 0
    close("BIO_CO2_2300", out[co2][-1])
    close("BIO_VEG_2100", out[port.OUT_NAMES.index("veg_c")][2100 - 1746])
    close("BIO_BOREAL_VEG_2100", bio[0][port.BIOME_OUT_NAMES.index("veg_c")][2100 - 1746])
