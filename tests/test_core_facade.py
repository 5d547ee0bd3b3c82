"""The C++ host facade (include/hector_b200_core.hpp: Core / EnsembleCore / message_data /
unitval / h_exception over the C ABI), driven by a C++ program written like the reference's own
callers (tests/cpp/test_core_facade.cpp) and checked against the CPU oracle.

Needs the reference's input data (ini + csv): /root/reference in the build container, or the
copy under oracle/_ref/input that travels to the GPU box."""
import os
import subprocess

import numpy as np
import pytest

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUT_DIRS = ["/root/reference/inst/input", os.path.join(ROOT, "oracle", "_ref", "input")]
LIBDIR = os.path.join(ROOT, "hector_b200")


def ini_path():
    for d in INPUT_DIRS:
        p = os.path.join(d, "hector_ssp245.ini")
        if os.path.exists(p):
            return p
    pytest.skip("reference input data not available")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libhector_b200.so")):
        pytest.skip("libhector_b200.so not built")
    out = str(tmp_path_factory.mktemp("facade") / "test_core_facade")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I",
           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_core_facade.cpp"),
           "-L", LIBDIR, "-lhector_b200", "-Wl,-rpath," + LIBDIR, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def parse(stdout):
    kv = {}
    for ln in stdout.splitlines():
        assert not ln.startswith("FAILED"), stdout
        if "=" in ln:
            k, v = ln.split("=", 1)
            kv[k] = v
    return kv


def test_facade_without_gpu_throws_h_exception(exe):
    """no CPU fallback behind the facade either: parse() raises h_exception naming CUDA"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, ini_path(), "nogpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "CUDA" in parse(r.stdout)["NOGPU_MSG"]


@pytest.mark.gpu
def test_facade_matches_oracle(exe):
    from oracle import port
    r = subprocess.run([exe, ini_path()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    kv = parse(r.stdout)
    assert kv["FAILURES"] == "0"
    raw = util.scenarios()["ssp245"]
    co2, tas, ph = (port.OUT_NAMES.index(v) for v in ("CO2_concentration", "global_tas", "HL_pH"))

    def yr(y):
        return y - 1746

    def close(key, ref, floor=1e-3):
        got = float(kv[key])
        assert abs(got - ref) / max(abs(ref), floor) < 1e-10, (key, got, ref)

    st, _, out, _, _ = port.run_member(raw)
    assert st == 0
    close("D_CO2_2100", out[co2][yr(2100)])
    close("D_CO2_2300", out[co2][yr(2300)])
    close("D_TAS_2300", out[tas][yr(2300)], 0.01)
    close("D_PH_2000", out[ph][yr(2000)])
    close("AGAIN_CO2_2300", out[co2][yr(2300)])      # the core survived a failed run
    st, _, out, _, _ = port.run_member(raw, S=4.5, beta=0.4)
    close("S45_CO2_2300", out[co2][yr(2300)])
    close("S45_TAS_2300", out[tas][yr(2300)], 0.01)
    close("ENS_TAS_2300_2", out[tas][yr(2300)], 0.01)  # member 2 of the batch = the same member
    close("ENS_CO2_2300_2", out[co2][yr(2300)])
    for m, S in ((0, 2.0), (1, 3.0), (3, 6.0)):
        st, _, out, _, _ = port.run_member(raw, S=S)
        close("ENS_TAS_2300_%d" % m, out[tas][yr(2300)], 0.01)
    # tracking + CH4 constraint block: 31 recorded years, at least 11 pools x 1 source each
    assert int(kv["TRACK_ROWS"]) > 31 * 11
    st, _, out = port.run_member_constrained(raw, {"CH4_constrain": {y: 1800.0 for y in range(2000, 2011)}})
    close("TRACK_CH4_2020", out[port.OUT_NAMES.index("CH4_concentration")][yr(2020)])
    # biomes: the two_ssp245 run of the unmodified reference
    bio = util.ref_biomes()[0]
    assert bio["name"] == "two_ssp245"
    close("BIOME_CO2_2300", bio["values"]["CO2_concentration"][yr(2300)])
    close("BIOME_VEG_2100", bio["values"]["veg_c"][yr(2100)])
    close("BIOME_BOREAL_VEG_2100", bio["biome_values"]["boreal.veg_c"][yr(2100)])
    # beta = 50: the oracle and the engine must agree on whether the reference aborts
    st, _, _, _, _ = port.run_member(raw, beta=50.0)
    assert (st != 0) == (kv["BETA50_FAILED"] == "1")
