"""Random combinations of user constraints, land-ocean warming ratio and parameters: the oracle
against the UNMODIFIED reference (oracle/_ref), bit for bit, including the failing year.
Needs /root/reference (build container only).

usage: python tools/sweep_constraints_vs_ref.py [n_cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
from oracle import port, ref
from tests import util
import make_golden as mg

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
INI = "/root/reference/inst/input/hector_ssp245.ini"
raw = util.scenarios()["ssp245"]
st0, _, base, _, _ = port.run_member(raw)          # unconstrained trajectories to perturb
V = ["CO2_concentration", "global_tas", "RF_tot", "CH4_concentration", "N2O_concentration", "NBP",
     "ocean_c", "veg_c", "soil_c", "sst", "land_tas", "heatflux", "thawedp_c", "permafrost_c"]
SRC = {"CO2_constrain": "CO2_concentration", "tas_constrain": "global_tas",
       "RF_tot_constrain": "RF_tot", "CH4_constrain": "CH4_concentration",
       "N2O_constrain": "N2O_concentration", "NBP_constrain": "NBP"}
bad = 0
for case in range(N):
    kinds = [k for k in SRC if rng.random() < 0.4] or ["CO2_constrain"]
    spec = {}
    for k in kinds:
        a = int(rng.integers(1760, 2250))
        b = min(2300, a + int(rng.integers(1, 60)))
        series = base[port.OUT_NAMES.index(SRC[k])]
        scale = 0.3 if k in ("NBP_constrain", "tas_constrain", "RF_tot_constrain") else 0.03
        spec[k] = {y: float(series[y - 1746] * (1 + scale * rng.normal()) +
                            (0.2 * rng.normal() if k == "NBP_constrain" else 0.0))
                   for y in range(a, b + 1)}
    params = dict(S=rng.uniform(2.0, 4.5), q10_rh=rng.uniform(1.1, 2.4), beta=rng.uniform(0.2, 0.8),
                  diff=rng.uniform(0.6, 2.2))
    if rng.random() < 0.4:
        params["lo_warming_ratio"] = rng.uniform(0.9, 1.8)
    ok, err, o = mg._run_constrained(ref, INI, params, spec, V)
    fail = 0 if ok else 1746 + int(np.argmax(np.isnan(o[0])))
    st, fy, out = port.run_member_constrained(raw, spec, **params)
    n = 555 if not fail else fail - 1746
    same = (fail == (fy if st else 0))
    for k, v in enumerate(V):
        same = same and np.array_equal(out[port.OUT_NAMES.index(v)][:n], o[k][:n])
    same = same and np.array_equal(out[-1][:n], o[-1][:n])
    bad += not same
    print("case %2d %-60s ref %s  oracle status %d  %s" % (
        case, "+".join(k.replace("_constrain", "") for k in kinds) +
        (" lo" if "lo_warming_ratio" in params else ""),
        "ok" if ok else "fails %d" % fail, st, "BIT-IDENTICAL" if same else "MISMATCH"))
print("mismatches:", bad, "of", N)
