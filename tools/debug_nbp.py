import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from tests import util
np.set_printoptions(linewidth=220, precision=10)
case = [c for c in util.ref_constraints() if c["name"] == "nbp_near"][0]
variables = list(case["values"])
ens = hb.Ensemble(1, util.scenarios()["ssp245"], outputs=variables)
for name, d in case["spec"].items():
    ys = sorted(d)
    ens.setvar_series(name, ys, [d[y] for y in ys])
ens.run()
got = ens.fetchvars(np.arange(1746, 2301, dtype=np.float64))
for v in ["NBP", "veg_c", "soil_c", "CO2_concentration", "thawedp_c", "ocean_c", "DO_ocean_c"]:
    x = got[v][0]; r = case["values"][v]
    bad = np.nonzero(np.abs(x - r) > 1e-9 * np.maximum(1, np.abs(r)))[0]
    print(v, "first bad year", 1746 + bad[0] if len(bad) else None)
    if len(bad):
        i = bad[0]
        print("   gpu", x[i-1:i+3], "\n   ref", r[i-1:i+3], "\n   spec", [case["spec"]["NBP_constrain"].get(1746+k) for k in range(i-1, i+3)])
from oracle import port
import ctypes as C
for end in (1993, 1994, 1995, 2000):
    e = hb.Ensemble(1, util.scenarios()["ssp245"], outputs=["CO2_concentration", "ocean_timesteps"])
    for name, d in case["spec"].items():
        ys = sorted(d)
        e.setvar_series(name, ys, [d[y] for y in ys])
    e.run(end)
    c = e.counters()
    ts = e.fetch("ocean_timesteps", np.arange(1990, end + 1, dtype=np.float64))[0]
    # oracle
    p = port.default_params()
    raw = np.ascontiguousarray(util.scenarios()["ssp245"])
    cn, keep = port.make_constraints(556, 1745, case["spec"])
    out = np.empty((port.NOUT, end - 1745)); fy = C.c_int(0); cnt = port.Counters()
    st = port.lib().ho_run_member_ex(C.byref(p), port._dp(raw), C.byref(cn), end, port._dp(out), end - 1745, C.byref(fy), C.byref(cnt), None, 9999, None, None)
    print(end, "gpu", c["rhs_evals"], c["rk_steps"], c["rk_rejected"], c["stashes"], ts, "| oracle", cnt.rhs_evals, cnt.steps_accepted, cnt.steps_rejected, out[-1][1990-1746:])
    e.close()
