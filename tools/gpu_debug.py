"""debug helper: engine vs oracle, per-variable first year of divergence"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from oracle import port
from tests import util

cold = "--cold" in sys.argv
params = {}
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("="); params[k] = float(v)
ens = hb.Ensemble(1, util.scenarios()["ssp245"], outputs=hb.OUTPUT_VARIABLES, cold_newton=cold)
for k, v in params.items():
    ens.setvar(k, v)
ens.run()
print("status", ens.status(), "counters", ens.counters())
print("spinup", ens.spinup_state(0))
years = np.arange(1746, 2301, dtype=np.float64)
got = ens.fetchvars(years)
ost, ofy, out, cnt, sp = port.run_member(util.scenarios()["ssp245"], **params)
print("oracle spin", sp)
print("oracle cnt", cnt)
for v in hb.OUTPUT_VARIABLES:
    x = got[v][0]; r = out[port.OUT_NAMES.index(v)]
    rel = np.abs(x - r) / np.maximum(np.abs(r), util.FLOOR.get(v, 1e-3))
    bad = np.nonzero(rel > 3e-11)[0]
    first = bad[0] if len(bad) else -1
    print("%-20s max rel %.3e  first>1e-11: %s  %s" % (
        v, rel.max(), 1746 + first if first >= 0 else "-",
        ("got %.17g ref %.17g" % (x[first], r[first])) if first >= 0 else ""))

v="HL_pH"; x = got[v][0]; r = out[port.OUT_NAMES.index(v)]
d = np.abs(x-r); k = np.argsort(d)[-5:]
print("worst HL_pH years", 1746+k, d[k], "timesteps", got["ocean_timesteps"][0][k])
