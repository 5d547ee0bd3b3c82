"""GPU: results must not depend on how the engine orders the members internally.  Runs the
'extreme members' ensemble of tests/test_gpu_parity_at_size.py (per-member spin-up, many failures)
and the LHS headline ensemble with and without HX_FLAG_KEEP_ORDER, twice each, and compares bit
for bit; then lists the members that differ most from the oracle.
usage: python tools/gpu_order_invariance.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from oracle import port
from tests import util

years = np.arange(1746, 2301, dtype=np.float64)
outs = ["CO2_concentration", "global_tas", "ocean_timesteps"]


def run(raw, draw, M, **kw):
    e = hb.Ensemble(M, raw, outputs=outs, **kw)
    for k, v in draw.items():
        e.setvar(k, v)
    e.run()
    st, fy = e.status()
    g = e.fetchvars(years, outs)
    sp = np.array([[e.spinup_state(i)[k] for k in ("alk_HL", "alk_LL", "spinup_steps")] for i in range(min(M, 512))])
    e.close()
    return st, fy, g, sp


raw = util.scenarios()["ssp585"]
rng = np.random.default_rng(77)
M = 512
draw = {"S": rng.uniform(4.5, 9.0, M), "q10_rh": rng.uniform(2.5, 5.0, M),
        "beta": rng.uniform(0.01, 0.4, M), "diff": rng.uniform(0.1, 1.0, M),
        "detritus_c": rng.uniform(3.0, 40.0, M), "veg_c": rng.uniform(60.0, 400.0, M)}
a = run(raw, draw, M)
b = run(raw, draw, M)
c = run(raw, draw, M, keep_order=True)
for name, x, y in (("sorted vs sorted again", a, b), ("sorted vs caller order", a, c)):
    same = all(np.array_equal(x[2][v], y[2][v], equal_nan=True) for v in outs) and np.array_equal(x[0], y[0]) \
        and np.array_equal(x[1], y[1]) and np.array_equal(x[3], y[3])
    print("extreme ensemble,", name, ":", "bit-identical" if same else "DIFFERENT")
    if not same:
        for v in outs:
            d = np.nanmax(np.abs(x[2][v] - y[2][v]), axis=1)
            print("   ", v, "members differing:", np.nonzero(d > 0)[0][:10], "max", np.nanmax(d))
worst = []
for i in range(M):
    ost, ofy, out, _, sp = port.run_member(raw, **{k: float(v[i]) for k, v in draw.items()})
    n = 555 if not ost else ofy - 1746
    if n <= 0:
        continue
    e = util.parity_err(a[2]["CO2_concentration"][i][:n], out[0][:n], "CO2_concentration")
    worst.append((e, i, n, a[3][i][0] - sp["alk_HL"], a[3][i][1] - sp["alk_LL"], a[3][i][2] - sp["spinup_steps"]))
worst.sort(reverse=True)
print("worst members vs oracle (CO2 err, member, years, d alk_HL, d alk_LL, d spinup steps):")
for w in worst[:8]:
    print("   %.3g  %d  %d  %.3g  %.3g  %d" % w)
X = util.lhs(8192)
d4 = {n: X[:, j] for j, n in enumerate(["S", "q10_rh", "beta", "diff"])}
p = run(util.scenarios()["ssp245"], d4, 8192)
q = run(util.scenarios()["ssp245"], d4, 8192, keep_order=True)
same = all(np.array_equal(p[2][v], q[2][v], equal_nan=True) for v in outs)
print("LHS 8192, sorted vs caller order:", "bit-identical" if same else "DIFFERENT")
