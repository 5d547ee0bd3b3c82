"""Random tracked runs (scenario, tracking date, parameters, sometimes a CO2 or NBP constraint):
the oracle's source maps against the UNMODIFIED reference's (oracle/_ref), bit for bit, every
tenth year.  Needs /root/reference (build container only).

usage: python tools/sweep_tracking_vs_ref.py [n_cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
from oracle import port, ref
from tests import util
import make_golden as mg

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
SCN = ["ssp119", "ssp126", "ssp245", "ssp370", "ssp434", "ssp460", "ssp534-over", "ssp585"]
bad = 0
for case in range(N):
    scn = SCN[int(rng.integers(len(SCN)))]
    tdate = int(rng.integers(1750, 2100))
    params = dict(S=rng.uniform(2.0, 4.5), q10_rh=rng.uniform(1.1, 2.4), beta=rng.uniform(0.2, 0.8),
                  diff=rng.uniform(0.6, 2.2))
    spec = {}
    r = rng.random()
    if r < 0.25:
        a = int(rng.integers(tdate, 2250))
        spec["CO2_constrain"] = {y: float(rng.uniform(380, 520)) for y in range(a, a + 8)}
    elif r < 0.5:
        a = int(rng.integers(tdate, 2250))
        spec["NBP_constrain"] = {y: float(rng.uniform(-0.5, 1.0)) for y in range(a, a + 12)}
    c = ref.RefCore("/root/reference/inst/input/hector_%s.ini" % scn)
    c.setdata("core", "trackingDate", tdate)
    for k, v in params.items():
        c.setdata(ref.PARAM_COMPONENT[k], k, v)
    for var, d in spec.items():
        for y, v in d.items():
            c.setvar(var, float(v), mg.CONSTRAINT_UNITS[var], float(y))
    c.prepare()
    st, fy, out, frac, mask = port.run_member_constrained(util.scenarios()[scn], spec,
                                                          tracking_date=tdate, **params)
    ys = [y for y in mg.tracking_years(tdate)]
    same, fail = True, 0
    for y in ys:
        try:
            c.run(y)
        except ref.RefError:
            fail = y
            break
        on, v, f, pres = c.tracking_state()
        k = (pres.astype(np.uint32) << np.arange(12, dtype=np.uint32)).sum(1)
        i = y - 1746
        f = np.where(pres, f, 0.0)
        same = same and np.array_equal(np.where(pres, frac[i], 0.0), f) and np.array_equal(mask[i], k)
    c.close()
    if fail:
        same = same and st != 0 and fy <= fail
    else:
        same = same and st == 0
    bad += not same
    print("case %d %-11s tracking from %d %-14s ref %s  oracle status %d  %s" % (
        case, scn, tdate, "+".join(spec) or "-", "ok" if not fail else "fails <= %d" % fail, st,
        "BIT-IDENTICAL" if same else "MISMATCH"))
print("mismatches:", bad, "of", N)
