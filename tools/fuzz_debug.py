"""where the combination fuzz (tests/test_gpu_fuzz.py) differs from the oracle: per variable the
member, year and values of the largest error.  usage: python tools/fuzz_debug.py seed [var ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util
from tests import test_gpu_fuzz as F
seed = int(sys.argv[1])
want = sys.argv[2:] or ["thawedp_c"]
port, ens, tabs, names, specs, ms, vals, outs = F._setup(seed, 5, 40, list(F.SRC))
st, fy = ens.status()
got = ens.fetchvars(F.YEARS, outs)
for v in want:
    best = (0.0, -1, -1)
    for i in range(40):
        kw = {util.ALLPARAM_RANGES[n][0]: float(vals[n][i]) for n in vals}
        ost, ofy, out = port.run_member_constrained(tabs[names[ms[i]]], specs[ms[i]], **kw)
        ref = out[port.OUT_NAMES.index(v)]
        e = np.abs(got[v][i] - ref) / np.maximum(np.abs(ref), util.FLOOR.get(v, 1e-3))
        j = int(np.nanargmax(e))
        if e[j] > best[0]:
            best = (float(e[j]), i, j, ref.copy(), out.copy())
    e, i, j, ref, out = best
    print(v, "worst %.3g member %d scenario %s constraints %s year %d" % (
        e, i, names[ms[i]], {k: (min(d), max(d)) for k, d in specs[ms[i]].items()}, 1746 + j))
    print("  lo_warming_ratio", vals["lo_warming_ratio"][i])
    lo, hi = max(0, j - 3), min(555, j + 4)
    for u in [v, "permafrost_c", "soil_c", "ocean_timesteps", "land_tas"]:
        if u in outs:
            print("  %-16s gpu   " % u, " ".join("%.17g" % x for x in got[u][i][lo:hi]))
            print("  %-16s oracle" % "", " ".join("%.17g" % x for x in out[port.OUT_NAMES.index(u)][lo:hi]))
    d = np.abs(got[v][i] - ref)
    nz = np.nonzero(d > 0)[0]
    print("  first differing year", 1746 + int(nz[0]) if len(nz) else None, "n differing", len(nz))
    # conditioning: the oracle against its own build with FMA contraction on the same member
    import subprocess, tempfile
    so = os.path.join(tempfile.mkdtemp(), "fma.so")
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-mfma", "-ffp-contract=fast", "-o", so,
                           os.path.join(os.path.dirname(port.__file__), "hector_oracle.c"), "-lm"])
    keep = (port.SO, port._lib)
    port.SO, port._lib = so, None
    kw = {util.ALLPARAM_RANGES[n][0]: float(vals[n][i]) for n in vals}
    _, _, out_fma = port.run_member_constrained(tabs[names[ms[i]]], specs[ms[i]], **kw)
    port.SO, port._lib = keep
    k = port.OUT_NAMES.index(v)
    e_fma = np.abs(out_fma[k] - out[k]) / np.maximum(np.abs(out[k]), util.FLOOR.get(v, 1e-3))
    print("  oracle vs its FMA build on this member: %s worst %.3g in %d  (engine vs oracle: %.3g)" % (
        v, np.nanmax(e_fma), 1746 + int(np.nanargmax(e_fma)), e))
    for u in ["permafrost_c", "thawedp_c", "soil_c", "land_tas", "NBP", "atmos_co2" if "atmos_co2" in outs else "CO2_concentration"]:
        dd = np.abs(got[u][i] - out[port.OUT_NAMES.index(u)])
        print("  |gpu - oracle| %-18s" % u, " ".join("%d:%.1e" % (1746 + t, dd[t]) for t in list(range(0, 555, 30)) + list(range(478, 496))))
