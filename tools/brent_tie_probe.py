"""CPU only: how often does the alkalinity equilibration (oceanbox.cpp:382-445: Brent on
|flux(alk) - target|, alkalinity LEFT AT THE LAST POINT EVALUATED) end on a different probe when
the arithmetic differs in the last ulp?  Oracle as built vs the oracle rebuilt with FMA
contraction, over members with every spin-up-relevant parameter perturbed (the all-parameter
draw).  A member whose two builds disagree on alk_HL / alk_LL differs by ~1e-6 relative in CO2
from then on: no implementation -- another compiler's build of the reference included -- can be
expected to agree with the reference to 1e-10 on such a member.
usage: python tools/brent_tie_probe.py [members] [seed]"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import port
from tests import util

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 11
scen = util.scenarios()["ssp245"]
vals = util.allparams_draw(M, seed, port.default_params())


def sweep():
    res = []
    for i in range(M):
        kw = {util.ALLPARAM_RANGES[n][0]: float(vals[n][i]) for n in vals}
        st, fy, out, cnt, sp = port.run_member(scen, run_to=1760, **kw)
        res.append((st, sp["alk_HL"], sp["alk_LL"], sp["spinup_steps"], out[0][10] if st == 0 else np.nan))
    return np.array(res)


a = sweep()
tmp = tempfile.mkdtemp()
so = os.path.join(tmp, "libhector_oracle_fma.so")
subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-mfma", "-ffp-contract=fast",
                       "-o", so, os.path.join(ROOT, "oracle", "hector_oracle.c"), "-lm"])
port.SO, port._lib = so, None
b = sweep()
alk = (np.abs(a[:, 1] - b[:, 1]) > 1e-12) | (np.abs(a[:, 2] - b[:, 2]) > 1e-12)
steps = a[:, 3] != b[:, 3]
rel = np.abs(a[:, 4] - b[:, 4]) / np.abs(a[:, 4])
print("members", M, "| different last Brent probe:", int(alk.sum()), "| different spin-up step count:",
      int(steps.sum()), "| CO2(1756) rel diff > 1e-10:", int(np.nansum(rel > 1e-10)),
      "max %.3g" % np.nanmax(rel))
for i in np.nonzero(alk | steps)[0][:10]:
    print("  member %d alk_HL %.12g vs %.12g  alk_LL %.12g vs %.12g steps %d vs %d  CO2 rel %.3g" % (
        i, a[i, 1], b[i, 1], a[i, 2], b[i, 2], a[i, 3], b[i, 3], rel[i]))
