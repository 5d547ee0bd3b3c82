import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from tests import util
from oracle import port
np.set_printoptions(linewidth=220, precision=3)
name = sys.argv[1] if len(sys.argv) > 1 else "ssp119_2000"
case = [c for c in util.ref_tracking() if c["name"] == name][0]
tab = util.scenarios()[case["scenario"]]
st, fy, out, frac, mask = port.run_member_tracked(tab, case["tracking_date"], **case["params"])
ens = hb.Ensemble(1, tab, outputs=hb.OUTPUT_VARIABLES, tracking_date=case["tracking_date"], track_every=1)
for k, v in case["params"].items():
    ens.setvar(k, v)
ens.run()
got = ens.fetchvars(np.arange(1746, 2301, dtype=np.float64))
shown = 0
for y in range(case["tracking_date"], 2301):
    f, k = ens.fetch_tracking(y)
    d = np.abs(f[0] - frac[y - 1746])
    if d.max() > 1e-13 and shown < 6:
        shown += 1
        print("year", y, "max", d.max(), "steps gpu/oracle", got["ocean_timesteps"][0][y-1746], out[-1][y-1746])
        print(d.max(axis=1))
        print("raw luc_u, daccs:", tab[y - 1745 - 1][3], tab[y - 1745 - 1][1], "thawed", out[17][y-1746], out[17][y-1747])
