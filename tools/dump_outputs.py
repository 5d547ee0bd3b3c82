"""Run an M-member LHS ensemble with the library named by HECTOR_B200_LIB and save all selected
outputs (for bitwise A/B comparisons of kernel variants):
   HECTOR_B200_LIB=a.so python tools/dump_outputs.py /tmp/a.npz [M]
   python tools/dump_outputs.py --cmp /tmp/a.npz /tmp/b.npz"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if sys.argv[1] == "--cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for k in a.files:
        same = np.array_equal(a[k], b[k], equal_nan=True)
        if not same:
            bad += 1
            d = np.nanmax(np.abs(a[k] - b[k]) / np.maximum(np.abs(a[k]), 1e-300))
            print("DIFF", k, "max rel", d)
    print("bitwise identical" if bad == 0 else "%d arrays differ" % bad)
    sys.exit(1 if bad else 0)

import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS
out = sys.argv[1]
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
names = ["CO2_concentration", "global_tas", "RF_tot", "heatflux", "ocean_c", "HL_pH", "sst",
         "permafrost_c", "CH4_concentration", "NBP", "ocean_timesteps"]
X = lhs(M)
ens = hb.Ensemble(M, scenario_table(), outputs=names)
for j, n in enumerate(PARAMS):
    ens.setvar(n, np.ascontiguousarray(X[:, j]))
ens.run(2000)          # two segments: exercises the resume path (r0 != 0)
ens.run()
years = np.arange(1746, 2301, dtype=np.float64)
got = ens.fetchvars(years, names)
st, fy = ens.status()
np.savez(out, status=st, fail_year=fy, **got)
print("saved", out, "ms", ens.last_run_ms)
