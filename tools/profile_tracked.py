"""minimal driver for ncu: tracked ensemble (SSP5-8.5, tracking from 1750), a couple of runs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
X = lhs(M)
ens = hb.Ensemble(M, scenario_table("ssp585"), outputs=["CO2_concentration", "global_tas"],
                  tracking_date=1750, track_every=0)
for j, n in enumerate(PARAMS):
    ens.setvar(n, np.ascontiguousarray(X[:, j]))
ens.prepare()
for _ in range(2):
    ens.reset(); ens.run(); ens.synchronize()
    print("run ms", ens.last_run_ms)
