"""time the run kernel with carbon tracking on/off (BASELINE.json config 5 flavour, one GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
X = lhs(M)
for track in (None, 1750):
    ens = hb.Ensemble(M, scenario_table("ssp585"), outputs=["CO2_concentration", "global_tas"],
                      tracking_date=track, track_every=0)
    for j, n in enumerate(PARAMS):
        ens.setvar(n, np.ascontiguousarray(X[:, j]))
    ens.prepare()
    for _ in range(3):
        ens.reset()
        ens.run()
        ens.synchronize()
        print("tracking", track, "run ms", ens.last_run_ms, "member-years/s",
              M * 555 / ens.last_run_ms * 1e3)
    ens.close()
