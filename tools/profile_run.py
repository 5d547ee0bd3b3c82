"""minimal driver for ncu: one engine, a couple of full runs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X = lhs(M)
ens = hb.Ensemble(M, scenario_table(), outputs=["CO2_concentration", "global_tas"])
for j, n in enumerate(PARAMS):
    ens.setvar(n, np.ascontiguousarray(X[:, j]))
ens.prepare()
for _ in range(runs):
    ens.reset()
    ens.run()
    ens.synchronize()
    print("run ms", ens.last_run_ms)
print(ens.counters())
