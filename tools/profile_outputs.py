import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS
M = 65536
X = lhs(M)
for outs in (["CO2_concentration", "global_tas"], ["CO2_concentration", "global_tas", "RF_tot", "HL_pH", "veg_c"]):
    ens = hb.Ensemble(M, scenario_table(), outputs=outs)
    for j, n in enumerate(PARAMS):
        ens.setvar(n, np.ascontiguousarray(X[:, j]))
    ens.prepare()
    for _ in range(3):
        ens.reset(); ens.run(); ens.synchronize()
    print(len(outs), "outputs: run ms", ens.last_run_ms)
    ens.close()
