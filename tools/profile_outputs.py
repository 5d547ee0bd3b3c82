import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS
M = 65536
X = lhs(M)
R4 = ["CO2_concentration", "RF_tot", "RF_CO2", "global_tas"]          # R's default fetchvars
O = list(hb.OUTPUT_VARIABLES)
for outs in (["CO2_concentration", "global_tas"], R4, O[:8], O[:16], O, O + hb.STASH_OUTPUTS):
  for keep in (False, True):
    ens = hb.Ensemble(M, scenario_table(), outputs=outs, keep_order=keep)
    for j, n in enumerate(PARAMS):
        ens.setvar(n, np.ascontiguousarray(X[:, j]))
    ens.prepare()
    for _ in range(3):
        ens.reset(); ens.run(); ens.synchronize()
    print(len(outs), "outputs, keep_order", keep, ": run ms %.2f" % ens.last_run_ms)
    ens.close()
