"""wall-clock breakdown of one end-to-end step (host buffers in, trajectories out)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
X = lhs(M)
ens = hb.Ensemble(M, scenario_table(), outputs=["CO2_concentration", "global_tas"])
pin = [torch.from_numpy(np.ascontiguousarray(X[:, j])).pin_memory() for j in range(4)]
out = [torch.empty((M, 555), dtype=torch.float64).pin_memory() for _ in range(2)]
years = np.arange(1746, 2301, dtype=np.float64)
for j, n in enumerate(PARAMS):
    ens.setvar(n, pin[j].numpy())
ens.prepare(); ens.run(); ens.synchronize()
for it in range(3):
    t = [time.perf_counter()]
    for j, n in enumerate(PARAMS):
        ens.setvar(n, pin[j].numpy())
    ens.synchronize(); t.append(time.perf_counter())
    ens.reset(); ens.synchronize(); t.append(time.perf_counter())
    ens.run(); ens.synchronize(); t.append(time.perf_counter())
    ens.fetch("CO2_concentration", years, out=out[0].numpy()); t.append(time.perf_counter())
    ens.fetch("global_tas", years, out=out[1].numpy()); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("setvar x4 %.2f | reset(setup+spinup) %.2f | run %.2f | fetch CO2 %.2f | fetch tas %.2f | total %.2f ms"
          % (d[0], d[1], d[2], d[3], d[4], (t[-1] - t[0]) * 1e3))
