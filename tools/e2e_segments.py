"""end-to-end step time (bench.py's e2e leg) as a function of the hx_run_stream segment count"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS
M = 65536
X = lhs(M)
ens = hb.Ensemble(M, scenario_table(), outputs=["CO2_concentration", "global_tas"])
pin = [torch.from_numpy(np.ascontiguousarray(X[:, j])).pin_memory() for j in range(4)]
outs = [torch.empty((555, M), dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
V = ["CO2_concentration", "global_tas"]
for j, n in enumerate(PARAMS):
    ens.setvar(n, pin[j].numpy())
ens.prepare()
for seg in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 6, 8, 12]:
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for j, n in enumerate(PARAMS):
            ens.setvar(n, pin[j].numpy())
        ens.reset()
        ens.run_stream(V, outs=outs, segments=seg)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("segments %2d: min %.2f  median %.2f ms" % (seg, min(ts[1:]), sorted(ts[1:])[2]))
