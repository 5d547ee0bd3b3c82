set -x
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS
M = 131072
X = lhs(M)
ens = hb.Ensemble(M, scenario_table("ssp585", end=2500) if 'end' in scenario_table.__code__.co_varnames else scenario_table("ssp585"), outputs=["CO2_concentration", "global_tas"], tracking_date=1750, track_every=0)
for j, n in enumerate(PARAMS):
    ens.setvar(n, np.ascontiguousarray(X[:, j]))
ens.prepare()
free, total = torch.cuda.mem_get_info()
print("after prepare: used GB", (total - free) / 1e9)
for _ in range(2):
    ens.reset(); ens.run(); ens.synchronize()
    print("run ms", ens.last_run_ms)
st, fy = ens.status()
print("failed", int((st != 0).sum()))
PY
