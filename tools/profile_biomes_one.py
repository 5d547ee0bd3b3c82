"""minimal driver for ncu: the three-biome ensemble of bench.py's biome_ensemble leg"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from bench import lhs, scenario_table
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
X = lhs(M)
glob = dict(npp_flux0=56.2, veg_c=550.0, detritus_c=55.0, soil_c=917.0, permafrost_c=865.0)
fracs = {"tundra": 0.2, "amazon": 0.45, "midlat": 0.35}
ens = hb.Ensemble(M, scenario_table(), outputs=["CO2_concentration", "global_tas"], biomes=list(fracs))
for b, fr in fracs.items():
    ens.set_biome(b, f_nppv=0.35, f_nppd=0.60, f_litterd=0.98, **{k: v * fr for k, v in glob.items()})
    ens.setvar(b + ".q10_rh", np.ascontiguousarray(X[:, 1]))
    ens.setvar(b + ".beta", np.ascontiguousarray(X[:, 2]))
ens.setvar("S", np.ascontiguousarray(X[:, 0]))
ens.setvar("diff", np.ascontiguousarray(X[:, 3]))
ens.prepare()
for _ in range(2):
    ens.reset(); ens.run(); ens.synchronize()
    print("run ms", ens.last_run_ms)
