"""run time and work counters: single biome (constraint build) vs two / three biomes"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
X = lhs(M)
glob = dict(npp_flux0=56.2, veg_c=550.0, detritus_c=55.0, soil_c=917.0, permafrost_c=865.0)


def go(tag, fracs, wf=None, lo=0.0):
    ens = hb.Ensemble(M, scenario_table(), outputs=["CO2_concentration", "global_tas"],
                      biomes=list(fracs) if fracs else None)
    if fracs:
        for b, fr in fracs.items():
            ens.set_biome(b, f_nppv=0.35, f_nppd=0.60, f_litterd=0.98,
                          **{k: v * fr for k, v in glob.items()})
            ens.setvar(b + ".q10_rh", np.ascontiguousarray(X[:, 1]))
            ens.setvar(b + ".beta", np.ascontiguousarray(X[:, 2]))
        if wf:
            ens.setvar(list(fracs)[0] + ".warmingfactor", wf)
    else:
        ens.setvar("q10_rh", np.ascontiguousarray(X[:, 1]))
        ens.setvar("beta", np.ascontiguousarray(X[:, 2]))
    ens.setvar("S", np.ascontiguousarray(X[:, 0]))
    ens.setvar("diff", np.ascontiguousarray(X[:, 3]))
    if lo:
        ens.setvar("lo_warming_ratio", lo)
    ens.prepare()
    for _ in range(3):
        ens.reset(); ens.run(); ens.synchronize()
    c = ens.counters()
    my = c["member_years"]
    print("%-28s run %.2f ms  rhs/yr %.2f  stashes/yr %.3f  newton/yr %.2f  failed %d" % (
        tag, ens.last_run_ms, c["rhs_evals"] / my, c["stashes"] / my, c["newton_iterations"] / my,
        c["failed_members"]))
    ens.close()


go("global", None)
go("global, constraint build", None, lo=1.0000001)
go("2 biomes", {"a": 0.4, "b": 0.6})
go("3 biomes", {"tundra": 0.2, "amazon": 0.45, "midlat": 0.35})
go("3 biomes, tundra wf 2", {"tundra": 0.2, "amazon": 0.45, "midlat": 0.35}, wf=2.0)
