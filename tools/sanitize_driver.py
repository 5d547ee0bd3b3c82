"""Small runs of every run-kernel instantiation for compute-sanitizer (memcheck / racecheck /
synccheck): 300 members (two full tiles + a ragged one), 40 years unless told otherwise.

  compute-sanitizer --tool memcheck  python tools/sanitize_driver.py plain
  compute-sanitizer --tool racecheck python tools/sanitize_driver.py tracked
flavours: plain allout stashout constrained nbp tracked biomes stream spinup (default: all)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hector_b200 as hb
from bench import lhs, scenario_table, PARAMS

M = int(os.environ.get("HX_SAN_MEMBERS", "300"))
TO = float(os.environ.get("HX_SAN_TO", "1785"))
flavours = sys.argv[1:] or ["plain", "allout", "stashout", "constrained", "nbp", "tracked", "biomes", "stream",
                            "spinup"]
X = lhs(M)
tab = scenario_table()


def params(e, biomes=False):
    for j, n in enumerate(PARAMS):
        if biomes and n in ("q10_rh", "beta"):   # per-biome inputs then
            for b in ("boreal", "tropical"):
                e.setvar(b + "." + n, np.ascontiguousarray(X[:, j]))
        else:
            e.setvar(n, np.ascontiguousarray(X[:, j]))


for fl in flavours:
    kw = dict(outputs=["CO2_concentration", "global_tas"])
    if fl in ("allout", "constrained", "nbp"):
        kw["outputs"] = hb.OUTPUT_VARIABLES
    if fl == "stashout":   # + the per-stash outputs parked in the scratch rows X
        kw["outputs"] = list(hb.OUTPUT_VARIABLES) + hb.STASH_OUTPUTS
    if fl == "tracked":
        kw.update(tracking_date=1750, track_every=10)
    if fl == "biomes":
        kw["biomes"] = ["boreal", "tropical"]
    e = hb.Ensemble(M, tab, **kw)
    params(e, fl == "biomes")
    if fl == "constrained":
        e.setvar_series("tas_constrain", np.arange(1760, 1770), np.linspace(0.0, 0.2, 10))
        e.setvar("lo_warming_ratio", 1.4)
    if fl == "nbp":
        e.setvar_series("NBP_constrain", np.arange(1750, 1780), np.full(30, 0.5))
    if fl == "biomes":
        for b, fr in (("boreal", 0.4), ("tropical", 0.6)):
            e.set_biome(b, f_nppv=0.35, f_nppd=0.60, f_litterd=0.98, npp_flux0=56.2 * fr,
                        veg_c=550.0 * fr, detritus_c=55.0 * fr, soil_c=917.0 * fr,
                        permafrost_c=865.0 * fr)
    if fl == "spinup":
        e.setvar("f_nppv", np.linspace(0.3, 0.4, M))  # per-member spin-up + Brent on the device
    if fl == "stream":
        got = e.run_stream(["CO2_concentration", "global_tas"], to_date=TO, segments=3)
        assert np.isfinite(got["CO2_concentration"]).all()
    else:
        e.run(TO - 17)
        e.run(TO)          # resume: r0 != 0
    st, _ = e.status()
    co2 = e.fetch("CO2_concentration", np.array([TO]))
    print("%-12s members %d to %d  failed %d  CO2 %.6f..%.6f" % (fl, M, TO, int((st != 0).sum()),
                                                             co2.min(), co2.max()), flush=True)
    e.close()
print("sanitize_driver done")
