"""GPU: every per-member parameter perturbed at once (one random draw per member) against the
oracle, all 38 outputs.  Written at the very end of round 1; its one run (12 members, seed 5,
seconds of GPU time left) got through the status, failing-year and per-year sub-step-count
checks of all members and then tripped the blanket 1e-10 assertion over ALL outputs -- the line
naming the variable was cut off and there was no budget to rerun.  Secondary diagnostics of
strongly perturbed members are only held to 5e-8 elsewhere (conditioning of
the high-latitude box), so the script now reports contract variables (CO2, Tgav: 1e-10) and
secondary ones (5e-8) separately and prints everything before asserting.  A CPU experiment on
the same 12 members supports the conditioning reading -- moving ONE input (diff) by one ulp moves
the oracle's own ocean_uptake by up to 2.2e-11 and RF_tot by 8e-12 of their floors, with
identical sub-step counts, and the oracle built with -ffp-contract=fast differs from itself by up
to 5.1e-11 (ocean_uptake), 3.7e-11 (RF_tot), 2.9e-11 (global_tas) -- but does not prove it.  Run this first thing next round; once
understood and green, move it into tests/test_gpu_parity.py.

usage (under gpurun): python tools/gpu_all_params_vs_oracle.py [members] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hector_b200 as hb
from oracle import port
from tests import util

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
# engine name -> (oracle field, lo factor, hi factor); N0 and the host-side gas constants are
# per-scenario scalars in the engine and stay at their defaults here
RANGES = {
    "S": ("S", 0.6, 1.6), "diff": ("diff", 0.5, 2.2), "qco2": ("qco2", 0.9, 1.1),
    "beta": ("beta", 0.3, 1.4), "q10_rh": ("q10_rh", 0.9, 2.0), "f_nppv": ("f_nppv", 0.8, 1.1),
    "f_nppd": ("f_nppd", 0.8, 1.0), "f_litterd": ("f_litterd", 0.9, 1.0),
    "npp_flux0": ("npp_flux0", 0.85, 1.15), "C0": ("C0", 0.97, 1.03), "veg_c": ("veg_c", 0.8, 1.2),
    "detritus_c": ("detritus_c", 0.8, 1.2), "soil_c": ("soil_c", 0.8, 1.2),
    "permafrost_c": ("permafrost_c", 0.5, 1.3), "warmingfactor": ("warmingfactor", 0.8, 1.6),
    "rh_ch4_frac": ("rh_ch4_frac", 0.5, 2.0), "pf_mu": ("pf_mu", 0.85, 1.2),
    "pf_sigma": ("pf_sigma", 0.8, 1.2), "fpf_static": ("fpf_static", 0.7, 1.2),
    "tt": ("tt", 0.8, 1.2), "tu": ("tu", 0.8, 1.2), "twi": ("twi", 0.8, 1.2), "tid": ("tid", 0.8, 1.2),
    "preind_surface_c": ("preind_C_surface", 0.95, 1.05),
    "preind_interdeep_c": ("preind_C_ID", 0.95, 1.05),
    "eps_abs": ("eps_abs", 0.5, 2.0), "eps_rel": ("eps_rel", 0.5, 2.0), "dt": ("dt", 0.6, 1.6),
    "eps_spinup": ("eps_spinup", 0.5, 2.0),
    "aero_scalar": ("aero_scalar", 0.5, 1.5), "vol_scalar": ("vol_scalar", 0.8, 1.2),
    "delta_co2": ("delta_co2", 0.5, 1.5), "delta_ch4": ("delta_ch4", 0.5, 1.5),
    "delta_n2o": ("delta_n2o", 0.5, 1.5), "rho_bc": ("rho_bc", 0.5, 1.5), "rho_oc": ("rho_oc", 0.5, 1.5),
    "rho_so2": ("rho_so2", 0.5, 1.5), "rho_nh3": ("rho_nh3", 0.5, 1.5),
    "M0": ("M0", 0.97, 1.03), "Tsoil": ("Tsoil", 0.8, 1.2), "Tstrat": ("Tstrat", 0.8, 1.2),
    "UC_CH4": ("UC_CH4", 0.95, 1.05), "TOH0": ("TOH0", 0.85, 1.15), "CNOX": ("CNOX", 0.7, 1.3),
    "CCO": ("CCO", 0.7, 1.3), "CNMVOC": ("CNMVOC", 0.7, 1.3), "CCH4": ("CCH4", 0.8, 1.2),
    "PO3": ("PO3", 0.9, 1.1), "lo_warming_ratio": ("lo_warming_ratio", 0.0, 0.0),
}
missing = set(hb.PARAMETERS) - set(RANGES) - {"N0"}
assert not missing, missing
d = port.default_params()
vals = {}
for name, (field, lo, hi) in RANGES.items():
    if name == "lo_warming_ratio":
        vals[name] = np.where(rng.random(M) < 0.3, rng.uniform(0.9, 1.8, M), 0.0)
    else:
        vals[name] = float(getattr(d, field)) * rng.uniform(lo, hi, M)
outs = [v for v in hb.OUTPUT_VARIABLES]
ens = hb.Ensemble(M, util.scenarios()["ssp370"], outputs=outs)
for name, v in vals.items():
    ens.setvar(name, v)
ens.run()
st, fy = ens.status()
got = ens.fetchvars(np.arange(1746, 2301, dtype=np.float64), outs)
CONTRACT = ("CO2_concentration", "global_tas")
worst = ("", -1, 0.0)
worst2 = ("", -1, 0.0)
per_var = {}
for i in range(M):
    kw = {RANGES[n][0]: float(vals[n][i]) for n in vals}
    ost, ofy, out, _, _ = port.run_member(util.scenarios()["ssp370"], **kw)
    assert (ost != 0) == (st[i] != 0) and (ost == 0 or ofy == fy[i]), (i, ost, ofy, st[i], fy[i])
    n = 555 if not ost else ofy - 1746
    for v in outs:
        ref = out[port.OUT_NAMES.index(v)][:n]
        if v == "ocean_timesteps":
            assert np.array_equal(got[v][i][:n], ref), (i, v)
            continue
        e = util.parity_err(got[v][i][:n], ref, v)
        per_var[v] = max(per_var.get(v, 0.0), e)
        if v in CONTRACT and e > worst[2]:
            worst = (v, i, e)
        if v not in CONTRACT and e > worst2[2]:
            worst2 = (v, i, e)
print("members", M, "failed", int((st != 0).sum()))
for v, e in sorted(per_var.items(), key=lambda kv: -kv[1])[:12]:
    print("  %-20s %.3g" % (v, e))
print("worst contract variable", worst, "worst secondary", worst2, flush=True)
assert worst[2] < 1e-10, worst
assert worst2[2] < 5e-8, worst2
print("OK")
