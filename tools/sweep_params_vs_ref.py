"""Every perturbable input at once: random values for all engine parameters (and the halocarbon
tables' tau / rho / delta of a few gases), the oracle against the UNMODIFIED reference
(oracle/_ref), bit for bit.  Needs /root/reference (build container only).

usage: python tools/sweep_params_vs_ref.py [n_cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import port, ref
from tests import util

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
# engine / ini name -> (component, oracle Params field, lo factor, hi factor)
P = {
    "S": ("temperature", "S", 0.6, 1.6), "diff": ("temperature", "diff", 0.5, 2.2),
    "qco2": ("temperature", "qco2", 0.9, 1.1),
    "beta": ("simpleNbox", "beta", 0.3, 1.4), "q10_rh": ("simpleNbox", "q10_rh", 0.9, 2.0),
    "f_nppv": ("simpleNbox", "f_nppv", 0.8, 1.1), "f_nppd": ("simpleNbox", "f_nppd", 0.8, 1.0),
    "f_litterd": ("simpleNbox", "f_litterd", 0.9, 1.0),
    "npp_flux0": ("simpleNbox", "npp_flux0", 0.85, 1.15), "C0": ("simpleNbox", "C0", 0.97, 1.03),
    "veg_c": ("simpleNbox", "veg_c", 0.8, 1.2), "detritus_c": ("simpleNbox", "detritus_c", 0.8, 1.2),
    "soil_c": ("simpleNbox", "soil_c", 0.8, 1.2),
    "permafrost_c": ("simpleNbox", "permafrost_c", 0.5, 1.3),
    "warmingfactor": ("simpleNbox", "warmingfactor", 0.8, 1.6),
    "rh_ch4_frac": ("simpleNbox", "rh_ch4_frac", 0.5, 2.0), "pf_mu": ("simpleNbox", "pf_mu", 0.85, 1.2),
    "pf_sigma": ("simpleNbox", "pf_sigma", 0.8, 1.2), "fpf_static": ("simpleNbox", "fpf_static", 0.7, 1.2),
    "tt": ("ocean", "tt", 0.8, 1.2), "tu": ("ocean", "tu", 0.8, 1.2), "twi": ("ocean", "twi", 0.8, 1.2),
    "tid": ("ocean", "tid", 0.8, 1.2), "preind_surface_c": ("ocean", "preind_C_surface", 0.95, 1.05),
    "preind_interdeep_c": ("ocean", "preind_C_ID", 0.95, 1.05),
    "eps_abs": ("carbon-cycle-solver", "eps_abs", 0.5, 2.0), "eps_rel": ("carbon-cycle-solver", "eps_rel", 0.5, 2.0),
    "dt": ("carbon-cycle-solver", "dt", 0.6, 1.6), "eps_spinup": ("carbon-cycle-solver", "eps_spinup", 0.5, 2.0),
    "aero_scalar": ("forcing", "aero_scalar", 0.5, 1.5), "vol_scalar": ("forcing", "vol_scalar", 0.8, 1.2),
    "delta_co2": ("forcing", "delta_co2", 0.5, 1.5), "delta_ch4": ("forcing", "delta_ch4", 0.5, 1.5),
    "delta_n2o": ("forcing", "delta_n2o", 0.5, 1.5), "rho_bc": ("forcing", "rho_bc", 0.5, 1.5),
    "rho_oc": ("forcing", "rho_oc", 0.5, 1.5), "rho_so2": ("forcing", "rho_so2", 0.5, 1.5),
    "rho_nh3": ("forcing", "rho_nh3", 0.5, 1.5),
    "M0": ("CH4", "M0", 0.97, 1.03), "Tsoil": ("CH4", "Tsoil", 0.8, 1.2), "Tstrat": ("CH4", "Tstrat", 0.8, 1.2),
    "UC_CH4": ("CH4", "UC_CH4", 0.95, 1.05),
    "TOH0": ("OH", "TOH0", 0.85, 1.15), "CNOX": ("OH", "CNOX", 0.7, 1.3), "CCO": ("OH", "CCO", 0.7, 1.3),
    "CNMVOC": ("OH", "CNMVOC", 0.7, 1.3), "CCH4": ("OH", "CCH4", 0.8, 1.2),
    "PO3": ("ozone", "PO3", 0.9, 1.1),
    "N0": ("N2O", "N0", 0.98, 1.02), "UC_N2O": ("N2O", "UC_N2O", 0.95, 1.05), "TN2O0": ("N2O", "TN2O0", 0.9, 1.1),
}
SCN = ["ssp119", "ssp126", "ssp245", "ssp370", "ssp434", "ssp460", "ssp534-over", "ssp585"]
V = [v for v in port.OUT_NAMES if v not in ("NPP", "RH", "gmst", "ocean_tas", "heatflux_mixed",
                                            "heatflux_interior", "ocean_timesteps")]
V += ["NPP", "RH", "gmst", "ocean_tas", "heatflux_mixed", "heatflux_interior"]
d = port.default_params()
bad = 0
for case in range(N):
    scn = SCN[int(rng.integers(len(SCN)))]
    over_ref, over_or = {}, {}
    for name, (comp, field, lo, hi) in P.items():
        v = float(getattr(d, field)) * rng.uniform(lo, hi)
        over_ref[(comp, name)] = v
        over_or[field] = v
    # a few halocarbon table entries as well
    p = port.default_params()
    for g in rng.choice(len(port.HALOS), 4, replace=False):
        gas = port.HALOS[int(g)]
        for fld, ini in (("halo_tau", "tau"), ("halo_rho", "rho_" + gas), ("halo_delta", "delta_" + gas)):
            cur = getattr(p, fld)[int(g)]
            v = float(cur * rng.uniform(0.7, 1.3)) if cur != 0 else float(rng.uniform(-0.1, 0.1))
            getattr(p, fld)[int(g)] = v
            over_ref[(gas + "_halocarbon", ini)] = v
    ok, err, o, _ = ref.run_member("/root/reference/inst/input/hector_%s.ini" % scn, over_ref, V)
    fail = 0 if ok else 1746 + int(np.argmax(np.isnan(o[0])))
    st, fy, out, cnt, sp = port.run_member(util.scenarios()[scn], params=p, **over_or)
    n = 555 if not fail else fail - 1746
    same = (fail == (fy if st else 0))
    worst = ("", 0.0)
    for k, v in enumerate(V):
        a, b = out[port.OUT_NAMES.index(v)][:n], o[k][:n]
        if not np.array_equal(a, b):
            same = False
            e = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)))
            if e > worst[1]:
                worst = (v, e)
    same = same and np.array_equal(out[-1][:n], o[-1][:n])
    bad += not same
    print("case %2d %-11s ref %s  oracle status %d  %s %s" % (
        case, scn, "ok" if ok else "fails %d (%s)" % (fail, err[:50]), st,
        "BIT-IDENTICAL" if same else "MISMATCH", worst if not same else ""))
print("mismatches:", bad, "of", N)
