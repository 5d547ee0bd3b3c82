"""GPU vs the committed all-parameter runs of the unmodified reference
(tests/golden/ref_allparams.npz: every scalar parameter and four halocarbons' tau / rho / delta
perturbed at once).  Companion of gpu_all_params_vs_oracle.py (in the suite as
tests/test_gpu_parity.py::test_all_parameters_vs_reference_golden); the
per-scenario gas constants (N0, UC_N2O, TN2O0, halocarbon tables) go through
hx_set_param_scalar, everything else per member.

usage (under gpurun): python tools/gpu_allparams_vs_reference.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hector_b200 as hb
from oracle import port
from tests import util

ENGINE_NAME = {"preind_C_surface": "preind_surface_c", "preind_C_ID": "preind_interdeep_c"}
HALO_FIELD = {"halo_tau": "tau", "halo_rho": "rho", "halo_delta": "delta"}
YEARS = np.arange(1746, 2301, dtype=np.float64)
bad = 0
for case in util.ref_allparams():
    variables = list(case["values"])
    ens = hb.Ensemble(2, util.scenarios()[case["scenario"]], outputs=variables)
    for k, v in case["params"].items():
        ens.setvar(ENGINE_NAME.get(k, k), float(v))
    for key, v in case["halo"].items():
        fld, idx = key[:-1].split("[")
        ens.setvar("%s.%s" % (port.HALOS[int(idx)], HALO_FIELD[fld]), float(v))
    ens.run()
    st, fy = ens.status()
    got = ens.fetchvars(YEARS, variables)
    print(case["name"], "status", st.tolist())
    for v in variables:
        ref = case["values"][v]
        if v == "ocean_timesteps":
            ok = np.array_equal(got[v][0], ref)
            print("   %-20s %s" % (v, "equal" if ok else "DIFFER"))
            bad += not ok
            continue
        e = util.parity_err(got[v][0], ref, v)
        tol = 1e-10 if v in ("CO2_concentration", "global_tas") else 5e-8
        print("   %-20s %.3g %s" % (v, e, "" if e < tol else "ABOVE %g" % tol))
        bad += not (e < tol)
    ens.close()
print("violations:", bad, flush=True)
assert bad == 0
print("OK")
