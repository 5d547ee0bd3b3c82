"""Shared helpers for the tests: fixtures loading, parity metric."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_cache = {}


def scenarios():
    if "sc" not in _cache:
        z = np.load(os.path.join(GOLDEN, "scenarios.npz"))
        _cache["sc"] = {k.replace("_over", "-over") if k.endswith("_over") else k: z[k]
                        for k in z.files if k != "names"}
        _cache["sc_names"] = [str(s) for s in z["names"]]
    return _cache["sc"]


def raw_names():
    scenarios()
    return _cache["sc_names"]


def hector_comp():
    z = np.load(os.path.join(GOLDEN, "hector_comp.npz"))
    return dict(zip([str(v) for v in z["variables"]], z["values"])), z["years"]


def ref_runs():
    z = np.load(os.path.join(GOLDEN, "ref_runs.npz"))
    cases = []
    variables = [str(v) for v in z["variables"]]
    for i, name in enumerate(z["names"]):
        pn = str(z["param_names"][i])
        keys = pn.split(",") if pn else []
        params = {k: float(z["param_values"][i][j]) for j, k in enumerate(keys)}
        cases.append(dict(name=str(name), scenario=str(z["scenarios"][i]), params=params,
                          ok=bool(z["ok"][i]), error=str(z["errors"][i]),
                          values=dict(zip(variables, z["values"][i]))))
    return cases


def ref_tracking():
    """carbon-tracking known answers of the unmodified reference (tests/golden/make_golden.py)"""
    z = np.load(os.path.join(GOLDEN, "ref_tracking.npz"))
    cases = []
    for i, name in enumerate(z["names"]):
        pn = str(z["param_names"][i])
        keys = pn.split(",") if pn else []
        n = int((z["years"][i] >= 0).sum())
        cases.append(dict(name=str(name), scenario=str(z["scenarios"][i]),
                          tracking_date=int(z["tracking_dates"][i]),
                          params={k: float(z["param_values"][i][j]) for j, k in enumerate(keys)},
                          years=z["years"][i][:n], frac=z["frac"][i][:n], mask=z["mask"][i][:n],
                          pool_values=z["pool_values"][i][:n]))
    return cases


def ref_tracking_csv():
    return str(np.load(os.path.join(GOLDEN, "ref_tracking.npz"))["csv_1750_1755"])


def ref_outputs_more():
    """the rest of R's ALL_VARS() from the unmodified reference (make_golden.py more)"""
    return ref_outputs_extra("ref_outputs_more.npz")


def ref_outputs_extra(fixture="ref_outputs_extra.npz"):
    z = np.load(os.path.join(GOLDEN, fixture))
    variables = [str(v) for v in z["variables"]]
    cases = []
    for i, name in enumerate(z["names"]):
        pn = str(z["param_names"][i])
        keys = pn.split(",") if pn else []
        cases.append(dict(name=str(name), scenario=str(z["scenarios"][i]),
                          params={k: float(z["param_values"][i][j]) for j, k in enumerate(keys)},
                          values=dict(zip(variables, z["values"][i]))))
    return cases


def ref_constraints():
    """constraint known answers of the unmodified reference (tests/golden/make_golden.py)"""
    z = np.load(os.path.join(GOLDEN, "ref_constraints.npz"))
    cnames = [str(s) for s in z["constraint_names"]]
    variables = [str(v) for v in z["variables"]]
    cases = []
    for i, name in enumerate(z["names"]):
        spec = {}
        for k, y, v in z["spec"][i]:
            if not np.isnan(k):
                spec.setdefault(cnames[int(k)], {})[int(y)] = float(v)
        cases.append(dict(name=str(name), spec=spec, fail_year=int(z["fail_year"][i]),
                          values=dict(zip(variables, z["values"][i]))))
    return cases


# natural scale below which a relative error is meaningless (outputs that pass through zero
# or are differences of large pools); CO2 / Tgav floors are SURVEY.md section 8(d)'s
FLOOR = {"global_tas": 0.01, "CO2_concentration": 1.0, "sst": 0.01, "land_tas": 0.01,
         "heatflux": 0.1, "RF_tot": 0.01, "RF_CO2": 0.01, "RF_CH4": 0.01, "RF_N2O": 0.01,
         "NBP": 1.0, "ocean_uptake": 1.0, "thawedp_c": 1.0, "rh_ch4": 1e-3, "gmst": 0.01,
         "ocean_tas": 0.01, "heatflux_mixed": 0.1, "heatflux_interior": 0.1}


def parity_err(x, ref, var):
    """max_t |x - ref| / max(|ref|, floor_var)  """
    floor = FLOOR.get(var, 1e-3)
    return float(np.max(np.abs(x - ref) / np.maximum(np.abs(ref), floor)))


def lhs(M, seed=20241017):
    """SURVEY.md section 8(d) sampler: columns (S, q10_rh, beta, diff)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo = np.array([2.0, 1.0, 0.2, 0.5])
    hi = np.array([5.0, 2.6, 0.9, 2.5])
    cols = []
    for j in range(4):
        u = (rng.permutation(M) + rng.random(M)) / M
        cols.append(lo[j] + u * (hi[j] - lo[j]))
    return np.stack(cols, axis=1)


def ref_biomes():
    """multi-biome known answers of the unmodified reference (tests/golden/make_golden.py biomes)"""
    import json
    z = np.load(os.path.join(GOLDEN, "ref_biomes.npz"))
    spec = json.loads(str(z["spec"]))
    variables = [str(v) for v in z["variables"]]
    cases = []
    for i, name in enumerate(z["names"]):
        sp = spec[str(name)]
        own = {"%s.%s" % (b, str(v)): z["biome_values"][i][ib][k]
               for ib, b in enumerate(sp["biomes"]) for k, v in enumerate(z["biome_variables"])}
        cases.append(dict(name=str(name), scenario=sp["scenario"], biomes=sp["biomes"],
                          params=sp["params"], fail_year=int(z["fail_year"][i]),
                          constraints={k: {int(y): v for y, v in d.items()}
                                       for k, d in sp.get("constraints", {}).items()},
                          values=dict(zip(variables, z["values"][i])), biome_values=own))
    return cases


def ref_luc_pulse():
    """the reference's own LUC-pulse case (tests/testthat/test_pulse.R, input/luc_pulse.ini) as
    the unmodified reference ran it: input table [106, 44], the parameters the ini sets, outputs
    over 1746..1850 (tests/golden/make_golden.py luc_pulse)"""
    z = np.load(os.path.join(GOLDEN, "ref_luc_pulse.npz"))
    params = {str(n): float(v) for n, v in zip(z["param_names"], z["param_values"])}
    params["end_year"] = int(params["end_year"])
    variables = [str(v) for v in z["variables"]]
    return dict(table=z["table"], params=params, values=dict(zip(variables, z["values"])))


def ref_picontrol():
    """inst/input/hector_picontrol.ini as the unmodified reference ran it (make_golden.py
    picontrol): input table [556, 44], the parameters the ini sets, 31 variables x 555 years"""
    z = np.load(os.path.join(GOLDEN, "ref_picontrol.npz"))
    params = {str(n): float(v) for n, v in zip(z["param_names"], z["param_values"])}
    variables = [str(v) for v in z["variables"]]
    return dict(table=z["table"], params=params, values=dict(zip(variables, z["values"])),
                constraints={"CO2_constrain": {1745: 277.15}})


def ref_startdate():
    """what the unmodified reference answers for the start date (make_golden.py startdate):
    [{name, params {engine name: value}, values {variable: value or NaN = no entry}}]"""
    z = np.load(os.path.join(GOLDEN, "ref_startdate.npz"))
    variables = [str(v) for v in z["variables"]]
    pert = {str(n): float(v) for n, v in zip(z["perturbed_names"], z["perturbed_values"])}
    return [dict(name=str(n), params=pert if str(n) == "perturbed" else {},
                 values=dict(zip(variables, z["values"][i]))) for i, n in enumerate(z["names"])]


def ref_allparams():
    """every scalar parameter perturbed at once (tests/golden/make_golden.py allparams)"""
    import json
    z = np.load(os.path.join(GOLDEN, "ref_allparams.npz"))
    spec = json.loads(str(z["spec"]))
    variables = [str(v) for v in z["variables"]]
    return [dict(name=str(n), scenario=str(z["scenarios"][i]), params=spec[i]["params"],
                 halo=spec[i]["halo"], values=dict(zip(variables, z["values"][i])))
            for i, n in enumerate(z["names"])]


# engine parameter name -> (oracle field, lo factor, hi factor): the all-parameter draw of
# tools/gpu_all_params_vs_oracle.py / tests/test_gpu_parity.py::test_all_parameters_at_once.
# N0 and the host-side gas constants are per-scenario in the sweep and stay at their defaults.
ALLPARAM_RANGES = {
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


def allparams_draw(M, seed, defaults):
    """every per-member parameter perturbed at once: {engine name: values[M]}; `defaults` is
    the oracle's default parameter block (oracle.port.default_params())"""
    rng = np.random.default_rng(seed)
    vals = {}
    for name, (field, lo, hi) in ALLPARAM_RANGES.items():
        if name == "lo_warming_ratio":
            vals[name] = np.where(rng.random(M) < 0.3, rng.uniform(0.9, 1.8, M), 0.0)
        else:
            vals[name] = float(getattr(defaults, field)) * rng.uniform(lo, hi, M)
    return vals
