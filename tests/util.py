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


def ref_outputs_extra():
    z = np.load(os.path.join(GOLDEN, "ref_outputs_extra.npz"))
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


def ref_allparams():
    """every scalar parameter perturbed at once (tests/golden/make_golden.py allparams)"""
    import json
    z = np.load(os.path.join(GOLDEN, "ref_allparams.npz"))
    spec = json.loads(str(z["spec"]))
    variables = [str(v) for v in z["variables"]]
    return [dict(name=str(n), scenario=str(z["scenarios"][i]), params=spec[i]["params"],
                 halo=spec[i]["halo"], values=dict(zip(variables, z["values"][i])))
            for i, n in enumerate(z["names"])]
