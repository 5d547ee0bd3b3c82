"""Ensemble: the batched counterpart of a reference Core.

    reference (R, one member)                      here (M members at once)
    -------------------------                      ------------------------
    core <- newcore(inifile)                       ens = Ensemble(M, scenario tables)
    setvar(core, NA, ECS(), 3.5, "degC")           ens.setvar("S", per_member_array or scalar)
    reset(core); run(core, 2100)                   ens.reset(); ens.run(2100)
    fetchvars(core, 1746:2100, c(...))             ens.fetchvars(range(1746, 2101), [...])

Variable / parameter names are the reference's capability strings
(inst/include/component_data.hpp).
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import HxError

HALOS = ["CF4", "C2F6", "HFC23", "HFC32", "HFC4310", "HFC125", "HFC134a", "HFC143a", "HFC227ea",
         "HFC245fa", "SF6", "CFC11", "CFC12", "CFC113", "CFC114", "CFC115", "CCl4", "CH3CCl3",
         "HCFC22", "HCFC141b", "HCFC142b", "halon1211", "halon1301", "halon2402", "CH3Cl",
         "CH3Br"]
RAW_SERIES = ["ffi_emissions", "daccs_uptake", "luc_emissions", "luc_uptake", "CH4_emissions",
              "CH4N", "NOX_emissions", "CO_emissions", "NMVOC_emissions", "BC_emissions",
              "OC_emissions", "SO2_emissions", "NH3_emissions", "SV", "RF_albedo", "RF_misc",
              "N2O_emissions", "N2O_natural_emissions"] + ["%s_emissions" % h for h in HALOS]
PARAMETERS = ["S", "diff", "qco2", "beta", "q10_rh", "f_nppv", "f_nppd", "f_litterd", "npp_flux0",
              "C0", "veg_c", "detritus_c", "soil_c", "permafrost_c", "warmingfactor",
              "rh_ch4_frac", "pf_mu", "pf_sigma", "fpf_static", "tt", "tu", "twi", "tid",
              "preind_surface_c", "preind_interdeep_c", "eps_abs", "eps_rel", "dt", "eps_spinup",
              "aero_scalar", "vol_scalar", "delta_co2", "delta_ch4", "delta_n2o", "rho_bc",
              "rho_oc", "rho_so2", "rho_nh3", "M0", "Tsoil", "Tstrat", "UC_CH4", "TOH0", "CNOX",
              "CCO", "CNMVOC", "CCH4", "PO3", "N0", "lo_warming_ratio"]
OUTPUT_VARIABLES = ["CO2_concentration", "global_tas", "RF_tot", "RF_CO2", "heatflux", "ocean_c",
                    "HL_pH", "atmos_co2", "sst", "permafrost_c", "CH4_concentration",
                    "N2O_concentration", "O3_concentration", "land_tas", "veg_c", "detritus_c",
                    "soil_c", "thawedp_c", "earth_c", "NBP", "ocean_uptake", "LL_pH", "HL_PCO2",
                    "LL_PCO2", "HL_ocean_c", "LL_ocean_c", "IO_ocean_c", "DO_ocean_c", "RF_CH4",
                    "RF_N2O", "rh_ch4", "NPP", "RH", "gmst", "ocean_tas", "heatflux_mixed",
                    "heatflux_interior", "ocean_timesteps"]
MEMBER_STATUS = {0: "ok", 1: "negative flux/pool", 2: "mass not conserved",
                 3: "solver retries exhausted", 4: "no [H+] root", 5: "yearfraction out of bounds",
                 6: "CO2 SARF condition", 7: "ODE stepper", 8: "spin-up did not converge",
                 9: "tracking fractions out of range",
                 10: "needs exact_attempts=True (a skipped ODE attempt could have gone negative)"}
# carbon tracking: tracked pools (fluxpool names) and the possible source names
TRACK_POOLS = ["atmos_co2", "earth_c", "veg_c", "detritus_c", "soil_c", "permafrost_c",
               "thawedp_c", "HL", "LL", "intermediate", "deep"]
TRACK_SOURCES = TRACK_POOLS + ["untracked"]
TRACK_COMPONENT = ["simpleNbox"] * 7 + ["ocean"] * 4
# output variable holding each tracked pool's total (Pg C)
TRACK_POOL_OUTPUT = ["atmos_co2", "earth_c", "veg_c", "detritus_c", "soil_c", "permafrost_c",
                     "thawedp_c", "HL_ocean_c", "LL_ocean_c", "IO_ocean_c", "DO_ocean_c"]


# units and owning component of every recorded variable, as the reference reports them
# (unitval::unitsName, inst/include/unitval.hpp:68-130; getComponentName of the component that
# registers the capability)
VARIABLE_UNITS = {
    "CO2_concentration": "ppmv CO2", "global_tas": "degC", "RF_tot": "W/m2", "RF_CO2": "W/m2",
    "heatflux": "W/m2", "ocean_c": "Pg C", "HL_pH": "pH", "atmos_co2": "Pg C", "sst": "degC",
    "permafrost_c": "Pg C", "CH4_concentration": "ppbv CH4", "N2O_concentration": "ppbv N2O",
    "O3_concentration": "DU O3", "land_tas": "degC", "veg_c": "Pg C", "detritus_c": "Pg C",
    "soil_c": "Pg C", "thawedp_c": "Pg C", "earth_c": "Pg C", "NBP": "Pg C/yr",
    "ocean_uptake": "Pg C/yr", "LL_pH": "pH", "HL_PCO2": "uatm", "LL_PCO2": "uatm",
    "HL_ocean_c": "Pg C", "LL_ocean_c": "Pg C", "IO_ocean_c": "Pg C", "DO_ocean_c": "Pg C",
    "RF_CH4": "W/m2", "RF_N2O": "W/m2", "rh_ch4": "Pg C/yr", "ocean_timesteps": "(unitless)",
    "NPP": "Pg C/yr", "RH": "Pg C/yr", "gmst": "degC", "ocean_tas": "degC",
    "heatflux_mixed": "W/m2", "heatflux_interior": "W/m2",
    # functions of recorded outputs, evaluated at fetch time (FUNCTION_VARIABLES)
    "HL_sst": "degC", "LL_sst": "degC", "HL_DIC": "umol/kg", "LL_DIC": "umol/kg", "DIC": "umol/kg",
    "pH": "pH", "PCO2": "uatm", "ML_ocean_c": "Pg C", "TAU_OH": "Years", "f_frozen": "(unitless)",
    "HL_CO3": "umol/kg", "LL_CO3": "umol/kg", "CO3": "umol/kg",
    "HL_ocean_uptake": "Pg C/yr", "LL_ocean_uptake": "Pg C/yr", "rh_det": "Pg C/yr", "rh_soil": "Pg C/yr"}
VARIABLE_COMPONENT = {
    "HL_sst": "ocean", "LL_sst": "ocean", "HL_DIC": "ocean", "LL_DIC": "ocean", "DIC": "ocean",
    "pH": "ocean", "PCO2": "ocean", "ML_ocean_c": "ocean", "TAU_OH": "OH", "f_frozen": "simpleNbox",
    "HL_CO3": "ocean", "LL_CO3": "ocean", "CO3": "ocean", "HL_ocean_uptake": "ocean",
    "LL_ocean_uptake": "ocean", "rh_det": "simpleNbox", "rh_soil": "simpleNbox",
    "CO2_concentration": "simpleNbox", "atmos_co2": "simpleNbox", "veg_c": "simpleNbox",
    "detritus_c": "simpleNbox", "soil_c": "simpleNbox", "permafrost_c": "simpleNbox",
    "thawedp_c": "simpleNbox", "earth_c": "simpleNbox", "NBP": "simpleNbox",
    "rh_ch4": "simpleNbox", "global_tas": "temperature", "land_tas": "temperature",
    "sst": "temperature", "heatflux": "temperature", "RF_tot": "forcing", "RF_CO2": "forcing",
    "RF_CH4": "forcing", "RF_N2O": "forcing", "ocean_c": "ocean", "ocean_uptake": "ocean",
    "HL_pH": "ocean", "LL_pH": "ocean", "HL_PCO2": "ocean", "LL_PCO2": "ocean",
    "HL_ocean_c": "ocean", "LL_ocean_c": "ocean", "IO_ocean_c": "ocean", "DO_ocean_c": "ocean",
    "ocean_timesteps": "ocean", "NPP": "simpleNbox", "RH": "simpleNbox", "gmst": "temperature",
    "ocean_tas": "temperature", "heatflux_mixed": "temperature",
    "heatflux_interior": "temperature", "CH4_concentration": "CH4", "N2O_concentration": "N2O",
    "O3_concentration": "ozone"}


# Outputs that need no kernel: hx_fetch derives them from the scenario series, per-member
# parameters and recorded outputs (RF_O3_trop needs O3_concentration, RF_H2O_strat needs
# CH4_concentration among the selected outputs)
# per-biome inputs, set as "<biome>.<name>" once biomes are defined (simpleNbox.cpp:281-396)
BIOME_PARAMETERS = ["veg_c", "detritus_c", "soil_c", "permafrost_c", "npp_flux0", "beta", "q10_rh",
                    "warmingfactor", "f_nppv", "f_nppd", "f_litterd", "rh_ch4_frac", "pf_mu",
                    "pf_sigma", "fpf_static"]

# per-biome outputs, selected and fetched as "<biome>.<name>" (simpleNbox.cpp:533-697)
BIOME_OUTPUTS = ["veg_c", "detritus_c", "soil_c", "permafrost_c", "thawedp_c", "NPP", "RH"]

# Per-stash quantities the run kernel records on request (scratch rows in global memory, the
# all-output builds only; rh_det / rh_soil summed over the biomes): selected and fetched like
# OUTPUT_VARIABLES
STASH_OUTPUTS = ["HL_ocean_uptake", "LL_ocean_uptake", "rh_det", "rh_soil"]

# The rest of the reference's outputstream variables that are plain functions of recorded outputs
# (hx_fetch evaluates them on the host; each needs the outputs it depends on to be selected):
# {variable: recorded outputs it needs}.  With OUTPUT_VARIABLES, STASH_OUTPUTS and
# DERIVED_VARIABLES this is all of R's ALL_VARS().  With biomes f_frozen needs every
# "<biome>.permafrost_c" (it is their mean weighted with the current permafrost,
# simpleNbox.cpp:492-513) and "<biome>.f_frozen" that biome's.
FUNCTION_VARIABLES = {
    "HL_sst": ["sst"], "LL_sst": ["sst"], "HL_DIC": ["HL_ocean_c"], "LL_DIC": ["LL_ocean_c"],
    "DIC": ["HL_ocean_c", "LL_ocean_c"], "pH": ["HL_pH", "LL_pH"], "PCO2": ["HL_PCO2", "LL_PCO2"],
    "ML_ocean_c": ["HL_ocean_c", "LL_ocean_c"], "TAU_OH": ["CH4_concentration"],
    "f_frozen": ["land_tas", "permafrost_c"],
    "HL_CO3": ["sst", "HL_PCO2", "HL_pH"], "LL_CO3": ["sst", "LL_PCO2", "LL_pH"],
    "CO3": ["sst", "HL_PCO2", "HL_pH", "LL_PCO2", "LL_pH"],
    # not in R's ALL_VARS(), but the reference's getData and outputstream know them
    "HL_OmegaCa": ["sst", "HL_PCO2", "HL_pH"], "LL_OmegaCa": ["sst", "LL_PCO2", "LL_pH"],
    "HL_OmegaAr": ["sst", "HL_PCO2", "HL_pH"], "LL_OmegaAr": ["sst", "LL_PCO2", "LL_pH"]}

DERIVED_VARIABLES = (["RF_BC", "RF_OC", "RF_SO2", "RF_NH3", "RF_aci", "RF_vol", "RF_albedo",
                      "RF_misc", "RF_O3_trop", "RF_H2O_strat"]
                     + ["RF_%s" % h for h in HALOS] + ["Fadj%s" % h for h in HALOS]
                     + ["%s_concentration" % h for h in HALOS])


def variable_units(v):
    if v in VARIABLE_UNITS:
        return VARIABLE_UNITS[v]
    if "." in v and v.split(".", 1)[1] in BIOME_OUTPUTS:  # <biome>.<name>
        return VARIABLE_UNITS[v.split(".", 1)[1]]
    if v.startswith("RF_") or v.startswith("Fadj"):
        return "W/m2"
    if v.endswith("_concentration"):
        return "pptv"
    return "(unitless)"


def load_scenario_tables(path):
    """{scenario: table[nrow, len(RAW_SERIES)]} from an .npz written by
    tests/golden/make_golden.py (columns in RAW_SERIES order)."""
    z = np.load(path)
    names = [str(s) for s in z["names"]]
    if names != RAW_SERIES:
        raise HxError("scenario file columns do not match RAW_SERIES")
    return {(k[:-5] + "-over" if k.endswith("_over") else k): np.ascontiguousarray(z[k])
            for k in z.files if k != "names"}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Ensemble:
    def __init__(self, n_members, scenarios, member_scenario=None, start_year=1745, end_year=2300,
                 device=0, outputs=("CO2_concentration", "global_tas"), cold_newton=False,
                 spinup=True, stream=None, tracking_date=None, track_every=1, biomes=None,
                 exact_attempts=False, keep_order=False):
        """scenarios: one table [nrow, 44] (RAW_SERIES columns), or a list of them;
        member_scenario: int array [n_members] of indices into that list;
        tracking_date: [core] trackingDate -- carbon tracking from that year on, recorded every
        `track_every` years and in the end year (0: end year only);
        exact_attempts: execute the ODE attempts the reference abandons whenever one could raise
        its negativity exception (HX_FLAG_EXACT_ATTEMPTS; without it such a member stops with
        status 10); keep_order: no internal re-ordering of the members (HX_FLAG_KEEP_ORDER);
        biomes: names of 2..4 biomes, in creation order, that replace the global one -- their
        pools and parameters are then set as "<biome>.<name>" (BIOME_PARAMETERS) before prepare."""
        self.L = _capi.lib()
        if isinstance(scenarios, np.ndarray):
            scenarios = [scenarios]
        self.n_members = int(n_members)
        self.start_year, self.end_year = int(start_year), int(end_year)
        flags = (_capi.HX_FLAG_COLD_NEWTON if cold_newton else 0) | \
                (0 if spinup else _capi.HX_FLAG_NO_SPINUP) | \
                (_capi.HX_FLAG_EXACT_ATTEMPTS if exact_attempts else 0) | \
                (_capi.HX_FLAG_KEEP_ORDER if keep_order else 0)
        cfg = _capi.HxConfig(self.n_members, len(scenarios), self.start_year, self.end_year,
                             int(device), flags)
        self.h = C.c_void_p()
        rc = self.L.hx_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise HxError("hx_create: %s" % self.L.hx_last_error(None).decode())
        if stream is not None:
            self._chk(self.L.hx_set_stream(self.h, C.c_void_p(int(stream))))
        names = (C.c_char_p * len(RAW_SERIES))(*[s.encode() for s in RAW_SERIES])
        nrow = self.end_year - self.start_year + 1
        for sid, tab in enumerate(scenarios):
            tab = np.ascontiguousarray(tab, dtype=np.float64)
            if tab.shape != (nrow, len(RAW_SERIES)):
                raise HxError("scenario table must be [%d, %d]" % (nrow, len(RAW_SERIES)))
            self._chk(self.L.hx_set_scenario_table(self.h, sid, len(RAW_SERIES), names,
                                                   self.start_year, nrow, _dp(tab)))
        if member_scenario is not None:
            ms = np.ascontiguousarray(member_scenario, dtype=np.int32)
            self._chk(self.L.hx_set_member_scenario(
                self.h, ms.ctypes.data_as(C.POINTER(C.c_int32)), len(ms)))
        self.biomes = list(biomes) if biomes else []
        if self.biomes:  # before the outputs: "<biome>.<name>" outputs name a known biome
            arr = (C.c_char_p * len(self.biomes))(*[b.encode() for b in self.biomes])
            self._chk(self.L.hx_set_biomes(self.h, len(self.biomes), arr))
        self.select_outputs(outputs)
        if tracking_date is not None:
            self._chk(self.L.hx_set_tracking(self.h, int(tracking_date), int(track_every)))
        self.prepared = False

    def select_outputs(self, outputs):
        """which variables the run records (before prepare): OUTPUT_VARIABLES names and, once
        biomes are defined, "<biome>.<name>" with name in BIOME_OUTPUTS"""
        self.outputs = list(outputs)
        arr = (C.c_char_p * len(self.outputs))(*[s.encode() for s in self.outputs])
        self._chk(self.L.hx_select_outputs(self.h, len(self.outputs), arr))

    def split_biome(self, new_biomes, fveg_c=None, fdetritus_c=None, fsoil_c=None,
                    fpermafrost_c=None, fnpp_flux0=None, **params):
        """R split_biome(core, "global", new_biomes, ...) (R/biome.R:61-131): the global biome's
        pools and initial NPP are distributed over `new_biomes` by the given fractions (default:
        evenly; the others default to fveg_c), every other parameter is inherited unless given
        in `params` (one value, or one per new biome).  Call before prepare()."""
        n = len(new_biomes)
        fveg_c = [1.0 / n] * n if fveg_c is None else list(fveg_c)
        fr = {"veg_c": fveg_c,
              "detritus_c": fveg_c if fdetritus_c is None else list(fdetritus_c),
              "soil_c": fveg_c if fsoil_c is None else list(fsoil_c),
              "permafrost_c": fveg_c if fpermafrost_c is None else list(fpermafrost_c),
              "npp_flux0": fveg_c if fnpp_flux0 is None else list(fnpp_flux0)}
        for k, f in fr.items():
            if len(f) != n or abs(sum(f) - 1.0) > 1e-12 or min(f) < 0 or \
                    (k != "permafrost_c" and min(f) <= 0):
                raise HxError("split_biome: the %s fractions must be positive and sum to 1" % k)
        unknown = set(params) - set(BIOME_PARAMETERS)
        if unknown:
            raise HxError("split_biome: not a biome parameter: %s" % sorted(unknown))
        cur = {k: self.getvar(k) for k in BIOME_PARAMETERS}  # the global biome's, per member
        arr = (C.c_char_p * n)(*[b.encode() for b in new_biomes])
        self._chk(self.L.hx_set_biomes(self.h, n, arr))
        self.biomes = list(new_biomes)
        self.select_outputs(self.outputs)  # a new biome list drops per-biome selections
        for i, b in enumerate(new_biomes):
            for k in BIOME_PARAMETERS:
                if k in fr:
                    v = cur[k] * fr[k][i]
                elif k in params:
                    p = params[k]
                    v = np.full(self.n_members, float(p[i] if np.ndim(p) else p))
                else:
                    v = cur[k]
                self.setvar("%s.%s" % (b, k), v)

    def set_biome(self, biome, **values):
        """<biome>.<name> inputs (scalars or per-member arrays), like setvar("boreal.beta", ...)"""
        for k, v in values.items():
            self.setvar("%s.%s" % (biome, k), v)

    @classmethod
    def from_ini(cls, ini_paths, n_members, member_scenario=None, device=0,
                 outputs=("CO2_concentration", "global_tas"), cold_newton=False, stream=None):
        """newcore(inifile): scenario series and scalar parameters come from Hector ini files
        (one scenario per file), read by the library's own ini/csv reader."""
        if isinstance(ini_paths, str):
            ini_paths = [ini_paths]
        self = cls.__new__(cls)
        self.L = _capi.lib()
        self.n_members = int(n_members)
        arr = (C.c_char_p * len(ini_paths))(*[p.encode() for p in ini_paths])
        self.h = C.c_void_p()
        flags = _capi.HX_FLAG_COLD_NEWTON if cold_newton else 0
        rc = self.L.hx_create_from_ini(arr, len(ini_paths), self.n_members, int(device), flags,
                                       C.byref(self.h))
        if rc != 0:
            raise HxError("hx_create_from_ini: %s (code %d)" % (
                self.L.hx_last_error(None).decode(), rc))
        if stream is not None:
            self._chk(self.L.hx_set_stream(self.h, C.c_void_p(int(stream))))
        if member_scenario is not None:
            ms = np.ascontiguousarray(member_scenario, dtype=np.int32)
            self._chk(self.L.hx_set_member_scenario(
                self.h, ms.ctypes.data_as(C.POINTER(C.c_int32)), len(ms)))
        self.outputs = list(outputs)
        oarr = (C.c_char_p * len(self.outputs))(*[s.encode() for s in self.outputs])
        self._chk(self.L.hx_select_outputs(self.h, len(self.outputs), oarr))
        self.prepared = False
        self.start_year = self.end_year = None
        return self

    # -- plumbing --
    def _chk(self, rc):
        if rc != 0:
            raise HxError("%s (code %d)" % (self.L.hx_last_error(self.h).decode(), rc))

    def close(self):
        if getattr(self, "h", None):
            self.L.hx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- reference-shaped surface --
    def setvar(self, name, values):
        """R setvar (R/messages.R:107-140): scalar for all members or one value per member."""
        if np.isscalar(values):
            self._chk(self.L.hx_set_param_scalar(self.h, name.encode(), float(values)))
        else:
            v = np.ascontiguousarray(values, dtype=np.float64)
            self._chk(self.L.hx_set_param(self.h, name.encode(), _dp(v), v.size))

    def setvar_member(self, name, member, value):
        """setvar addressed to ONE member (the others keep their values)"""
        self._chk(self.L.hx_set_param_member(self.h, name.encode(), int(member), float(value)))

    def setvar_series(self, name, years, values, scenario=0):
        """R setvar(core, dates, var, values): a dated input of one scenario -- an emissions
        series (must cover start..end) or a user constraint (CO2_constrain, NBP_constrain, tas_constrain,
        RF_tot_constrain, CH4_constrain, N2O_constrain, <gas>_constrain; any subset of years,
        tests/testthat/test_constraints.R).  After prepare() it takes effect at the next
        reset()/run(), like the reference's setvar + reset."""
        years = np.asarray(years, dtype=np.int64)
        values = np.asarray(values, dtype=np.float64)
        y0, y1 = int(years.min()), int(years.max())
        dense = np.full(y1 - y0 + 1, np.nan)
        dense[years - y0] = values
        self._chk(self.L.hx_set_scenario_series(self.h, int(scenario), name.encode(), y0,
                                                dense.size, _dp(dense)))

    def setvar_device(self, name, dev_ptr, n):
        self._chk(self.L.hx_set_param_device(self.h, name.encode(), C.c_void_p(int(dev_ptr)), n))

    def getvar(self, name):
        out = np.empty(self.n_members)
        self._chk(self.L.hx_get_param(self.h, name.encode(), _dp(out), out.size))
        return out

    def prepare(self):
        """Core::prepareToRun incl. spin-up."""
        self._chk(self.L.hx_prepare(self.h))
        self.prepared = True

    def run(self, to_date=-1):
        if not self.prepared:
            self.prepare()
        self._chk(self.L.hx_run(self.h, float(to_date)))

    def run_stream(self, variables=None, to_date=-1, outs=None, segments=4):
        """run(to_date) and fetch every year of the segment for `variables` in one call, the
        device-to-host copies overlapped with the computation.  Returns {variable: array
        [n_years, n_members]} (year-major, R fetchvars' long format); `outs` may supply
        preallocated (pinned) arrays of that shape."""
        if not self.prepared:
            self.prepare()
        variables = list(variables or self.outputs)
        first = int(self.current_date) + 1
        last = int(to_date) if to_date >= 0 else int(self._end_year())
        ny = last - first + 1
        if outs is None:
            outs = [np.empty((ny, self.n_members)) for _ in variables]
        names = (C.c_char_p * len(variables))(*[v.encode() for v in variables])
        ptrs = (C.c_void_p * len(variables))(*[
            (o.ctypes.data if hasattr(o, "ctypes") else int(o)) for o in outs])
        self._chk(self.L.hx_run_stream(self.h, float(to_date), len(variables), names, ptrs,
                                       int(segments)))
        return dict(zip(variables, outs))

    def _end_year(self):
        if self.end_year is not None:
            return self.end_year
        raise HxError("run_stream on an ini-built engine needs an explicit to_date")

    def reset(self, date=None):
        """R reset(core, date = 0): back to the start (re-running the spin-up if parameters or
        inputs changed), or to a year inside the run already made."""
        if date is None:
            self._chk(self.L.hx_reset(self.h))
        else:
            self._chk(self.L.hx_reset_date(self.h, float(date)))

    def synchronize(self):
        self._chk(self.L.hx_synchronize(self.h))

    def fetch(self, var, dates, out=None):
        """-> array [n_members, n_dates]; `out` may be a preallocated (pinned) buffer address
        holder: a numpy array or anything with .ctypes / an int address."""
        dates = np.ascontiguousarray(dates, dtype=np.float64)
        if out is None:
            out = np.empty((self.n_members, dates.size))
        addr = out.ctypes.data if hasattr(out, "ctypes") else int(out)
        self._chk(self.L.hx_fetch(self.h, var.encode(), _dp(dates), dates.size,
                                  C.c_void_p(addr)))
        return out

    def fetchvars(self, dates, variables=None):
        """R fetchvars (R/messages.R:46-88): {variable: array[n_members, n_dates]}"""
        variables = variables or self.outputs
        return {v: self.fetch(v, dates) for v in variables}

    def fetchvars_frame(self, dates, variables=None, members=None, scenario="ensemble"):
        """R fetchvars (R/messages.R:46-88, src/rcpp_hector.cpp:349-355) as a long data frame:
        columns scenario, member, year, variable, value, units (one row per member x year x
        variable; `member` is the column the single-core reference does not have)."""
        import pandas as pd
        variables = list(variables or self.outputs)
        dates = np.asarray(dates, dtype=np.float64)
        members = np.arange(self.n_members) if members is None else np.asarray(members)
        frames = []
        for v in variables:
            x = self.fetch(v, dates)[members]
            frames.append(pd.DataFrame({
                "scenario": scenario, "member": np.repeat(members, dates.size),
                "year": np.tile(dates.astype(int), members.size), "variable": v,
                "value": x.reshape(-1), "units": variable_units(v)}))
        return pd.concat(frames, ignore_index=True)

    def write_outputstream(self, path, member=0, run_name="hector_b200", dates=None,
                           variables=None):
        """One member as the reference's outputstream_<run>.csv (CSVOutputStreamVisitor,
        src/csv_outputstream_visitor.cpp:55-71, 100-140): a comment line, then
        year,run_name,spinup,component,variable,value,units rows, values with the C++ stream's
        default six significant digits."""
        variables = list(variables or self.outputs)
        if dates is None:
            dates = np.arange(self.start_year + 1, int(self.current_date) + 1)
        dates = np.asarray(dates, dtype=np.float64)
        cols = {v: self.fetch(v, dates)[member] for v in variables}
        with open(path, "w") as f:
            f.write("# Output from hector_b200 (Hector v3.5.0 hot path) member %d\n" % member)
            f.write("year,run_name,spinup,component,variable,value,units\n")
            for k, y in enumerate(dates.astype(int)):
                for v in variables:
                    f.write("%d,%s,0,%s,%s,%.6g,%s\n" % (
                        y, run_name, VARIABLE_COMPONENT.get(v, "?"), v, cols[v][k],
                        VARIABLE_UNITS.get(v, "(unitless)")))

    def output_device(self, var):
        """(device pointer, member stride, n_years) of the [year][member] block of `var`."""
        p = C.c_void_p()
        stride = C.c_int64()
        ny = C.c_int32()
        self._chk(self.L.hx_output_device(self.h, var.encode(), C.byref(p), C.byref(stride),
                                          C.byref(ny)))
        return p.value, stride.value, ny.value

    def fetch_tracking(self, date):
        """-> (frac[n_members, 11, 12], mask[n_members, 11]): source fractions of the tracked
        pools (TRACK_POOLS x TRACK_SOURCES) in year `date`, and which sources are keys of each
        pool's map (bit s of mask) -- the rows Core::getTrackingData prints."""
        frac = np.empty((self.n_members, len(TRACK_POOLS), len(TRACK_SOURCES)))
        mask = np.zeros((self.n_members, len(TRACK_POOLS)), dtype=np.uint32)
        self._chk(self.L.hx_fetch_tracking(self.h, float(date), _dp(frac),
                                           mask.ctypes.data_as(C.POINTER(C.c_uint32))))
        return frac, mask

    def tracking_years(self):
        n = self.L.hx_tracking_years(self.h, None, 0)
        if n < 0:
            self._chk(n)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.L.hx_tracking_years(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return out[:n]

    def tracking_data(self, member, dates):
        """get_tracking_data (R/messages.R, Core::getTrackingData, csv_tracking_visitor.cpp:88-103)
        for one member: rows (year, component, pool_name, pool_value, pool_units, source_name,
        source_fraction).  Pool totals need the pool outputs to be selected."""
        rows = []
        for y in dates:
            frac, mask = self.fetch_tracking(y)
            for k, pool in enumerate(TRACK_POOLS):
                val = float("nan")
                if TRACK_POOL_OUTPUT[k] in self.outputs:
                    val = float(self.fetch(TRACK_POOL_OUTPUT[k], [y])[member, 0])
                for s, src in enumerate(TRACK_SOURCES):
                    if mask[member, k] >> s & 1:
                        rows.append((int(y), TRACK_COMPONENT[k], pool, val, "Pg C", src,
                                     float(frac[member, k, s])))
        return rows

    # -- exchange between the GPUs of a node over peer memory (no kernels) --
    def ipc_export(self):
        """64-byte CUDA IPC handle of this engine's output block"""
        buf = C.create_string_buffer(64)
        self._chk(self.L.hx_ipc_export(self.h, buf, None))
        return buf.raw

    def ipc_open(self, handles, self_index):
        blob = b"".join(handles)
        self._chk(self.L.hx_ipc_open(self.h, len(handles), blob, int(self_index)))

    def ipc_pull(self, var, year_a, year_b, dst_dev_ptr):
        self._chk(self.L.hx_ipc_pull(self.h, var.encode(), int(year_a), int(year_b),
                                     C.c_void_p(int(dst_dev_ptr))))

    def ipc_wait(self):
        self._chk(self.L.hx_ipc_wait(self.h))

    # -- push exchange: finished slabs go to every peer's gather block while the run computes --
    def xchg_create(self, n_peers, self_index):
        """-> 64-byte CUDA IPC handle of this rank's gather block"""
        buf = C.create_string_buffer(64)
        self._chk(self.L.hx_xchg_create(self.h, int(n_peers), int(self_index), buf, None))
        return buf.raw

    def xchg_open(self, handles):
        self._chk(self.L.hx_xchg_open(self.h, len(handles), b"".join(handles)))

    def run_exchange(self, to_date=-1):
        if not self.prepared:
            self.prepare()
        self._chk(self.L.hx_run_exchange(self.h, float(to_date)))

    def xchg_block(self):
        """(device pointer, elements per rank slot) of the gather block"""
        p = C.c_void_p()
        n = C.c_int64()
        self._chk(self.L.hx_xchg_block(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def event_record(self, idx):
        self._chk(self.L.hx_event_record(self.h, int(idx)))

    def event_synchronize(self, idx):
        self._chk(self.L.hx_event_synchronize(self.h, int(idx)))

    def status(self):
        st = np.empty(self.n_members, dtype=np.int32)
        fy = np.empty(self.n_members, dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        self._chk(self.L.hx_member_status(self.h, st.ctypes.data_as(ip), fy.ctypes.data_as(ip),
                                          self.n_members))
        return st, fy

    def counters(self):
        out = (C.c_uint64 * _capi.HX_NCOUNTERS)()
        self._chk(self.L.hx_counters(self.h, out, _capi.HX_NCOUNTERS))
        return dict(zip(_capi.COUNTER_NAMES, [int(x) for x in out]))

    def spinup_state(self, member=0):
        out = np.empty(14)
        self._chk(self.L.hx_spinup_state(self.h, member, _dp(out)))
        keys = ["atmos", "veg", "det", "soil", "permafrost", "thawed", "earth", "HL", "LL", "IO",
                "DO", "alk_HL", "alk_LL", "spinup_steps"]
        return dict(zip(keys, out))

    @property
    def current_date(self):
        return self.L.hx_current_date(self.h)

    @property
    def last_run_ms(self):
        return self.L.hx_last_run_ms(self.h)
