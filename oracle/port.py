"""oracle/port.py -- ctypes binding of oracle/libhector_oracle.so, the plain-C restatement
(TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke(), bench.py cpu_baseline)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhector_oracle.so")
NHALO = 26
NRAW = 18 + NHALO

OUT_NAMES = ["CO2_concentration", "global_tas", "RF_tot", "RF_CO2", "heatflux", "ocean_c", "HL_pH",
             "atmos_co2", "sst", "permafrost_c", "CH4_concentration", "N2O_concentration",
             "O3_concentration", "land_tas", "veg_c", "detritus_c", "soil_c", "thawedp_c",
             "earth_c", "NBP", "ocean_uptake", "LL_pH", "HL_PCO2", "LL_PCO2", "HL_ocean_c",
             "LL_ocean_c", "IO_ocean_c", "DO_ocean_c", "RF_CH4", "RF_N2O", "rh_ch4", "NPP", "RH",
             "gmst", "ocean_tas", "heatflux_mixed", "heatflux_interior", "ocean_timesteps"]
NOUT = len(OUT_NAMES)
STATUS = {0: "OK", 1: "NEGATIVE", 2: "MASS", 3: "RETRIES", 4: "NOROOT", 5: "YEARFRACTION",
          6: "CO2SARF", 7: "STEPPER", 8: "TRACKING"}
TRACK_POOLS = ["atmos_co2", "earth_c", "veg_c", "detritus_c", "soil_c", "permafrost_c",
               "thawedp_c", "HL", "LL", "intermediate", "deep"]
TRACK_SOURCES = TRACK_POOLS + ["untracked"]


MAX_BIOMES = 4
BIOME_FIELDS = ("veg_c", "detritus_c", "soil_c", "permafrost_c", "npp_flux0", "beta", "q10_rh",
                "warmingfactor", "f_nppv", "f_nppd", "f_litterd", "rh_ch4_frac", "pf_mu",
                "pf_sigma", "fpf_static")
BIOME_DEFAULTS = dict(warmingfactor=1.0, rh_ch4_frac=0.023, pf_mu=1.67, pf_sigma=0.986,
                      fpf_static=0.74)  # simpleNbox-runtime.cpp:109-143


class Biome(C.Structure):
    _fields_ = [(n, C.c_double) for n in BIOME_FIELDS]


class Params(C.Structure):
    _fields_ = (
        [(n, C.c_int) for n in ("start_year", "end_year", "do_spinup", "max_spinup")]
        + [(n, C.c_double) for n in (
            "S", "diff", "qco2", "beta", "q10_rh", "f_nppv", "f_nppd", "f_litterd", "npp_flux0",
            "C0", "veg_c", "detritus_c", "soil_c", "permafrost_c", "warmingfactor", "rh_ch4_frac",
            "pf_mu", "pf_sigma", "fpf_static", "tt", "tu", "twi", "tid", "preind_C_surface",
            "preind_C_ID")]
        + [("spinup_chem", C.c_int)]
        + [(n, C.c_double) for n in (
            "eps_abs", "eps_rel", "dt", "eps_spinup", "baseyear", "aero_scalar", "vol_scalar",
            "delta_co2", "delta_ch4", "delta_n2o", "rho_bc", "rho_oc", "rho_so2", "rho_nh3",
            "M0", "Tsoil", "Tstrat", "UC_CH4", "TOH0", "CNOX", "CCO", "CNMVOC", "CCH4", "PO3",
            "N0", "UC_N2O", "TN2O0")]
        + [(n, C.c_double * NHALO) for n in (
            "halo_tau", "halo_rho", "halo_delta", "halo_H0", "halo_molarMass")]
        + [("lo_warming_ratio", C.c_double), ("n_biomes", C.c_int),
           ("biome_order", C.c_int * MAX_BIOMES), ("biome", Biome * MAX_BIOMES)]
    )

    def set_biomes(self, biomes):
        """biomes: {name: {field: value}} in creation (biome_list) order; every pool and
        parameter the reference insists on (simpleNbox-runtime.cpp:66-101) must be given, the
        rest default like its prepareToRun does (:102-143)"""
        names = list(biomes)
        assert 2 <= len(names) <= MAX_BIOMES
        self.n_biomes = len(names)
        for k, i in enumerate(sorted(range(len(names)), key=lambda i: names[i])):
            self.biome_order[k] = i
        for i, n in enumerate(names):
            vals = dict(BIOME_DEFAULTS)
            vals.update(biomes[n])
            assert set(vals) == set(BIOME_FIELDS), (n, set(BIOME_FIELDS) ^ set(vals))
            for f, v in vals.items():
                setattr(self.biome[i], f, float(v))
        return self


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "rhs_evals", "steps_accepted", "steps_rejected", "integrate_calls", "newton_iterations",
        "newton_calls", "spinup_steps")]


class SpinState(C.Structure):
    _fields_ = ([(n, C.c_double) for n in ("atmos", "veg", "det", "soil", "permafrost", "thawed",
                                           "earth")]
                + [("ocean", C.c_double * 4), ("alk_HL", C.c_double), ("alk_LL", C.c_double),
                   ("spinup_steps", C.c_int)])


class Constraints(C.Structure):
    _fields_ = ([(n, C.POINTER(C.c_double)) for n in ("co2", "nbp", "ch4", "n2o", "halo", "rf_tot",
                                                      "tas")]
                + [(n, C.c_int) for n in ("rf_tot_last_row", "tas_first_row", "tas_last_row")])


HALOS = ["CF4", "C2F6", "HFC23", "HFC32", "HFC4310", "HFC125", "HFC134a", "HFC143a", "HFC227ea",
         "HFC245fa", "SF6", "CFC11", "CFC12", "CFC113", "CFC114", "CFC115", "CCl4", "CH3CCl3",
         "HCFC22", "HCFC141b", "HCFC142b", "halon1211", "halon1301", "halon2402", "CH3Cl",
         "CH3Br"]


def make_constraints(nrow, start_year, spec):
    """spec: {name: {year: value}} with names CO2_constrain, NBP_constrain, CH4_constrain,
    N2O_constrain, <gas>_constrain, RF_tot_constrain, tas_constrain.  Returns (Constraints,
    keep-alive list).  RF_tot / tas are densified the way tseries::get does it (linear
    interpolation between entries; RF_tot flat below its first entry)."""
    cn = Constraints()
    keep = []

    def dense(d):
        a = np.full(nrow, np.nan)
        for y, v in d.items():
            a[int(y) - start_year] = v
        return a

    def interp(d, flat_below):
        ys = np.array(sorted(d))
        vs = np.array([d[y] for y in ys], dtype=np.float64)
        a = np.full(nrow, np.nan)
        for r in range(nrow):
            y = start_year + r
            if y in d:
                a[r] = d[y]
            elif y < ys[0]:
                if flat_below:
                    a[r] = vs[0]
            elif y <= ys[-1]:
                i = np.searchsorted(ys, y) - 1
                a[r] = vs[i] + (y - ys[i]) * (vs[i + 1] - vs[i]) / (ys[i + 1] - ys[i])
        return a, int(ys[0]) - start_year, int(ys[-1]) - start_year

    halo = None
    for name, d in spec.items():
        if name == "RF_tot_constrain":
            a, _, last = interp(d, True)
            cn.rf_tot = _dp(a); cn.rf_tot_last_row = last; keep.append(a)
        elif name == "tas_constrain":
            a, first, last = interp(d, False)
            cn.tas = _dp(a); cn.tas_first_row = first; cn.tas_last_row = last; keep.append(a)
        elif name in ("CO2_constrain", "NBP_constrain", "CH4_constrain", "N2O_constrain"):
            a = dense(d)
            setattr(cn, name.split("_")[0].lower(), _dp(a)); keep.append(a)
        elif name.endswith("_constrain") and name[:-10] in HALOS:
            if halo is None:
                halo = np.full((nrow, NHALO), np.nan)
            halo[:, HALOS.index(name[:-10])] = dense(d)
        else:
            raise KeyError(name)
    if halo is not None:
        cn.halo = _dp(halo); keep.append(halo)
    return cn, keep


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        L.ho_default_params.argtypes = [C.POINTER(Params)]
        L.ho_run_member.argtypes = [C.POINTER(Params), C.POINTER(C.c_double), C.c_int,
                                    C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int),
                                    C.POINTER(Counters), C.POINTER(SpinState)]
        L.ho_run_member.restype = C.c_int
        L.ho_run_member_tracked.argtypes = L.ho_run_member.argtypes + [
            C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
        L.ho_run_member_tracked.restype = C.c_int
        L.ho_run_member_ex.argtypes = [C.POINTER(Params), C.POINTER(C.c_double),
                                       C.POINTER(Constraints), C.c_int, C.POINTER(C.c_double),
                                       C.c_int, C.POINTER(C.c_int), C.POINTER(Counters),
                                       C.POINTER(SpinState), C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_uint32)]
        L.ho_run_member_ex.restype = C.c_int
        L.ho_csys.argtypes = [C.c_double] * 6 + [C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ho_csys.restype = C.c_double
        L.ho_gas_series.argtypes = [C.POINTER(Params), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ho_doeclim_kernel.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def default_params(**over):
    p = Params()
    lib().ho_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def run_member(raw, params=None, run_to=-1, **over):
    """-> (status, fail_year, out[NOUT, nyears], counters dict, spin-up state dict)"""
    p = params if params is not None else default_params()
    for k, v in over.items():
        setattr(p, k, v)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    assert raw.shape == (p.end_year - p.start_year + 1, NRAW), raw.shape
    ny = (p.end_year if run_to < 0 else run_to) - p.start_year
    out = np.empty((NOUT, ny))
    fy = C.c_int(0)
    cnt = Counters()
    sp = SpinState()
    st = lib().ho_run_member(C.byref(p), _dp(raw), run_to, _dp(out), ny, C.byref(fy),
                             C.byref(cnt), C.byref(sp))
    cd = {n: int(getattr(cnt, n)) for n, _ in Counters._fields_}
    sd = {n: getattr(sp, n) for n in ("atmos", "veg", "det", "soil", "permafrost", "thawed",
                                      "earth", "alk_HL", "alk_LL", "spinup_steps")}
    sd["ocean"] = list(sp.ocean)
    return st, fy.value, out, cd, sd


def run_member_constrained(raw, spec, params=None, run_to=-1, tracking_date=None, **over):
    """run_member with user constraints (see make_constraints) -> (status, fail_year, out), or
    (status, fail_year, out, frac, mask) when carbon tracking is requested as well"""
    p = params if params is not None else default_params()
    for k, v in over.items():
        setattr(p, k, v)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    nrow = p.end_year - p.start_year + 1
    assert raw.shape == (nrow, NRAW), raw.shape
    cn, keep = make_constraints(nrow, p.start_year, spec)
    ny = (p.end_year if run_to < 0 else run_to) - p.start_year
    out = np.empty((NOUT, ny))
    fy = C.c_int(0)
    if tracking_date is None:
        st = lib().ho_run_member_ex(C.byref(p), _dp(raw), C.byref(cn), run_to, _dp(out), ny,
                                    C.byref(fy), None, None, 9999, None, None)
        return st, fy.value, out
    frac = np.empty((ny, len(TRACK_POOLS), len(TRACK_SOURCES)))
    mask = np.zeros((ny, len(TRACK_POOLS)), dtype=np.uint32)
    st = lib().ho_run_member_ex(C.byref(p), _dp(raw), C.byref(cn), run_to, _dp(out), ny,
                                C.byref(fy), None, None, int(tracking_date), _dp(frac),
                                mask.ctypes.data_as(C.POINTER(C.c_uint32)))
    return st, fy.value, out, frac, mask


BIOME_OUT_NAMES = ["veg_c", "detritus_c", "soil_c", "permafrost_c", "thawedp_c", "NPP", "RH"]


def run_member_biomes(raw, params, spec=None, run_to=-1, **over):
    """a multi-biome run (params.set_biomes) -> (status, fail_year, out, bio_out[nb, 7, nyears]);
    bio_out[b, k] = BIOME_OUT_NAMES[k] of biome b in creation order"""
    p = params
    for k, v in over.items():
        setattr(p, k, v)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    nrow = p.end_year - p.start_year + 1
    cn, keep = make_constraints(nrow, p.start_year, spec or {})
    ny = (p.end_year if run_to < 0 else run_to) - p.start_year
    out = np.empty((NOUT, ny))
    bio = np.full((max(p.n_biomes, 1), len(BIOME_OUT_NAMES), ny), np.nan)
    fy = C.c_int(0)
    f = lib().ho_run_member_biomes
    f.restype = C.c_int
    st = f(C.byref(p), _dp(raw), C.byref(cn), run_to, _dp(out), ny, C.byref(fy), _dp(bio))
    return st, fy.value, out, bio


def run_member_tracked(raw, tracking_date, params=None, run_to=-1, **over):
    """run_member with carbon tracking from `tracking_date`:
    -> (status, fail_year, out, frac[nyears, 11, 12] (NaN before tracking), mask[nyears, 11])"""
    p = params if params is not None else default_params()
    for k, v in over.items():
        setattr(p, k, v)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    assert raw.shape == (p.end_year - p.start_year + 1, NRAW), raw.shape
    ny = (p.end_year if run_to < 0 else run_to) - p.start_year
    out = np.empty((NOUT, ny))
    frac = np.empty((ny, len(TRACK_POOLS), len(TRACK_SOURCES)))
    mask = np.zeros((ny, len(TRACK_POOLS)), dtype=np.uint32)
    fy = C.c_int(0)
    st = lib().ho_run_member_tracked(C.byref(p), _dp(raw), run_to, _dp(out), ny, C.byref(fy),
                                     None, None, int(tracking_date), _dp(frac),
                                     mask.ctypes.data_as(C.POINTER(C.c_uint32)))
    return st, fy.value, out, frac, mask


def csys(Tbox, carbon, alk, volume, S=34.5, U=6.7):
    out = np.empty(8)
    it = C.c_int(0)
    h = lib().ho_csys(Tbox, carbon, alk, volume, S, U, _dp(out), C.byref(it))
    return h, out, it.value


def gas_series(raw, params=None):
    p = params if params is not None else default_params()
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    n = raw.shape[0]
    n2o = np.empty(n)
    hrf = np.empty((n, NHALO))
    lib().ho_gas_series(C.byref(p), _dp(raw), _dp(n2o), _dp(hrf))
    return n2o, hrf


def doeclim_kernel(diff, ns):
    k = np.empty(ns)
    lib().ho_doeclim_kernel(diff, ns, _dp(k))
    return k
