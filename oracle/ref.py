"""oracle/ref.py -- ctypes binding of oracle/_ref/libhector_ref.so (TEST INFRASTRUCTURE ONLY).

The .so is the UNMODIFIED reference C++ (JGCRI/hector v3.5.0) compiled by oracle/Makefile
against the header-only Boost shim in oracle/shim.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libhector_ref.so")
REF_INPUT = os.path.join(REF_DIR, "input")

_lib = None

# variable -> owning component for ini-style overrides (ref_setdata), cf. inst/input/*.ini
PARAM_COMPONENT = {
    "S": "temperature", "diff": "temperature", "qco2": "temperature",
    "lo_warming_ratio": "temperature",
    "beta": "simpleNbox", "q10_rh": "simpleNbox", "f_nppv": "simpleNbox", "f_nppd": "simpleNbox",
    "f_litterd": "simpleNbox", "npp_flux0": "simpleNbox", "C0": "simpleNbox",
    "aero_scalar": "forcing", "vol_scalar": "forcing",
    "delta_co2": "forcing", "delta_ch4": "forcing", "delta_n2o": "forcing",
    "endDate": "core", "startDate": "core", "trackingDate": "core",
}


# tracked pools (fluxpool names) and the universe of source names, fixed order
TRACK_POOLS = ["atmos_co2", "earth_c", "veg_c", "detritus_c", "soil_c", "permafrost_c",
               "thawedp_c", "HL", "LL", "intermediate", "deep"]
TRACK_SOURCES = TRACK_POOLS + ["untracked"]


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libhector_ref.so missing: run `make -C oracle ref` "
                               "(needs /root/reference)")
        L = C.CDLL(REF_SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_open.argtypes = [C.c_char_p]
        L.ref_setdata.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_double, C.c_char_p]
        L.ref_setvar.argtypes = [C.c_int, C.c_char_p, C.c_double, C.c_double, C.c_char_p]
        L.ref_prepare.argtypes = [C.c_int]
        L.ref_run.argtypes = [C.c_int, C.c_double]
        L.ref_reset.argtypes = [C.c_int, C.c_double]
        L.ref_fetch.argtypes = [C.c_int, C.c_char_p, C.c_double, C.POINTER(C.c_double)]
        L.ref_fetch_series.argtypes = [C.c_int, C.c_char_p, C.c_double, C.c_int,
                                       C.POINTER(C.c_double)]
        L.ref_fetch_component.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_double,
                                          C.POINTER(C.c_double)]
        for f in (L.ref_start_date, L.ref_end_date, L.ref_current_date):
            f.argtypes = [C.c_int]
            f.restype = C.c_double
        L.ref_close.argtypes = [C.c_int]
        L.ref_tracking_pool.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_char_p),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(C.c_int)]
        L.ref_tracking_data.argtypes = [C.c_int, C.c_char_p, C.c_long]
        L.ref_tracking_data.restype = C.c_long
        L.ref_counters.argtypes = [C.POINTER(C.c_uint64), C.c_int]
        L.ref_run_member.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p),
                                     C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.c_int,
                                     C.POINTER(C.c_char_p), C.c_double, C.POINTER(C.c_double),
                                     C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


class RefError(RuntimeError):
    pass


def ini_path(scenario="ssp245"):
    return os.path.join(REF_INPUT, "hector_%s.ini" % scenario)


class RefCore:
    """One reference Core (newcore / setvar / run / fetchvars, R/hector.R:57-87)."""

    def __init__(self, ini):
        self.L = lib()
        self.h = self.L.ref_open(ini.encode())
        if self.h < 0:
            raise RefError(self.L.ref_last_error().decode())

    def _chk(self, rc):
        if rc != 0:
            raise RefError(self.L.ref_last_error().decode())

    def setdata(self, component, var, value, date=-1.0):
        v = value if isinstance(value, str) else "%.17g" % value
        self._chk(self.L.ref_setdata(self.h, component.encode(), var.encode(), date, v.encode()))

    def setvar(self, var, value, unit, date=-1.0):
        self._chk(self.L.ref_setvar(self.h, var.encode(), date, value, unit.encode()))

    def prepare(self):
        self._chk(self.L.ref_prepare(self.h))

    def run(self, to_date=-1.0):
        self._chk(self.L.ref_run(self.h, to_date))

    def reset(self, date):
        self._chk(self.L.ref_reset(self.h, date))

    def fetch(self, var, date=-1.0):
        out = C.c_double()
        self._chk(self.L.ref_fetch(self.h, var.encode(), date, C.byref(out)))
        return out.value

    def fetch_series(self, var, date0, n):
        out = np.empty(n)
        self._chk(self.L.ref_fetch_series(self.h, var.encode(), date0, n,
                                          out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def fetch_component(self, component, var, date=-1.0):
        out = C.c_double()
        self._chk(self.L.ref_fetch_component(self.h, component.encode(), var.encode(), date,
                                             C.byref(out)))
        return out.value

    def tracking_state(self):
        """-> (tracking?, value[11], frac[11, 12], present[11, 12]) at the current date, pools
        in TRACK_POOLS order, sources in TRACK_SOURCES order (full precision)."""
        ns = len(TRACK_SOURCES)
        srcs = (C.c_char_p * ns)(*[s.encode() for s in TRACK_SOURCES])
        val = np.zeros(len(TRACK_POOLS))
        frac = np.zeros((len(TRACK_POOLS), ns))
        pres = np.zeros((len(TRACK_POOLS), ns), dtype=np.int32)
        on = True
        for i, p in enumerate(TRACK_POOLS):
            v = C.c_double()
            rc = self.L.ref_tracking_pool(self.h, p.encode(), ns, srcs, C.byref(v),
                                          frac[i].ctypes.data_as(C.POINTER(C.c_double)),
                                          pres[i].ctypes.data_as(C.POINTER(C.c_int)))
            if rc < 0:
                raise RefError(self.L.ref_last_error().decode())
            on = on and rc == 1
            val[i] = v.value
        return on, val, frac, pres

    def tracking_csv(self):
        """Core::getTrackingData() as text"""
        n = self.L.ref_tracking_data(self.h, None, 0)
        if n < 0:
            raise RefError(self.L.ref_last_error().decode())
        buf = C.create_string_buffer(n + 1)
        self.L.ref_tracking_data(self.h, buf, n + 1)
        return buf.value.decode()

    @property
    def start_date(self):
        return self.L.ref_start_date(self.h)

    @property
    def end_date(self):
        return self.L.ref_end_date(self.h)

    def close(self):
        if self.h >= 0:
            self.L.ref_close(self.h)
            self.h = -1

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def counters(reset=False):
    out = (C.c_uint64 * 6)()
    lib().ref_counters(out, int(reset))
    keys = ["rhs_evals", "steps_accepted", "steps_rejected", "integrate_calls",
            "newton_iterations", "newton_calls"]
    return dict(zip(keys, [int(x) for x in out]))


def run_member(ini, params=None, variables=("CO2_concentration", "global_tas"), to_date=-1.0,
               nyears=None):
    """Run one member start->to_date; returns (ok, err, out[nvars+1, nyears], run_seconds).
    Row nvars holds the per-year ocean sub-step (stash) count.  params: {name: value} using
    PARAM_COMPONENT, or {(component, name): value}."""
    L = lib()
    params = params or {}
    comps, names, vals = [], [], []
    for k, v in params.items():
        if isinstance(k, tuple):
            comps.append(k[0]); names.append(k[1])
        else:
            comps.append(PARAM_COMPONENT[k]); names.append(k)
        vals.append(float(v))
    n = len(vals)
    c_comps = (C.c_char_p * max(n, 1))(*[s.encode() for s in comps])
    c_names = (C.c_char_p * max(n, 1))(*[s.encode() for s in names])
    c_vals = (C.c_double * max(n, 1))(*vals)
    nv = len(variables)
    c_vars = (C.c_char_p * max(nv, 1))(*[s.encode() for s in variables])
    if nyears is None:
        nyears = 555 if to_date < 0 else int(to_date - 1745)
    out = np.full((nv + 1, nyears), np.nan)
    secs = C.c_double(0.0)
    rc = L.ref_run_member(ini.encode(), n, c_comps, c_names, c_vals, nv, c_vars, to_date,
                          out.ctypes.data_as(C.POINTER(C.c_double)) if nv else None, nyears,
                          C.byref(secs))
    err = L.ref_last_error().decode() if rc else ""
    return rc == 0, err, out, secs.value
