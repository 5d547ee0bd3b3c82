"""ctypes binding of libhector_b200.so (include/hector_b200.h)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

HX_NCOUNTERS = 8
COUNTER_NAMES = ["rhs_evals", "rk_steps", "rk_rejected", "stashes", "newton_iterations",
                 "newton_calls", "failed_members", "member_years"]
HX_FLAG_COLD_NEWTON = 1
HX_FLAG_NO_SPINUP = 2
HX_FLAG_EXACT_ATTEMPTS = 4
HX_FLAG_KEEP_ORDER = 8

EXPORTS = ["hx_create", "hx_create_from_ini", "hx_ini_read", "hx_ini_scalar", "hx_ini_string", "hx_destroy", "hx_last_error", "hx_set_stream", "hx_set_scenario_series", "hx_set_param_member",
           "hx_set_scenario_table", "hx_set_member_scenario", "hx_set_param_scalar",
           "hx_set_param", "hx_set_param_device", "hx_get_param", "hx_select_outputs",
           "hx_prepare", "hx_run", "hx_run_stream", "hx_reset", "hx_reset_date", "hx_synchronize", "hx_fetch", "hx_output_device",
           "hx_ipc_export", "hx_ipc_open", "hx_ipc_pull", "hx_ipc_wait", "hx_ipc_close",
           "hx_xchg_create", "hx_xchg_open", "hx_run_exchange", "hx_xchg_block", "hx_xchg_close", "hx_event_record", "hx_event_synchronize", "hx_member_status", "hx_set_tracking", "hx_set_biomes", "hx_biome_count", "hx_biome_name", "hx_tracking_date", "hx_fetch_tracking", "hx_tracking_years", "hx_counters", "hx_current_date", "hx_last_run_ms",
           "hx_spinup_state", "hx_measure_fp64_peak", "hx_measure_hbm_copy", "hx_diag_transcendentals", "hx_version"]


class HxError(RuntimeError):
    pass


class HxConfig(C.Structure):
    _fields_ = [("n_members", C.c_int32), ("n_scenarios", C.c_int32), ("start_year", C.c_int32),
                ("end_year", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32)]


def lib_path():
    # HECTOR_B200_LIB: alternative build of the same library (tuning experiments)
    return os.environ.get("HECTOR_B200_LIB") or os.path.join(HERE, "libhector_b200.so")


def lib():
    """Load the CUDA library; fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise HxError("%s not found: build it with `python -c 'import __graft_entry__ as g; "
                      "g.build()'` or `make -C hector_b200/csrc` (needs nvcc, sm_100a)" % path)
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    L.hx_version.restype = C.c_char_p
    L.hx_last_error.restype = C.c_char_p
    L.hx_last_error.argtypes = [vp]
    L.hx_create.argtypes = [C.POINTER(HxConfig), C.POINTER(vp)]
    L.hx_create_from_ini.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.c_int32,
                                     C.c_uint32, C.POINTER(vp)]
    L.hx_ini_read.argtypes = [C.c_char_p, ip, ip, dp, C.c_int32]
    L.hx_ini_scalar.argtypes = [C.c_char_p, C.c_char_p, dp]
    L.hx_destroy.argtypes = [vp]
    L.hx_set_stream.argtypes = [vp, vp]
    L.hx_set_scenario_series.argtypes = [vp, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, dp]
    L.hx_set_scenario_table.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_char_p),
                                        C.c_int32, C.c_int32, dp]
    L.hx_set_member_scenario.argtypes = [vp, ip, C.c_int32]
    L.hx_set_param_scalar.argtypes = [vp, C.c_char_p, C.c_double]
    L.hx_set_param.argtypes = [vp, C.c_char_p, dp, C.c_int32]
    L.hx_set_param_member.argtypes = [vp, C.c_char_p, C.c_int32, C.c_double]
    L.hx_set_param_device.argtypes = [vp, C.c_char_p, vp, C.c_int32]
    L.hx_get_param.argtypes = [vp, C.c_char_p, dp, C.c_int32]
    L.hx_select_outputs.argtypes = [vp, C.c_int32, C.POINTER(C.c_char_p)]
    L.hx_prepare.argtypes = [vp]
    L.hx_run.argtypes = [vp, C.c_double]
    L.hx_run_stream.argtypes = [vp, C.c_double, C.c_int32, C.POINTER(C.c_char_p),
                                C.POINTER(C.c_void_p), C.c_int32]
    L.hx_reset.argtypes = [vp]
    L.hx_reset_date.argtypes = [vp, C.c_double]
    L.hx_synchronize.argtypes = [vp]
    L.hx_fetch.argtypes = [vp, C.c_char_p, dp, C.c_int32, vp]
    L.hx_output_device.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(C.c_int64), ip]
    L.hx_ipc_export.argtypes = [vp, vp, C.POINTER(C.c_int64)]
    L.hx_ipc_open.argtypes = [vp, C.c_int32, vp, C.c_int32]
    L.hx_ipc_pull.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32, vp]
    L.hx_ipc_wait.argtypes = [vp]
    L.hx_ipc_close.argtypes = [vp]
    L.hx_xchg_create.argtypes = [vp, C.c_int32, C.c_int32, vp, C.POINTER(C.c_int64)]
    L.hx_xchg_open.argtypes = [vp, C.c_int32, vp]
    L.hx_run_exchange.argtypes = [vp, C.c_double]
    L.hx_xchg_block.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64)]
    L.hx_xchg_close.argtypes = [vp]
    L.hx_event_record.argtypes = [vp, C.c_int32]
    L.hx_event_synchronize.argtypes = [vp, C.c_int32]
    L.hx_set_tracking.argtypes = [vp, C.c_int32, C.c_int32]
    L.hx_set_biomes.argtypes = [vp, C.c_int32, C.POINTER(C.c_char_p)]
    L.hx_biome_count.argtypes = [vp]
    L.hx_fetch_tracking.argtypes = [vp, C.c_double, dp, C.POINTER(C.c_uint32)]
    L.hx_tracking_years.argtypes = [vp, ip, C.c_int32]
    L.hx_member_status.argtypes = [vp, ip, ip, C.c_int32]
    L.hx_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int32]
    L.hx_current_date.argtypes = [vp]
    L.hx_current_date.restype = C.c_double
    L.hx_last_run_ms.argtypes = [vp]
    L.hx_last_run_ms.restype = C.c_double
    L.hx_spinup_state.argtypes = [vp, C.c_int32, dp]
    L.hx_measure_fp64_peak.argtypes = [C.c_int32, dp, dp]
    L.hx_measure_hbm_copy.argtypes = [C.c_int32, dp]
    if hasattr(L, "hx_diag_transcendentals"):  # absent from older A/B builds (tools/gpu_ab.sh)
        L.hx_diag_transcendentals.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, C.c_int32]
    _LIB = L
    return L
