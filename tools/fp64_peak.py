"""Measure the FP64 FMA peak and the device copy bandwidth of GPU 0 (hx_measure_fp64_peak /
hx_measure_hbm_copy, hector_b200/csrc/hx_diag.cu) -> JSON on stdout and, with a path, a file.
   python tools/fp64_peak.py [profiles/fp64_peak.json]"""
import ctypes as C, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hector_b200 import _capi
L = _capi.lib()
tf, per = C.c_double(), C.c_double()
assert L.hx_measure_fp64_peak(0, C.byref(tf), C.byref(per)) == 0
gb = C.c_double()
assert L.hx_measure_hbm_copy(0, C.byref(gb)) == 0
name = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.max.sm", "--format=csv,noheader"],
                      capture_output=True, text=True).stdout.strip()
out = {"fp64_fma_tflops": tf.value, "fma_per_clk_per_sm_at_nominal_clock": per.value,
       "hbm_copy_gbs": gb.value, "gpu": name,
       "method": "8 independent dependent-FMA chains per thread, 256-thread CTAs, full occupancy, "
                 "CUDA events, best of 3; copy = 1 GiB device-to-device, read + write"}
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
