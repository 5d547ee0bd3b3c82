"""CPU probe behind the replay kernel's mixing form (hx_model.cuh, tm_mix).

The reference mixes two source maps as (a fd + b fs) / (a + b) per source (fluxpool.hpp:197-257).
The replay kernel evaluates the same mean as fd + v (fs - fd) with v = b (1 / (a + b)) -- two
FP64 operations per source instead of six.  This script measures what that costs in parity: it
builds a copy of the oracle's C restatement whose tm_add uses the kernel's form (the one
statement is rewritten with a regular expression; nothing else changes), runs the tracked
SSP5-8.5 configuration held to 2500 (BASELINE.json configs[4] shape) for a few Monte-Carlo
members with both libraries and reports the largest difference in any source fraction, the key
sets, the drift of the fractions' sums from 1 and the largest fraction.

Result (this container): max |difference| 7e-15 over 755 years, key sets identical, sums within
4e-15 of 1 for both forms, largest fraction exactly 1.

usage: python tools/track_lerp_probe.py [members]   (test infrastructure: uses oracle/)"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import port
from tests import util

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
src = open(os.path.join(ROOT, "oracle", "hector_oracle.c")).read()
exact_stmt = "A->f[s] = new_total ? pool / new_total : 1.0 / n;"
assert exact_stmt in src
kernel_stmt = ("A->f[s] = new_total ? __builtin_fma(b * (1.0 / new_total), B->f[s] - A->f[s], A->f[s])"
               " : 1.0 / n;")
tmp = tempfile.mkdtemp()
with open(os.path.join(tmp, "ho_lerp.c"), "w") as f:
    f.write(src.replace(exact_stmt, kernel_stmt))
subprocess.check_call(["cp", os.path.join(ROOT, "oracle", "hector_oracle.h"), tmp])
so = os.path.join(tmp, "libho_lerp.so")
subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                       os.path.join(tmp, "ho_lerp.c"), "-lm"])
exact = port.lib()
lerp = C.CDLL(so)
lerp.ho_run_member_tracked.argtypes = exact.ho_run_member_tracked.argtypes
lerp.ho_run_member_tracked.restype = C.c_int

raw = util.scenarios()["ssp585"]
ext = np.vstack([raw, np.repeat(raw[-1:], 200, axis=0)])
rng = np.random.Generator(np.random.PCG64(20241018))
lo = np.array([2.0, 1.0, 0.2, 0.5, 0.5, 0.8])
hi = np.array([5.0, 2.6, 0.9, 2.5, 1.5, 1.2])
X = lo + rng.random((96, 6)) * (hi - lo)
names = ["S", "q10_rh", "beta", "diff", "aero_scalar", "vol_scalar"]
worst = 0.0
for i in range(0, 8 * n, 8):
    p = port.default_params(end_year=2500, **{k: X[i, j] for j, k in enumerate(names)})
    st, _, _, frac, mask = port.run_member_tracked(ext, 1750, p)
    port._lib = lerp
    try:
        st2, _, _, frac2, mask2 = port.run_member_tracked(ext, 1750, p)
    finally:
        port._lib = exact
    assert st == 0 and st2 == 0 and np.array_equal(mask, mask2)
    t0 = 1750 - 1746
    d = float(np.abs(frac[t0:] - frac2[t0:]).max())
    print("member %2d  max |df| %.2e  |sum - 1| exact %.2e kernel form %.2e  largest fraction %.17g"
          % (i, d, np.abs(frac[t0:].sum(axis=2) - 1).max(), np.abs(frac2[t0:].sum(axis=2) - 1).max(),
             frac2[t0:].max()))
    worst = max(worst, d)
print("worst difference in any fraction:", worst)
assert worst < 1e-13
