"""CPU only: how far does the oracle move under perturbations that are legitimate between two
correct IEEE builds?  For the members of the all-parameter draw (tests/util.allparams_draw) on
which the GPU differs most from the oracle, compare

  (a) the oracle as built (gcc -O2 -ffp-contract=off),
  (b) the same source built with FMA contraction (-O2 -mfma -ffp-contract=fast) -- what nvcc does
      to the device code by default,
  (c) the oracle with ONE input (diff) moved by one ulp,

on every output, with the parity metric of tests/util.parity_err.  If |a - b| or |a - c| on a
variable is of the size of the GPU-vs-oracle difference on the same member, that difference is
conditioning of the model, not an error of the kernel.

usage: python tools/conditioning_probe.py [M seed member ...]   (default: 64 5 5 39)"""
import ctypes, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import port
from tests import util

args = [int(a) for a in sys.argv[1:]] or [64, 5, 5, 39]
M, seed, members = args[0], args[1], args[2:]
scen = util.scenarios()["ssp370"]
vals = util.allparams_draw(M, seed, port.default_params())


def run(i, ulp_on=None):
    kw = {util.ALLPARAM_RANGES[n][0]: float(vals[n][i]) for n in vals}
    if ulp_on:
        kw[ulp_on] = float(np.nextafter(kw[ulp_on], np.inf))
    st, fy, out, cnt, _ = port.run_member(scen, **kw)
    assert st == 0
    return out, cnt


base = {i: run(i) for i in members}
ulp = {i: run(i, "diff") for i in members}
# (b): rebuild the oracle with contraction into a scratch directory and rebind port to it
tmp = tempfile.mkdtemp()
so = os.path.join(tmp, "libhector_oracle_fma.so")
subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-mfma", "-ffp-contract=fast",
                       "-o", so, os.path.join(ROOT, "oracle", "hector_oracle.c"), "-lm"])
port.SO, port._lib = so, None
fma = {i: run(i) for i in members}

for i in members:
    print("member %d  (sub-step counts equal under fma: %s, under 1 ulp: %s)" % (
        i, np.array_equal(base[i][0][-1], fma[i][0][-1]), np.array_equal(base[i][0][-1], ulp[i][0][-1])))
    rows = []
    for k, v in enumerate(port.OUT_NAMES[:-1]):
        e_f = util.parity_err(fma[i][0][k], base[i][0][k], v)
        e_u = util.parity_err(ulp[i][0][k], base[i][0][k], v)
        rows.append((max(e_f, e_u), v, e_f, e_u))
    for _, v, e_f, e_u in sorted(rows, reverse=True)[:10]:
        print("   %-20s fma %.3g   one-ulp(diff) %.3g" % (v, e_f, e_u))
