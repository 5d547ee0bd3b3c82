"""Random multi-biome configurations: the oracle against the UNMODIFIED reference (oracle/_ref),
bit for bit, including the year a run fails in.  Needs /root/reference (build container only).

usage: python tools/sweep_biomes_vs_ref.py [n_cases] [seed]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import port, ref
from tests import util
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import make_golden as mg

N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
SCN = ["ssp119", "ssp126", "ssp245", "ssp370", "ssp434", "ssp460", "ssp534-over", "ssp585"]
NAMES = ["tundra", "amazon", "midlat", "boreal", "steppe", "alpine"]
tmp = tempfile.mkdtemp()
bad = 0
for case in range(N):
    nb = int(rng.integers(2, 5))
    names = list(rng.permutation(NAMES)[:nb])
    fr = rng.dirichlet(np.ones(nb) * 3.0)
    over = {}
    for b in names:
        over[b] = dict(beta=rng.uniform(0.1, 0.9), q10_rh=rng.uniform(1.0, 2.8),
                       warmingfactor=rng.uniform(0.6, 2.4), f_nppv=rng.uniform(0.25, 0.45),
                       f_nppd=rng.uniform(0.4, 0.55), f_litterd=rng.uniform(0.9, 1.0),
                       rh_ch4_frac=rng.uniform(0.0, 0.06), pf_mu=rng.uniform(1.2, 2.2),
                       pf_sigma=rng.uniform(0.7, 1.3), fpf_static=rng.uniform(0.5, 0.9))
    biomes = mg._split(dict(zip(names, fr)), **over)
    pf = rng.dirichlet(np.ones(nb))
    if rng.random() < 0.4:
        pf[int(rng.integers(nb))] = 0.0
        pf = pf / pf.sum()
    for b, f in zip(names, pf):
        biomes[b]["permafrost_c"] = 865.0 * f
    scn = SCN[int(rng.integers(len(SCN)))]
    params = dict(S=rng.uniform(1.8, 5.0), diff=rng.uniform(0.5, 2.5))
    if rng.random() < 0.3:
        params["lo_warming_ratio"] = rng.uniform(0.9, 1.8)
    ini = os.path.join(tmp, "c%d.ini" % case)
    mg.biome_ini(scn, biomes, ini)
    own = ["%s.%s" % (b, v) for b in names for v in mg.BIOME_OWN_VARS]
    ok, err, o, _ = ref.run_member(ini, params, mg.BIOME_VARS + own)
    fail = 0 if ok else 1746 + int(np.argmax(np.isnan(o[0])))
    p = port.default_params()
    p.set_biomes(biomes)
    st, fy, out, bio = port.run_member_biomes(util.scenarios()[scn], p, **params)
    n = 555 if not fail else fail - 1746
    same = (fail == (fy if st else 0))
    for k, v in enumerate(mg.BIOME_VARS):
        if v in port.OUT_NAMES:
            same = same and np.array_equal(out[port.OUT_NAMES.index(v)][:n], o[k][:n])
    same = same and np.array_equal(out[-1][:n], o[-1][:n])
    for ib in range(nb):
        for k in range(len(mg.BIOME_OWN_VARS)):
            same = same and np.array_equal(bio[ib, k][:n], o[len(mg.BIOME_VARS) + ib * 7 + k][:n])
    bad += not same
    print("case %2d %-11s %d biomes  ref %s  oracle status %d  %s" % (
        case, scn, nb, "ok" if ok else "fails %d" % fail, st, "BIT-IDENTICAL" if same else "MISMATCH"))
print("mismatches:", bad, "of", N)
