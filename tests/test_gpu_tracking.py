"""GPU parity of carbon tracking: the engine's per-member source maps,
called through the C ABI, against the unmodified reference's committed known answers
(tests/golden/ref_tracking.npz) and against the CPU oracle on perturbed members.

Tolerance: the fractions are contractive mixtures (no error growth), observed agreement with the
reference is ~1e-15; the test holds 1e-12 absolute on fractions in [0, 1] and exact key sets."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

TOL_FRAC = 1e-12
# the extreme-corner member (S = 6, q10 = 3.5, beta = 1, diff = 3): its high-latitude surface box
# amplifies 1-ulp differences for decades (tests/test_gpu_parity.py), and the
# box fluxes are what mixes the source maps
TOL_FRAC_CORNER = 5e-8


@pytest.mark.parametrize("case", util.ref_tracking(), ids=lambda c: c["name"])
def test_tracking_vs_reference_golden(case):
    import hector_b200 as hb
    tab = util.scenarios()[case["scenario"]]
    ens = hb.Ensemble(3, tab, outputs=["CO2_concentration", "global_tas"] + hb.TRACK_POOL_OUTPUT,
                      tracking_date=case["tracking_date"], track_every=1)
    for k, v in case["params"].items():
        ens.setvar(k, v)
    ens.run()
    st, _ = ens.status()
    assert (st == 0).all()
    worst = 0.0
    for i, y in enumerate(case["years"]):
        frac, mask = ens.fetch_tracking(y)
        assert np.array_equal(mask[0], case["mask"][i]), (y, mask[0], case["mask"][i])
        assert np.array_equal(mask[0], mask[2]) and np.array_equal(frac[0], frac[2])
        worst = max(worst, float(np.abs(frac[0] - case["frac"][i]).max()))
    assert worst < (TOL_FRAC_CORNER if "corner" in case["name"] else TOL_FRAC), worst
    # years before the tracking date are not available
    with pytest.raises(hb.HxError):
        ens.fetch_tracking(case["tracking_date"] - 1)
    # pool totals that go with the fractions
    got = ens.fetchvars(case["years"].astype(np.float64), hb.TRACK_POOL_OUTPUT)
    for k, v in enumerate(hb.TRACK_POOL_OUTPUT):
        tol = TOL_FRAC_CORNER if "corner" in case["name"] else 1e-10
        assert util.parity_err(got[v][0], case["pool_values"][:, k], v) < tol, v
    ens.close()


def test_tracking_leaves_trajectories_bit_identical_and_matches_oracle():
    """tracking on leaves CO2/Tgav and all work counters bit-identical;
    perturbed members (BASELINE.json config 5 parameters) agree with the oracle's maps."""
    from oracle import port
    import hector_b200 as hb
    M = 40
    rng = np.random.Generator(np.random.PCG64(20241018))
    lo = np.array([2.0, 1.0, 0.2, 0.5, 0.5, 0.8])
    hi = np.array([5.0, 2.6, 0.9, 2.5, 1.5, 1.2])
    names = ["S", "q10_rh", "beta", "diff", "aero_scalar", "vol_scalar"]
    X = lo + rng.random((M, 6)) * (hi - lo)
    tab = util.scenarios()["ssp585"]
    outs = ["CO2_concentration", "global_tas", "ocean_timesteps"]

    def make(track):
        e = hb.Ensemble(M, tab, outputs=outs, tracking_date=1750 if track else None,
                        track_every=50)
        for j, n in enumerate(names):
            e.setvar(n, X[:, j])
        e.run()
        return e
    a, b = make(False), make(True)
    yrs = np.arange(1746, 2301, dtype=np.float64)
    ga, gb = a.fetchvars(yrs), b.fetchvars(yrs)
    for v in outs:
        assert np.array_equal(ga[v], gb[v]), v
    ca, cb = a.counters(), b.counters()
    assert ca == cb
    rec_years = list(range(1750, 2301, 50)) + [2300]
    for i in (0, 7, 39):
        p = port.default_params(**{n: X[i, j] for j, n in enumerate(names)})
        st, _, out, frac, mask = port.run_member_tracked(tab, 1750, p)
        assert st == 0
        for y in rec_years:
            f, k = b.fetch_tracking(y)
            assert np.array_equal(k[i], mask[y - 1746]), (i, y)
            assert np.abs(f[i] - frac[y - 1746]).max() < TOL_FRAC, (i, y)
    # a year that was not recorded is an error, not a silent interpolation
    with pytest.raises(hb.HxError):
        b.fetch_tracking(1777)
    # reset() restores the untracked initial maps; a second run reproduces the first bit for bit
    f1, k1 = b.fetch_tracking(2300)
    b.reset()
    b.run()
    f2, k2 = b.fetch_tracking(2300)
    assert np.array_equal(f1, f2) and np.array_equal(k1, k2)
    a.close()
    b.close()


def test_tracking_resume_and_live_fetch():
    """run(d1); run(d2) carries the maps; the current date is always fetchable even when it is
    not a recorded year; tracking_data() has the reference's row shape."""
    import hector_b200 as hb
    tab = util.scenarios()["ssp245"]
    outs = ["CO2_concentration"] + hb.TRACK_POOL_OUTPUT
    one = hb.Ensemble(2, tab, outputs=outs, tracking_date=1800, track_every=0)
    one.run()
    two = hb.Ensemble(2, tab, outputs=outs, tracking_date=1800, track_every=0)
    two.run(1799)
    with pytest.raises(hb.HxError):
        two.fetch_tracking(1799)
    two.run(1853)
    f_mid, k_mid = two.fetch_tracking(1853)       # live maps, not a recorded year
    assert abs(f_mid[0].sum(axis=1) - 1.0).max() < 1e-12
    two.run()
    fa, ka = one.fetch_tracking(2300)
    fb, kb = two.fetch_tracking(2300)
    assert np.array_equal(fa, fb) and np.array_equal(ka, kb)
    rows = two.tracking_data(0, [2300])
    assert rows[0][:3] == (2300, "simpleNbox", "atmos_co2") and rows[0][4] == "Pg C"
    assert {r[2] for r in rows} == set(hb.TRACK_POOLS)
    by_pool = {}
    for r in rows:
        by_pool.setdefault(r[2], 0.0)
        by_pool[r[2]] += r[6]
    assert all(abs(v - 1.0) < 1e-12 for v in by_pool.values())
    one.close()
    two.close()


def test_tracking_csv_matches_reference_text():
    """the reference's own getTrackingData() text (6 significant digits) for 1750-1755"""
    import hector_b200 as hb
    tab = util.scenarios()["ssp245"]
    ens = hb.Ensemble(1, tab, outputs=hb.TRACK_POOL_OUTPUT, tracking_date=1750)
    ens.run(1755)
    ref_rows = {}
    for line in util.ref_tracking_csv().splitlines()[1:]:
        y, comp, pool, val, units, src, fr = line.split(",")
        ref_rows[(int(y), comp, pool, src)] = (float(val), float(fr))
    rows = ens.tracking_data(0, range(1750, 1756))
    assert len(rows) == len(ref_rows)
    for y, comp, pool, val, units, src, fr in rows:
        rv, rf = ref_rows[(y, comp, pool, src)]
        assert abs(val - rv) <= 6e-6 * max(abs(rv), 1e-30) + 1e-12, (y, pool, val, rv)
        assert abs(fr - rf) <= 6e-6 * max(abs(rf), 1e-30) + 1e-12, (y, pool, src, fr, rf)
    ens.close()


def test_tracking_with_run_stream_and_reset_to_date():
    """the streaming run (segments of whole slabs) and reset(date) drive the same record/replay
    pipeline: identical maps"""
    import hector_b200 as hb
    tab = util.scenarios()["ssp245"]
    outs = ["CO2_concentration", "global_tas"]
    a = hb.Ensemble(130, tab, outputs=outs, tracking_date=1760, track_every=20)
    b = hb.Ensemble(130, tab, outputs=outs, tracking_date=1760, track_every=20)
    S = np.linspace(2.0, 5.0, 130)
    for e in (a, b):
        e.setvar("S", S)
    a.run()
    got = b.run_stream(outs, segments=5)
    ref = a.fetchvars(np.arange(1746, 2301, dtype=np.float64))
    for v in outs:
        assert np.array_equal(got[v], ref[v].T), v
    assert list(a.tracking_years()) == list(range(1760, 2301, 20))
    for y in (1760, 2000, 2300):
        fa, ka = a.fetch_tracking(y)
        fb, kb = b.fetch_tracking(y)
        assert np.array_equal(fa, fb) and np.array_equal(ka, kb), y
    b.reset(1900)                     # re-derives the state of 1900, maps included
    b.run()
    fb, kb = b.fetch_tracking(2300)
    fa, ka = a.fetch_tracking(2300)
    assert np.array_equal(fa, fb) and np.array_equal(ka, kb)
    a.close()
    b.close()
