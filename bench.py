#!/usr/bin/env python
"""bench.py -- throughput of the Hector ensemble hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU path

Metric (BASELINE.json): ensemble-member-years/s, SSP2-4.5, 1745->2300 (555 yearly steps per
member).  One "step" = one complete run of the whole ensemble (reset to the post-spin-up
state + 555 coupled years for every member).  Workload at N GPUs: --members per GPU
(default 65 536 = BASELINE.json configs[2], "1 GPU, HBM-bound sizing"; weak scaling), the
(S, q10_rh, beta, diff) Latin hypercube of SURVEY.md section 8(d), seed 20241017.

Numbers on the JSON line:
  value     member-years/s with parameters already resident in HBM (timed: state restore + run
            kernel [+ the final NCCL all-gather of CO2/Tgav trajectories when N > 1])
  e2e       same metric through the public API with HOST buffers: per step the 4 parameter
            vectors go host->device, set-up + spin-up + run execute, and the CO2 and Tgav
            trajectories of every member come back to (pinned) host memory
  roofline  algorithmic bytes (5 048 B per member-year) / run-kernel
            time measured with CUDA events, against the measured HBM copy bandwidth
  cpu_baseline  the unmodified reference (oracle/_ref) or the C oracle port on this box's host
            cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

YEARS = 555
E2E_VARS = ["CO2_concentration", "global_tas"]
B_ALG = 5048.0  # algorithmic bytes per member-year, SURVEY.md section 8(d)
METRIC = "ensemble_member_years_per_sec"
UNIT = "member-years/s"
PARAMS = ["S", "q10_rh", "beta", "diff"]
LO = np.array([2.0, 1.0, 0.2, 0.5])
HI = np.array([5.0, 2.6, 0.9, 2.5])


def lhs(M, seed=20241017):
    rng = np.random.Generator(np.random.PCG64(seed))
    cols = []
    for j in range(4):
        u = (rng.permutation(M) + rng.random(M)) / M
        cols.append(LO[j] + u * (HI[j] - LO[j]))
    return np.stack(cols, axis=1)


def scenario_table(name="ssp245"):
    import hector_b200 as hb
    return hb.load_scenario_tables(os.path.join(ROOT, "tests", "golden", "scenarios.npz"))[name]


# ------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for k, nm in ((5, "hw_slowdown"), (6, "hw_thermal_slowdown"),
                          (7, "sw_thermal_slowdown"), (8, "sw_power_cap")):
                if len(r) > k and r[k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU baseline: the reference's own implementation on the host cores
def _ref_worker(args):
    kind, rows, to_date = args
    secs, years = 0.0, 0
    if kind == "reference":
        from oracle import ref
        for S, q10, beta, diff in rows:
            ok, err, _, s = ref.run_member(ref.ini_path("ssp245"),
                                           dict(S=S, q10_rh=q10, beta=beta, diff=diff), (),
                                           to_date)
            secs += s
            years += YEARS
    else:
        from oracle import port
        raw = scenario_table()
        for S, q10, beta, diff in rows:
            t0 = time.perf_counter()
            port.run_member(raw, S=S, q10_rh=q10, beta=beta, diff=diff)
            secs += time.perf_counter() - t0   # includes the spin-up: the port has one entry
            years += YEARS
    return secs, years


def cpu_reference_pass(members_per_core=2, cores=None):
    """One bounded sample: `cores` processes x members_per_core member-runs each; returns
    (member-years/s over run() time, kind, cores, sample description)."""
    import multiprocessing as mp
    from oracle import ref
    kind = "reference" if ref.available() else "port"
    cores = cores or os.cpu_count() or 1
    if kind == "port":
        members_per_core = max(members_per_core, 64)
    X = lhs(cores * members_per_core)
    chunks = [(kind, [tuple(r) for r in X[i::cores]], -1.0) for i in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, chunks)
    worst = max(s for s, _ in res)
    years = sum(y for _, y in res)
    value = years / worst
    what = ("%d member-runs (first %d of the LHS) x 555 yr, %d processes, timing %s" %
            (len(X), len(X), cores,
             "Core::run() only (ini parse + spin-up excluded)" if kind == "reference"
             else "ho_run_member (spin-up included)"))
    return value, kind, cores, what


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, kind, cores, what = cpu_reference_pass(args.ref_members_per_core)
        if i >= args.warmup:
            vals.append((v, time.perf_counter() - t0))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([t for _, t in vals])) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "%d-member-per-GPU (S, q10_rh, beta, diff) Latin-hypercube ensemble, "
                        "SSP2-4.5, 1745->2300 (BASELINE.json configs[2] sizing)" % args.members,
            "members_per_gpu": args.members, "years": YEARS, "scenario": "ssp245",
            "sampler": "LHS seed 20241017",
            "cache": "working set (state + histories + outputs) >> 126 MB L2"
                     if args.members >= 16384 else "L2 flushed between timed iterations"}


# ------------------------------------------------------------------------------------------
def roofline_of(kernel, members, years, kernel_ms, hbm_peak, peaks_measured, fp64_peak,
                traffic_file, counts_file):
    """The three fractions of one run-kernel launch:
      frac       SURVEY.md 8(d)'s ALGORITHMIC bytes (5 048 B per member-year of a year-stepped SoA
                 model) / live kernel time, against the measured HBM copy bandwidth -- how fast
                 the model is advanced, NOT how busy HBM is;
      dram_frac  real DRAM traffic of the launch (ncu dram__bytes_read + write, profiles/) /
                 live kernel time, against the same bandwidth -- how busy HBM really is;
      fp64       executed FP64 flops (ncu smsp__sass_thread_inst_executed_op_{dfma x2, dmul,
                 dadd}, per member-year, profiles/) / live kernel time, against the DFMA peak
                 measured on this GPU right now (hx_measure_fp64_peak)."""
    achieved = members * years * B_ALG / (kernel_ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
         "frac": achieved / hbm_peak, "traffic": None,
         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks_measured
         else "B200_PROFILING.md fallback (of fallback)",
         "kernel": kernel, "kernel_ms": kernel_ms,
         "algorithmic_bytes_per_member_year": B_ALG,
         "binds": "FP64 issue/latency at low occupancy, not HBM: see dram_frac and fp64",
         "frac_note": "frac counts SURVEY 8(d)'s algorithmic bytes (a year-stepped model re-reading "
                      "its DOECLIM history every year); the kernel reads the history once per 16 "
                      "years and keeps the state on chip, so frac can exceed 1 while HBM is idle"}
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))
        # the capture's ensemble may differ in size from this run's: scale per member-year
        traffic = t["dram_bytes_per_launch"] * (members * years) / float(
            t.get("member_years", 65536 * 555))
        r["traffic"] = traffic
        r["dram_gbs"] = traffic / (kernel_ms * 1e-3) / 1e9
        r["dram_frac"] = r["dram_gbs"] / hbm_peak
        r["traffic_source"] = t.get("source")
    except Exception:
        pass
    try:
        c = json.load(open(os.path.join(ROOT, "profiles", counts_file)))
        flops = (2.0 * c["dfma_per_member_year"] + c["dmul_per_member_year"]
                 + c["dadd_per_member_year"]) * members * years
        tf = flops / (kernel_ms * 1e-3) / 1e12
        r["fp64"] = {"achieved_tflops": tf, "peak": fp64_peak,
                     "frac": tf / fp64_peak if fp64_peak else None, "unit": "TFLOP/s",
                     "flops_per_member_year": flops / (members * years),
                     "peak_source": "hx_measure_fp64_peak on this GPU, this run (dependent-FMA "
                                    "chains, all SMs, CUDA events)",
                     "counts_source": c.get("source")}
    except Exception:
        pass
    return r


def parity_spot(ens, X, n=8, seed=7):
    """Outside the timed region: CO2 / Tgav of n random members of the ensemble just run against
    the CPU oracle (the checker; SURVEY.md 8(d) parity metric).  A regression that kept every
    member's status at 0 would otherwise still print a throughput."""
    from oracle import port
    table = scenario_table()
    pick = sorted(np.random.default_rng(seed).choice(X.shape[0], n, replace=False).tolist())
    years = np.arange(1746, 2301, dtype=np.float64)
    co2 = ens.fetch("CO2_concentration", years)
    tas = ens.fetch("global_tas", years)
    e_co2 = e_tas = 0.0
    for i in pick:
        st, _, out, _, _ = port.run_member(table, S=X[i, 0], q10_rh=X[i, 1], beta=X[i, 2],
                                           diff=X[i, 3])
        if st != 0:
            continue
        e_co2 = max(e_co2, float(np.max(np.abs(co2[i] - out[0]) / np.maximum(np.abs(out[0]), 1.0))))
        e_tas = max(e_tas, float(np.max(np.abs(tas[i] - out[1]) / np.maximum(np.abs(out[1]), 0.01))))
    return {"members": pick, "CO2_concentration": e_co2, "global_tas": e_tas, "tolerance": 1e-10,
            "ok": bool(e_co2 <= 1e-10 and e_tas <= 1e-10),
            "metric": "max_t |x - oracle| / max(|oracle|, floor), floor 1 ppm / 0.01 degC"}


class CudaArrayView:
    """zero-copy view of an engine output block for torch (NCCL gathers)"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (ptr, False),
                                         "version": 2, "strides": None}


SSPS = ["ssp119", "ssp126", "ssp245", "ssp370", "ssp434", "ssp460", "ssp534-over", "ssp585"]


def run_config4(total, rank, world, local_rank, stream, dist, gloo, timed, args):
    """BASELINE.json configs[3]: `total` members, all 8 SSP scenarios interleaved in API order
    (member i runs scenario i mod 8), sharded as SURVEY.md 8(e) says: members sorted by scenario,
    the sorted list cut into `world` contiguous ranges.  Exchange: pushed over peer memory when
    available (hx_run_exchange), else one NCCL all-gather.  Afterwards rank 0 un-permutes a random
    subset of API members out of the gathered block and compares it with (a) a single-GPU run of
    exactly those members and (b) the CPU oracle."""
    import torch
    import hector_b200 as hb
    from hector_b200.sharding import scenario_sorted_shards, PushExchange
    ms = np.arange(total) % 8
    X = lhs(total)
    order, bounds = scenario_sorted_shards(ms, world)
    lo, hi = bounds[rank]
    if any(b - a != hi - lo for a, b in bounds):
        raise SystemExit("bench.py --config4-members: not divisible by the number of ranks")
    mine = order[lo:hi]
    scen_ids = sorted(set(ms[mine].tolist()))
    for sid in scen_ids:
        if int((ms[mine] == sid).sum()) % 128:
            raise SystemExit("bench.py --config4-members: per-rank scenario groups must be whole "
                             "128-member tiles")
    remap = {sid: k for k, sid in enumerate(scen_ids)}
    ens = hb.Ensemble(len(mine), [scenario_table(SSPS[sid]) for sid in scen_ids],
                      member_scenario=np.array([remap[int(x)] for x in ms[mine]], dtype=np.int32),
                      device=local_rank, outputs=E2E_VARS, stream=stream.cuda_stream)
    for j, nme in enumerate(PARAMS):
        ens.setvar(nme, np.ascontiguousarray(X[mine, j]))
    ens.prepare()
    ens.synchronize()
    _, stride, ny = ens.output_device(E2E_VARS[0])
    assert stride == len(mine)
    push = None
    gathered = None
    if world > 1:
        ok = 1
        try:
            if gloo is None:
                raise RuntimeError("no gloo group")
            push = PushExchange(ens, gloo)
        except Exception as ex:
            sys.stderr.write("bench.py config4: push exchange unavailable (%r), NCCL all-gather\n" % (ex,))
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            push = None
            gathered = torch.empty((world, 2, ny, stride), dtype=torch.float64, device="cuda")
    views = []
    for v in E2E_VARS:
        ptr, _, _ = ens.output_device(v)
        views.append(torch.as_tensor(CudaArrayView(ptr, (ny, stride)), device="cuda"))

    def step(e):
        e.reset()
        if world == 1:
            e.run()
        elif push is not None:
            push.run()
        else:
            e.run()
            mineblk = torch.stack(views)                       # [2, ny, stride]
            dist.all_gather_into_tensor(gathered.reshape(-1), mineblk.reshape(-1))

    nsteps = max(2, args.steps // 2)
    ms_total, _ = timed(ens, step, nsteps, 2, False)
    ms_step = ms_total / nsteps
    st, _ = ens.status()
    failed = torch.tensor([int((st != 0).sum())], dtype=torch.int64, device="cuda")
    if dist:
        dist.all_reduce(failed)
    res = None
    if rank == 0:
        block = (push.block if push is not None else gathered) if world > 1 else torch.stack(views)[None]
        inv = np.argsort(order)
        pick = np.sort(np.random.default_rng(4).choice(total, 64, replace=False))
        pos = inv[pick]
        rk, col = pos // stride, pos % stride
        got = block[torch.as_tensor(rk, device="cuda"), :, :, torch.as_tensor(col, device="cuda")]
        got = got.cpu().numpy()                                # [64, 2, ny]
        # (a) the same members on one GPU, in API order
        one = hb.Ensemble(64, [scenario_table(n) for n in SSPS], member_scenario=ms[pick].astype(np.int32),
                          device=local_rank, outputs=E2E_VARS)
        for j, nme in enumerate(PARAMS):
            one.setvar(nme, np.ascontiguousarray(X[pick, j]))
        one.run()
        years = np.arange(1746, 2301, dtype=np.float64)
        ref = np.stack([one.fetch(v, years) for v in E2E_VARS], axis=1)   # [64, 2, ny]
        one.close()
        same = bool(np.array_equal(got, ref))
        maxrel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 0.01)))
        # (b) the oracle on eight of them (one per scenario)
        from oracle import port
        e_co2 = e_tas = 0.0
        for k in range(8):
            i = int(np.nonzero(ms[pick] == k)[0][0]) if (ms[pick] == k).any() else k
            stt, _, out, _, _ = port.run_member(scenario_table(SSPS[int(ms[pick][i])]), S=X[pick[i], 0],
                                                q10_rh=X[pick[i], 1], beta=X[pick[i], 2], diff=X[pick[i], 3])
            e_co2 = max(e_co2, float(np.max(np.abs(got[i, 0] - out[0]) / np.maximum(np.abs(out[0]), 1.0))))
            e_tas = max(e_tas, float(np.max(np.abs(got[i, 1] - out[1]) / np.maximum(np.abs(out[1]), 0.01))))
        res = {"members": total, "members_per_gpu": len(mine), "scenarios": 8,
               "scenarios_on_rank0": [SSPS[s_] for s_ in scen_ids],
               "value": total * YEARS / (ms_step * 1e-3), "unit": UNIT, "ms_per_step": ms_step,
               "failed_members": int(failed.item()),
               "exchange": "none (1 GPU)" if world == 1 else ("peer memory, pushed" if push is not None
                                                                else "NCCL all-gather"),
               "gathered_bytes_per_gpu": int(world * 2 * ny * stride * 8),
               "unpermute_check": {"api_members": 64, "bit_identical_to_single_gpu_run": same,
                                   "max_rel_diff": maxrel},
               "parity_spot": {"members": 8, "CO2_concentration": e_co2, "global_tas": e_tas,
                               "ok": bool(e_co2 <= 1e-10 and e_tas <= 1e-10)},
               "note": "BASELINE.json configs[3]: member i runs scenario i mod 8 (API order); "
                       "members sorted by scenario and cut into contiguous per-GPU ranges "
                       "(SURVEY.md 8(e)); every GPU ends with all members' CO2 + Tgav"}
    if dist:
        dist.barrier()
    ens.close()
    return res


def run_config5(total, rank, world, local_rank, stream, dist, timed, args):
    """BASELINE.json configs[4]: `total` members, SSP5-8.5 with every series held at its 2300 value
    to 2500, carbon tracking from 1750 (maps recorded in 2500), Monte-Carlo over (S, q10_rh, beta,
    diff, aero_scalar, vol_scalar), seed 20241018; contiguous shards; ONE NCCL all-gather of the
    CO2 + Tgav trajectories after the run, and an all-reduced summary of the tracked maps."""
    import torch
    import hector_b200 as hb
    from hector_b200.sharding import shard_range
    raw = scenario_table("ssp585")
    ext = np.vstack([raw, np.repeat(raw[-1:], 200, axis=0)])
    rng = np.random.Generator(np.random.PCG64(20241018))
    lo6 = np.array([2.0, 1.0, 0.2, 0.5, 0.5, 0.8]); hi6 = np.array([5.0, 2.6, 0.9, 2.5, 1.5, 1.2])
    X6 = lo6 + rng.random((total, 6)) * (hi6 - lo6)
    a, b = shard_range(total, rank, world)
    if (b - a) * world != total or (b - a) % 128:
        raise SystemExit("bench.py --config5-members: shards must be equal and whole tiles")
    free0 = torch.cuda.mem_get_info()[0]
    ens = hb.Ensemble(b - a, ext, end_year=2500, device=local_rank, outputs=E2E_VARS,
                      tracking_date=1750, track_every=0, stream=stream.cuda_stream)
    names6 = PARAMS + ["aero_scalar", "vol_scalar"]
    for j, nme in enumerate(names6):
        ens.setvar(nme, np.ascontiguousarray(X6[a:b, j]))
    ens.prepare()
    ens.synchronize()
    engine_bytes = free0 - torch.cuda.mem_get_info()[0]
    _, stride, ny = ens.output_device(E2E_VARS[0])
    views = []
    for v in E2E_VARS:
        ptr, _, _ = ens.output_device(v)
        views.append(torch.as_tensor(CudaArrayView(ptr, (ny, stride)), device="cuda"))
    gathered = torch.empty((world, 2, ny, stride), dtype=torch.float64, device="cuda") if world > 1 else None

    def step(e):
        e.reset()
        e.run()
        if world > 1:
            dist.all_gather_into_tensor(gathered.reshape(-1), torch.stack(views).reshape(-1))

    nsteps = max(2, args.steps // 2)
    ms_total, _ = timed(ens, step, nsteps, 1, False)
    ms_step = ms_total / nsteps
    st, _ = ens.status()
    frac, _ = ens.fetch_tracking(2500)                       # [members, 11 pools, 12 sources]
    ok = st == 0
    sums = torch.tensor(np.concatenate([frac[ok, 0, :].sum(axis=0), [ok.sum(), (~ok).sum()]]),
                        dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(sums)
    sums = sums.cpu().numpy()
    res = None
    if rank == 0:
        from oracle import port
        # parity of this rank's first member against the oracle (trajectory and tracked maps)
        p = port.default_params(end_year=2500, **{n: X6[a, j] for j, n in enumerate(names6)})
        stt, _, out, ofrac, _ = port.run_member_tracked(ext, 1750, p)
        years = np.arange(1746, 2501, dtype=np.float64)
        mine0 = (gathered[0, :, :, 0] if world > 1 else torch.stack(views)[:, :, 0]).cpu().numpy()
        e_co2 = float(np.max(np.abs(mine0[0] - out[0]) / np.maximum(np.abs(out[0]), 1.0)))
        e_tas = float(np.max(np.abs(mine0[1] - out[1]) / np.maximum(np.abs(out[1]), 0.01)))
        e_map = float(np.abs(frac[0] - ofrac[2500 - 1746]).max())
        res = {"members": total, "members_per_gpu": b - a, "years": 755,
               "value": total * 755 / (ms_step * 1e-3), "unit": UNIT, "ms_per_step": ms_step,
               "failed_members": int(sums[-1]),
               "engine_device_memory_gb": engine_bytes / 1e9,
               "gathered_bytes_per_gpu": int(world * 2 * ny * stride * 8),
               "exchange": "none (1 GPU)" if world == 1 else "one NCCL all-gather of CO2 + Tgav "
                           "(755 years, every member) after the run",
               "tracking_summary": {"pool": "atmos_co2", "year": 2500,
                                    "sources": ["atmos_co2", "earth_c", "veg_c", "detritus_c", "soil_c",
                                                "permafrost_c", "thawedp_c", "HL", "LL", "intermediate",
                                                "deep", "untracked"],
                                    "ensemble_mean_fraction": (sums[:12] / max(sums[12], 1.0)).tolist()},
               "parity_spot": {"member": int(a), "CO2_concentration": e_co2, "global_tas": e_tas,
                               "tracked_fractions_abs": e_map,
                               "ok": bool(e_co2 <= 1e-10 and e_tas <= 1e-10 and e_map <= 1e-12)},
               "note": "BASELINE.json configs[4]: Monte-Carlo over six parameters, SSP5-8.5 "
                       "1745->2500 (series held at their 2300 values), carbon tracking from 1750 "
                       "(11 pools x 12 sources per member), contiguous member shards"}
    if dist:
        dist.barrier()
    ens.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members", type=int, default=65536, help="members per GPU")
    ap.add_argument("--small-members", type=int, default=1024,
                    help="also time BASELINE.json configs[1] (0 = skip)")
    ap.add_argument("--multi-scenario-members", type=int, default=65536,
                    help="also time an ensemble over all 8 SSP scenarios interleaved (0 = skip)")
    ap.add_argument("--tracked-members", type=int, default=65536,
                    help="also time a carbon-tracking ensemble to 2500 (0 = skip)")
    ap.add_argument("--biome-members", type=int, default=65536,
                    help="also time a three-biome ensemble (0 = skip)")
    ap.add_argument("--e2e-segments", type=int, default=4)
    ap.add_argument("--exchange", default="auto", choices=["auto", "push", "ipc", "nccl"],
                    help="N > 1: how the trajectories are exchanged.  push (auto): ONE launch per "
                         "rank, every finished 16-year slab copied into the peers' gather blocks "
                         "over NVLink while the kernel runs (hx_run_exchange); ipc: round 1's "
                         "pulls of 4 run segments; nccl: one all-gather after the run")
    ap.add_argument("--config4-members", type=int, default=-1,
                    help="BASELINE.json configs[3] leg: total members over all ranks, 8 SSPs "
                         "interleaved (default 262 144 at 8 GPUs, else skipped; 0 = skip)")
    ap.add_argument("--config5-members", type=int, default=-1,
                    help="BASELINE.json configs[4] leg: total members, SSP5-8.5 to 2500 with carbon "
                         "tracking (default 1 048 576 at 8 GPUs, else skipped; 0 = skip)")
    ap.add_argument("--exchange-segments", type=int, default=4)
    ap.add_argument("--gather-segments", type=int, default=1,
                    help="N > 1: run segments per step, each followed by its share of the "
                         "all-gather (measured at N = 2 and 8: no gain over one gather at the "
                         "end, NCCL's CTAs do not fit next to the persistent run kernel)")
    ap.add_argument("--ref-members-per-core", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cold-newton", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import hector_b200 as hb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    from hector_b200.sharding import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local_rank)   # before any pinned host buffer exists
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    M = args.members
    X_all = lhs(M * world)
    X = np.ascontiguousarray(X_all[rank * M:(rank + 1) * M])
    table = scenario_table()
    # a dedicated (non-default) stream: the engine launches on it and torch's events time it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    years = np.arange(1746, 2301, dtype=np.float64)

    def make_engine(m, x):
        e = hb.Ensemble(m, table, device=local_rank, outputs=["CO2_concentration", "global_tas"],
                        cold_newton=args.cold_newton, stream=stream.cuda_stream)
        for j, nme in enumerate(PARAMS):
            e.setvar(nme, np.ascontiguousarray(x[:, j]))
        e.prepare()
        e.synchronize()
        return e

    ens = make_engine(M, X)
    gather_buf = None
    views = []
    if world > 1:
        for v in ("CO2_concentration", "global_tas"):
            ptr, stride, ny = ens.output_device(v)
            views.append(torch.as_tensor(CudaArrayView(ptr, (ny, stride)), device="cuda"))
        # The job's one exchange: every rank ends up with all members' CO2 / Tgav trajectories --
        # by default a single all-gather per variable after the last year.  --gather-segments k
        # issues it per finished run segment (whole 16-year slabs) instead, asynchronously.
        nseg = max(1, args.gather_segments)
        nslab = (YEARS + 15) // 16
        cuts = sorted({min(YEARS, ((nslab * (k + 1)) // nseg) * 16) for k in range(nseg)} | {YEARS})
        cuts = [c for c in cuts if c > 0]
        seg_rows = list(zip([0] + cuts[:-1], cuts))
        gather_buf = [[torch.empty(world * (b - a) * t.shape[1], dtype=torch.float64, device="cuda")
                       for (a, b) in seg_rows] for t in views]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")  # 256 MB > L2

    # Exchange over peer memory (default at N > 1): every rank opens its peers' output blocks
    # through CUDA IPC and pulls each finished run segment with copy-engine transfers over
    # NVLink while its own kernel computes the next segment -- no collective kernel has to find
    # room next to the persistent run kernel.  Host-side ordering: a gloo barrier per segment.
    peer_all = None
    gloo = None
    exchange = None
    push = None
    want = args.exchange if args.exchange != "auto" else "push"
    if world > 1 and want in ("push", "ipc"):
        from hector_b200.sharding import PeerExchange, PushExchange
        gloo = dist.new_group(backend="gloo")
        ok = 1
        try:
            if want == "push":
                push = PushExchange(ens, gloo)
            else:
                exchange = PeerExchange(ens, E2E_VARS, gloo, segments=args.exchange_segments)
        except Exception as ex:  # e.g. CUDA IPC not permitted in this container
            sys.stderr.write("bench.py: peer-memory exchange unavailable on rank %d (%r); "
                             "using the NCCL all-gather\n" % (rank, ex))
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=gloo)   # all ranks or none
        if int(flag.item()) == 1:
            peer_all = ([push.block[:, k] for k in range(len(E2E_VARS))] if push is not None
                        else [exchange.blocks[v] for v in E2E_VARS])
        else:
            exchange = push = None

    def step_ipc(e):
        e.reset()
        if push is not None:
            push.run()
        else:
            exchange.run()

    def step_device(e, do_gather=True):
        if world > 1 and do_gather and peer_all is not None:
            return step_ipc(e)
        e.reset()
        if world > 1 and do_gather:
            works = []
            for k, (a, b) in enumerate(seg_rows):
                e.run(1745 + b)
                for t, g in zip(views, gather_buf):
                    works.append(dist.all_gather_into_tensor(g[k], t[a:b].reshape(-1),
                                                             async_op=True))
            for wk in works:
                wk.wait()
        else:
            e.run()

    def timed(e, fn, steps, warmup, flush_l2):
        for _ in range(warmup):
            fn(e)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        kernel_ms = []
        t_ms = 0.0
        for _ in range(steps):
            if flush_l2:
                flush.fill_(1.0)
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            fn(e)
            ev1.record(stream)
            torch.cuda.synchronize()
            t_ms += ev0.elapsed_time(ev1)
            kernel_ms.append(e.last_run_ms)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), float(np.mean(kernel_ms))

    if peer_all is not None:
        # one-off check of the exchange against NCCL's all-gather of the same blocks
        step_ipc(ens)
        for t, g in zip(views, peer_all):
            ref = torch.empty(g.shape, dtype=g.dtype, device=g.device)
            dist.all_gather_into_tensor(ref.reshape(-1), t.reshape(-1))
            torch.cuda.synchronize()
            if not torch.equal(torch.nan_to_num(ref), torch.nan_to_num(g.contiguous())):
                raise SystemExit("bench.py: peer-memory exchange differs from the NCCL all-gather")
        dist.barrier(group=gloo)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, kernel_ms = timed(ens, step_device, args.steps, args.warmup, M < 16384)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        # the step above runs in segments; the roofline wants the whole run kernel's duration
        _, kernel_ms = timed(ens, lambda e: (e.reset(), e.run()), 2, 1, False)
    ms_per_step = total_ms / args.steps
    value = world * M * YEARS / (ms_per_step * 1e-3)
    cnt = ens.counters()
    st, _ = ens.status()
    failed = int((st != 0).sum())
    spot = None
    fp64_peak = None
    if rank == 0:
        try:
            spot = parity_spot(ens, X)
        except Exception as ex:  # the checker is optional for the measurement
            spot = {"ok": None, "error": repr(ex)}
        import ctypes as C
        tf = C.c_double()
        if hb._capi.lib().hx_measure_fp64_peak(local_rank, C.byref(tf), None) == 0:
            fp64_peak = tf.value

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        pinned_in = [torch.from_numpy(np.ascontiguousarray(X[:, j])).pin_memory() for j in range(4)]
        pinned_out = [torch.empty((YEARS, M), dtype=torch.float64).pin_memory() for _ in range(2)]
        e2e_vars = ["CO2_concentration", "global_tas"]
        e2e_outs = [p.numpy() for p in pinned_out]

        def step_e2e(e):
            for j, nme in enumerate(PARAMS):
                e.setvar(nme, pinned_in[j].numpy())      # host -> device
            e.reset()                                    # set-up + spin-up (parameters changed)
            # run + device -> host of every year of both variables; a slab's rows are copied
            # while the kernel computes the later slabs (hx_run_stream)
            e.run_stream(e2e_vars, outs=e2e_outs, segments=args.e2e_segments)

        e2e_ms, _ = timed(ens, step_e2e, max(2, args.steps // 2), 1, False)
        e2e_ms /= max(2, args.steps // 2)
        e2e = {"value": world * M * YEARS / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(4 * M * 8), "d2h_bytes_per_step": int(2 * M * YEARS * 8),
               "ms_per_step": e2e_ms,
               "delivers": "every rank's own members to that rank's pinned host buffers (no cross-"
                           "rank exchange on this path: the host side of a sharded job reads its "
                           "shard)" if world > 1 else "all members to the caller's host buffers",
               "includes": "4 parameter vectors H2D, set-up + spin-up, run, CO2+Tgav x 555 yr D2H "
                           "(hx_run_stream: one launch of the persistent run kernel, every "
                           "16-year slab's rows copied out as the kernel reports it complete)"}

    # ---- BASELINE.json configs[1]: the 1 024-member ensemble on one GPU ----
    small = None
    if args.small_members and rank == 0 and world == 1:
        ms_ = args.small_members
        es = make_engine(ms_, lhs(ms_))
        sm_ms, sm_k = timed(es, lambda e: (e.reset(), e.run()), args.steps, 3, True)
        small = {"members": ms_, "value": ms_ * YEARS / (sm_ms / args.steps * 1e-3), "unit": UNIT,
                 "ms_per_step": sm_ms / args.steps, "kernel_ms": sm_k,
                 "note": "BASELINE.json configs[1]; fills "
                 "%d of 148 SMs' worth of CTAs" % ((ms_ + 127) // 128)}
        es.close()

    # ---- BASELINE.json configs[3] flavour: all 8 SSP scenarios interleaved in one engine ----
    multi = None
    if args.multi_scenario_members and rank == 0 and world == 1:
        names = ["ssp119", "ssp126", "ssp245", "ssp370", "ssp434", "ssp460", "ssp534-over",
                 "ssp585"]
        mm = args.multi_scenario_members
        em = hb.Ensemble(mm, [scenario_table(n) for n in names],
                         member_scenario=np.arange(mm) % 8, device=local_rank,
                         outputs=["CO2_concentration", "global_tas"], stream=stream.cuda_stream)
        Xm = lhs(mm)
        for j, nme in enumerate(PARAMS):
            em.setvar(nme, np.ascontiguousarray(Xm[:, j]))
        em.prepare()
        em.synchronize()
        mm_ms, _ = timed(em, lambda e: (e.reset(), e.run()), max(2, args.steps // 2), 3, False)
        mm_ms /= max(2, args.steps // 2)
        stm, _ = em.status()
        multi = {"members": mm, "scenarios": 8, "value": mm * YEARS / (mm_ms * 1e-3), "unit": UNIT,
                 "ms_per_step": mm_ms, "failed_members": int((stm != 0).sum()),
                 "note": "BASELINE.json configs[3] flavour on one GPU: member i runs scenario "
                         "i mod 8 (the engine groups members by scenario internally and "
                         "un-permutes on fetch)"}
        em.close()

    # ---- BASELINE.json configs[4] flavour: SSP5-8.5 to 2500 with carbon tracking on ----
    tracked = None
    if args.tracked_members and rank == 0 and world == 1:
        raw = scenario_table("ssp585")
        ext = np.vstack([raw, np.repeat(raw[-1:], 200, axis=0)])  # series held at 2300 values
        mt = args.tracked_members
        rng = np.random.Generator(np.random.PCG64(20241018))
        lo6 = np.array([2.0, 1.0, 0.2, 0.5, 0.5, 0.8]); hi6 = np.array([5.0, 2.6, 0.9, 2.5, 1.5, 1.2])
        X6 = lo6 + rng.random((mt, 6)) * (hi6 - lo6)
        et = hb.Ensemble(mt, ext, end_year=2500, device=local_rank,
                         outputs=["CO2_concentration", "global_tas"], tracking_date=1750,
                         track_every=0, stream=stream.cuda_stream)
        for j, nme in enumerate(PARAMS + ["aero_scalar", "vol_scalar"]):
            et.setvar(nme, np.ascontiguousarray(X6[:, j]))
        et.prepare()
        et.synchronize()
        tr_ms, _ = timed(et, lambda e: (e.reset(), e.run()), max(2, args.steps // 2), 3, False)
        tr_ms /= max(2, args.steps // 2)
        stt, _ = et.status()
        tracked = {"members": mt, "years": 755, "value": mt * 755 / (tr_ms * 1e-3), "unit": UNIT,
                   "ms_per_step": tr_ms, "failed_members": int((stt != 0).sum()),
                   "note": "BASELINE.json configs[4] flavour on one GPU: Monte-Carlo over (S, "
                           "q10_rh, beta, diff, aero_scalar, vol_scalar), SSP5-8.5 1745->2500, "
                           "carbon tracking from 1750 (11 pools x 12 sources per member)"}
        et.close()

    # ---- biome-split land pools: three biomes, per-member biome parameters ----
    biome = None
    if args.biome_members and rank == 0 and world == 1:
        mb_ = args.biome_members
        glob = dict(npp_flux0=56.2, veg_c=550.0, detritus_c=55.0, soil_c=917.0, permafrost_c=865.0)
        fracs = {"tundra": 0.2, "amazon": 0.45, "midlat": 0.35}
        eb = hb.Ensemble(mb_, scenario_table(), device=local_rank, biomes=list(fracs),
                         outputs=["CO2_concentration", "global_tas"], stream=stream.cuda_stream)
        Xb = lhs(mb_)
        for b, fr in fracs.items():
            eb.set_biome(b, f_nppv=0.35, f_nppd=0.60, f_litterd=0.98,
                         **{k: v * fr for k, v in glob.items()})
            eb.setvar(b + ".q10_rh", np.ascontiguousarray(Xb[:, 1]))
            eb.setvar(b + ".beta", np.ascontiguousarray(Xb[:, 2]))
        eb.setvar("tundra.warmingfactor", 2.0)
        eb.setvar("S", np.ascontiguousarray(Xb[:, 0]))
        eb.setvar("diff", np.ascontiguousarray(Xb[:, 3]))
        eb.prepare()
        eb.synchronize()
        bi_ms, _ = timed(eb, lambda e: (e.reset(), e.run()), max(2, args.steps // 2), 3, False)
        bi_ms /= max(2, args.steps // 2)
        stb, _ = eb.status()
        biome = {"members": mb_, "biomes": 3, "value": mb_ * YEARS / (bi_ms * 1e-3), "unit": UNIT,
                 "ms_per_step": bi_ms, "failed_members": int((stb != 0).sum()),
                 "note": "the headline ensemble with the land split into three biomes (pools and "
                         "NPP 0.2 / 0.45 / 0.35, per-member beta and q10_rh in every biome, "
                         "tundra warming factor 2): the BIOMES instantiation of the run kernel"}
        eb.close()

    # ---- BASELINE.json configs[3] and configs[4] AT SIZE, sharded over all ranks ----
    config4 = config5 = None
    n4 = args.config4_members if args.config4_members >= 0 else (262144 if world == 8 else 0)
    n5 = args.config5_members if args.config5_members >= 0 else (1048576 if world == 8 else 0)
    if n4:
        config4 = run_config4(n4, rank, world, local_rank, stream, dist, gloo, timed, args)
    if n5:
        config5 = run_config5(n5, rank, world, local_rank, stream, dist, timed, args)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = roofline_of("hx_run_kernel", M, YEARS, kernel_ms, peak, bool(peaks), fp64_peak,
                           "latest_traffic.json", "latest_fp64_counts.json")
    if small is not None:
        small["roofline"] = roofline_of("hx_run_kernel (1 024 members)", small["members"], YEARS,
                                        small["kernel_ms"], peak, bool(peaks), fp64_peak,
                                        "latest_traffic_small.json", "latest_fp64_counts.json")

    cpu = None
    if not args.no_cpu_baseline:
        try:
            v, kind, cores, what = cpu_reference_pass(args.ref_members_per_core)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": what}
        except Exception as ex:  # the checker is optional for the measurement
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable",
                   "sample": repr(ex)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args), "clocks": clocks, "e2e": e2e,
        # per step: one run kernel + one NaN-fill kernel per run segment (1 segment at N = 1,
        # --exchange-segments / --gather-segments of them at N > 1); restores and exchanges
        # are copies, not kernels
        "gpu_launches": 2 * args.steps * (1 if (world == 1 or push is not None) else
                                          (len(exchange.segments) if exchange is not None
                                           else len(seg_rows))),
        "roofline": roofline, "cpu_baseline": cpu,
        "work_per_member_year": {k: cnt[k] / max(1, cnt["member_years"]) for k in
                                 ("rhs_evals", "rk_steps", "stashes", "newton_iterations",
                                  "newton_calls")},
        "failed_members": failed, "parity_spot": spot, "small_ensemble": small, "multi_scenario_ensemble": multi,
        "tracked_ensemble": tracked,
        "biome_ensemble": biome,
        "exchange": None if world == 1 else
        ("peer memory, pushed: one launch per rank, each finished 16-year slab copied into every "
         "peer's gather block over NVLink while the kernel computes (hx_run_exchange), one host "
         "barrier" if push is not None else
         "peer memory (CUDA IPC pulls, %d run segments)" % len(exchange.segments)
         if exchange is not None else "NCCL all-gather (%d run segments)" % len(seg_rows)),
        "numa": numa, "config4": config4, "config5": config5,
    }
    print(json.dumps(line))
    ens.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
