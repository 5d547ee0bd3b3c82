"""Summarise an .ncu-rep (one kernel launch) into a small text + json that can be committed
under profiles/.   usage: python tools/ncu_summary.py <rep> <out-prefix> [members]"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
members = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max"]
def gb(k):
    v, u = d[k]
    v = float(v)
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[u]
lines = ["ncu --set full --clock-control none, one launch of hx_run_kernel (%d members x 555 years)" % members, ""]
for k in keys:
    if k in d:
        lines.append("%-80s %s %s" % (k, d[k][0], d[k][1]))
traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
alg = members * 555 * 5048.0
ms = float(d["gpu__time_duration.sum"][0]) * {"ms": 1, "us": 1e-3, "s": 1e3, "msecond": 1, "usecond": 1e-3, "second": 1e3}.get(d["gpu__time_duration.sum"][1], 1)
lines += ["", "derived:",
          "  dram traffic per launch      %.1f GB  (read + write)" % (traffic / 1e9),
          "  algorithmic bytes per launch %.1f GB  (5048 B x %d member-years)" % (alg / 1e9, members * 555),
          "  traffic / algorithmic        %.2f" % (traffic / alg),
          "  warp instructions per warp-member-year  %.0f" % (float(d["smsp__inst_executed.sum"][0]) / (members * 555 / 32.0)),
          "  duration under ncu (cold caches, serialised)  %.2f ms" % ms]
open(out + ".txt", "w").write("\n".join(lines) + "\n")
json.dump({"dram_bytes_per_launch": traffic, "algorithmic_bytes_per_launch": alg, "members": members,
           "ncu_duration_ms": ms, "source": rep.split("/")[-1]}, open(out + ".json", "w"), indent=1)
print("\n".join(lines))
