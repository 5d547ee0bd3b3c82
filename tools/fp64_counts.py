"""Turn an `ncu --metrics ...op_{dfma,dmul,dadd,fp64}..., dram__bytes... --csv` log of ONE run-kernel
launch (tools/gpu_call_r02_r.sh) into the small json files bench.py's roofline reads:
profiles/latest_fp64_counts.json, profiles/latest_traffic.json (65 536 members) or
profiles/latest_traffic_small.json (1 024 members).
usage: python tools/fp64_counts.py <log.csv> <members> [small]"""
import csv, json, os, sys
log, members = sys.argv[1], int(sys.argv[2])
small = len(sys.argv) > 3
rows = [r for r in csv.reader(open(log)) if len(r) > 14 and r[0].isdigit()]
v = {r[12]: float(r[14].replace(",", "")) for r in rows}
kern = rows[0][4]
my = members * 555
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
src = "profiles/" + os.path.basename(log)
traffic = {"kernel": "hx_run_kernel" + (" (1 024 members)" if small else ""), "member_years": my,
           "dram_bytes_per_launch": v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"],
           "source": src + " (ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch, %d members x 555 years)" % members}
json.dump(traffic, open(os.path.join(root, "latest_traffic_small.json" if small else "latest_traffic.json"), "w"), indent=1)
if not small:
    d = {"kernel": kern, "member_years": my,
         "dfma_per_member_year": v["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] / my,
         "dmul_per_member_year": v["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] / my,
         "dadd_per_member_year": v["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] / my,
         "fp64_inst_per_member_year": v["smsp__sass_thread_inst_executed_op_fp64_pred_on.sum"] / my,
         "warp_inst_per_launch": v["smsp__inst_executed.sum"],
         "threads_per_warp_instruction": v["smsp__thread_inst_executed_per_inst_executed.ratio"],
         "source": src + " (ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd,fp64}_pred_on.sum, one launch, "
                         "65536 members x 555 years, SSP2-4.5 LHS, k-d member order)"}
    json.dump(d, open(os.path.join(root, "latest_fp64_counts.json"), "w"), indent=1)
    print(json.dumps(d, indent=1))
print(json.dumps(traffic, indent=1))
