set -x
python tools/fp64_peak.py gpurun_out/fp64_peak.json
python tools/gpu_all_params_vs_oracle.py 64 5 gpurun_out/allparams_64_5.npz > gpurun_out/allparams_64_5.log 2>&1
# FP64 instruction counts + dram bytes of the headline launch (metrics pass, not a timing)
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum --clock-control none -k regex:hx_run_kernel -c 1 --csv --log-file gpurun_out/r02_fp64_counts_65536.csv python tools/profile_run.py 65536 1 > gpurun_out/ncu_counts.log 2>&1
# the 1 024-member launch: where does one warp's latency go
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -c 1 -o gpurun_out/r02_small1024 python tools/profile_run.py 1024 1 > gpurun_out/ncu_small.log 2>&1
# sanitizers
for fl in plain constrained tracked biomes stream spinup; do
  timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_$fl.log python tools/sanitize_driver.py $fl > gpurun_out/san_mem_$fl.out 2>&1
done
for fl in plain tracked stream; do
  HX_SAN_TO=1765 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck_$fl.log python tools/sanitize_driver.py $fl > gpurun_out/san_race_$fl.out 2>&1
done
HX_SAN_TO=1765 timeout 600 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_synccheck_plain.log python tools/sanitize_driver.py plain > gpurun_out/san_sync_plain.out 2>&1
tail -3 gpurun_out/r02_memcheck_*.log gpurun_out/r02_racecheck_*.log gpurun_out/r02_synccheck_*.log
cat gpurun_out/fp64_peak.json
