set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_j.log 2>&1; tail -6 gpurun_out/r02_gputests_j.log
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -c 1 -o gpurun_out/r02_v13 python tools/profile_run.py 65536 1 > gpurun_out/ncu_v13.log 2>&1; tail -3 gpurun_out/ncu_v13.log
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:hx_run_kernel -c 1 --csv --log-file gpurun_out/r02_fp64_counts_v13.csv python tools/profile_run.py 65536 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_v13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --small-members 0 --multi-scenario-members 0 --tracked-members 0 --biome-members 0 > gpurun_out/b_ncu_v13.log 2>&1
tail -3 gpurun_out/b_ncu_v13.log | cut -c1-300
