set -x
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -c 1 -o gpurun_out/r02_v14 python tools/profile_run.py 65536 1 > gpurun_out/ncu_v14.log 2>&1; tail -2 gpurun_out/ncu_v14.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_v14.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --small-members 0 --multi-scenario-members 0 --tracked-members 0 --biome-members 0 > gpurun_out/b_ncu_v14.log 2>&1
