set -x
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -c 1 -o gpurun_out/r02_v16 python tools/profile_run.py 65536 1 > gpurun_out/ncu_v16.log 2>&1; tail -2 gpurun_out/ncu_v16.log
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -c 1 -o gpurun_out/r02_small_v16 python tools/profile_run.py 1024 1 > gpurun_out/ncu_small_v16.log 2>&1; tail -2 gpurun_out/ncu_small_v16.log
