set -x
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -s 5 -c 1 -o gpurun_out/r02_trkrun python tools/profile_tracked.py 65536 > gpurun_out/ncu_trkrun.log 2>&1; tail -2 gpurun_out/ncu_trkrun.log
