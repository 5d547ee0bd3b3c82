set -x
ncu --set full --clock-control none --import-source on -k regex:hx_track_kernel -s 20 -c 1 -o gpurun_out/r02_trk_v5 python tools/profile_tracked.py 65536 > gpurun_out/ncu_trk_v5.log 2>&1; tail -2 gpurun_out/ncu_trk_v5.log
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -s 5 -c 1 -o gpurun_out/r02_trkrun_v2 python tools/profile_tracked.py 65536 > gpurun_out/ncu_trkrun_v2.log 2>&1; tail -2 gpurun_out/ncu_trkrun_v2.log
