set -x
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -c 1 -o gpurun_out/r02_biomes_v3 python tools/profile_biomes_one.py > gpurun_out/ncu_biomes_v3.log 2>&1; tail -2 gpurun_out/ncu_biomes_v3.log
