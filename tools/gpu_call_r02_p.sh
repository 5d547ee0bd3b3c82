set -x
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02_ab_epsrel.log
ncu --set full --clock-control none --import-source on -k regex:hx_run_kernel -c 1 -o gpurun_out/r02_v18 python tools/profile_run.py 65536 1 > gpurun_out/ncu_v18.log 2>&1; tail -2 gpurun_out/ncu_v18.log
python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
