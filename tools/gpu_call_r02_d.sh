set -x
python -m pytest tests/test_compat.py tests/test_core_facade.py -q -m gpu > gpurun_out/r02_compat.log 2>&1; tail -5 gpurun_out/r02_compat.log
HECTOR_B200_LIB=$PWD/hector_b200/libhector_b200_nofma.so python tools/gpu_all_params_vs_oracle.py 64 5 gpurun_out/allparams_nofma.npz > gpurun_out/r02_allparams_nofma.log 2>&1; head -8 gpurun_out/r02_allparams_nofma.log
python tools/gpu_all_params_vs_oracle.py 64 5 gpurun_out/allparams_fma.npz > gpurun_out/r02_allparams_fma.log 2>&1; head -8 gpurun_out/r02_allparams_fma.log
