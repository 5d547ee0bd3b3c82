set -x
python -m pytest tests/test_compat.py tests/test_core_facade.py -q -m gpu > gpurun_out/r02_compat.log 2>&1; tail -25 gpurun_out/r02_compat.log
# FMA discriminator: the same all-parameter probe with a build that rounds like x86-64 (no contraction)
HECTOR_B200_LIB=$PWD/hector_b200/libhector_b200_nofma.so python tools/gpu_all_params_vs_oracle.py 64 5 > gpurun_out/r02_allparams_nofma.log 2>&1; cat gpurun_out/r02_allparams_nofma.log
python tools/gpu_all_params_vs_oracle.py 64 5 > gpurun_out/r02_allparams_fma.log 2>&1; cat gpurun_out/r02_allparams_fma.log
