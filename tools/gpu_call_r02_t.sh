# NVTX build: quick GPU tests; no-year-barrier A/B at 1 024 and 65 536 members; all-parameter probe, more seeds
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do
for f in hector_b200/libhector_b200.so hector_b200/ab_nosync.so; do
  echo "== small $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_run.py 1024 4 | grep "run ms" | tail -3 | tr '\n' ' '; echo
  echo "== big $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_run.py 65536 3 | grep "run ms" | tail -2 | tr '\n' ' '; echo
done; done 2>&1 | tee gpurun_out/r02_ab_nosync.log
for seed in 11 12 13; do python tools/gpu_all_params_vs_oracle.py 64 $seed 2>&1 | tail -6; done | tee gpurun_out/r02_allparams_seeds.log
