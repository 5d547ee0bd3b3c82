set -x
python -m pytest tests/test_gpu_parity.py -q -k "vector_transcendentals or early_test_division" 2>&1 | tail -5
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02_ab_div.log
for f in hector_b200/libhector_b200.so hector_b200/ab_v15.so; do
  echo "== small $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_run.py 1024 4 | grep "run ms" | tail -3 | tr '\n' ' '; echo
done 2>&1 | tee -a gpurun_out/r02_ab_div.log
python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_p.log 2>&1; tail -6 gpurun_out/r02_gputests_p.log
