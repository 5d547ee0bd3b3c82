set -x
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02_ab_lds128.log
for rep in 1 2; do
for f in hector_b200/libhector_b200.so hector_b200/ab_head.so; do
  echo "== small $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_run.py 1024 4 | grep "run ms" | tail -3 | tr '\n' ' '; echo
done; done 2>&1 | tee -a gpurun_out/r02_ab_lds128.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests_p.log 2>&1; tail -4 gpurun_out/r02_gputests_p.log
