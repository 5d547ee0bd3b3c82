set -x
python -m pytest tests/test_gpu_tracking.py tests/test_gpu_parity_at_size.py -q -x 2>&1 | tail -4
for g in 1 2 4 8; do
  echo "== HX_TRK_GROUP=$g"; HX_TRK_GROUP=$g python tools/profile_tracked.py 65536 | tr '\n' ' '; echo
done 2>&1 | tee gpurun_out/r02_ab_trk_group.log
