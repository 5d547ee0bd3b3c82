python -m pytest tests/test_gpu_tracking.py tests/test_gpu_parity_at_size.py tests/test_gpu_constraints.py -q -x 2>&1 | tail -3
for rep in 1 2; do
for f in hector_b200/libhector_b200.so hector_b200/ab_membermajor.so; do
  echo "== $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_tracked.py 65536 | tr '\n' ' '; echo
done; done
