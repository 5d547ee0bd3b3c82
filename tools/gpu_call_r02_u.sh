# replay variants: tracking tests, then same-call A/B
python -m pytest tests/test_gpu_tracking.py tests/test_gpu_parity_at_size.py -m gpu -q -x -k "track" 2>&1 | tail -3
for rep in 1 2; do
for f in hector_b200/libhector_b200.so hector_b200/ab_*.so; do
  echo "== $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_tracked.py 65536 | tr '\n' ' '; echo
done; done 2>&1 | tee gpurun_out/r02_ab_trkmix.log
