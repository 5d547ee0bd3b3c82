# replay variants: tracking tests, then same-call A/B, then executed-instruction counts of the replay
python -m pytest tests/test_gpu_tracking.py tests/test_gpu_parity_at_size.py -m gpu -q -x -k "track" 2>&1 | tail -3
for rep in 1 2; do
for f in hector_b200/libhector_b200.so hector_b200/ab_*.so; do
  echo "== $f"; HECTOR_B200_LIB=$PWD/$f python tools/profile_tracked.py 65536 | tr '\n' ' '; echo
done; done 2>&1 | tee gpurun_out/r02_ab_trkmix.log
ncu --set full --clock-control none --import-source on -k regex:hx_track_kernel -s 20 -c 1 -o gpurun_out/r02_trk_v8 python tools/profile_tracked.py 65536 > gpurun_out/ncu_trk_v8.log 2>&1; tail -2 gpurun_out/ncu_trk_v8.log
