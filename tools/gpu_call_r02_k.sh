set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_k.log 2>&1; tail -4 gpurun_out/r02_gputests_k.log
python tools/gpu_order_invariance.py > gpurun_out/r02_order_invariance.log 2>&1; grep -v "^   " gpurun_out/r02_order_invariance.log
for fl in plain constrained nbp tracked biomes stream spinup; do
  timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_$fl.log python tools/sanitize_driver.py $fl > gpurun_out/san_mem_$fl.out 2>&1
done
for fl in plain tracked stream biomes; do
  HX_SAN_TO=1765 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck_$fl.log python tools/sanitize_driver.py $fl > gpurun_out/san_race_$fl.out 2>&1
done
for fl in plain tracked; do
HX_SAN_TO=1765 timeout 600 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_synccheck_$fl.log python tools/sanitize_driver.py $fl > gpurun_out/san_sync_$fl.out 2>&1
done
HX_SAN_TO=1765 timeout 600 compute-sanitizer --tool initcheck --log-file gpurun_out/r02_initcheck_plain.log python tools/sanitize_driver.py plain > gpurun_out/san_init_plain.out 2>&1
for f in gpurun_out/r02_*check_*.log; do echo $f; tail -n 1 $f; done
cat gpurun_out/san_*_*.out | grep -v "^sanitize_driver done" | sort | uniq -c
