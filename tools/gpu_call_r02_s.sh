set -x
for fl in plain tracked biomes stream; do
  timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02b_memcheck_$fl.log python tools/sanitize_driver.py $fl > gpurun_out/san_mem_$fl.out 2>&1
done
HX_NO_LAT=1 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02b_memcheck_plain_general.log python tools/sanitize_driver.py plain allout > gpurun_out/san_mem_plain_general.out 2>&1
for fl in plain tracked biomes; do
  HX_SAN_TO=1765 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02b_racecheck_$fl.log python tools/sanitize_driver.py $fl > gpurun_out/san_race_$fl.out 2>&1
done
HX_NO_LAT=1 HX_SAN_TO=1765 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02b_racecheck_plain_general.log python tools/sanitize_driver.py plain > gpurun_out/san_race_plain_general.out 2>&1
HX_SAN_TO=1765 timeout 600 compute-sanitizer --tool synccheck --log-file gpurun_out/r02b_synccheck_plain.log python tools/sanitize_driver.py plain > gpurun_out/san_sync_plain.out 2>&1
HX_SAN_TO=1765 timeout 600 compute-sanitizer --tool initcheck --log-file gpurun_out/r02b_initcheck_plain.log python tools/sanitize_driver.py plain > gpurun_out/san_init_plain.out 2>&1
for f in gpurun_out/r02b_*check_*.log; do echo $f; tail -n 1 $f; done
cat gpurun_out/san_*_*.out | grep -v "^sanitize_driver done" | sort | uniq -c | head
