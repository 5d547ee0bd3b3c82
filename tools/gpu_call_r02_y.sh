set -x
# the all-output builds with the per-stash outputs (scratch rows X): latency and general build
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02c_memcheck_stashout.log python tools/sanitize_driver.py stashout > gpurun_out/san_mem_stashout.out 2>&1
HX_NO_LAT=1 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02c_memcheck_stashout_general.log python tools/sanitize_driver.py stashout allout > gpurun_out/san_mem_stashout_general.out 2>&1
HX_NO_LAT=1 HX_SAN_TO=1765 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02c_racecheck_stashout_general.log python tools/sanitize_driver.py stashout > gpurun_out/san_race_stashout_general.out 2>&1
HX_SAN_TO=1765 timeout 600 compute-sanitizer --tool initcheck --log-file gpurun_out/r02c_initcheck_stashout.log python tools/sanitize_driver.py stashout > gpurun_out/san_init_stashout.out 2>&1
for f in gpurun_out/r02c_*check_*.log; do echo $f; tail -n 1 $f; done
cat gpurun_out/san_*_stashout*.out | sort | uniq -c | head
