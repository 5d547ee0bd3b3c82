set -x
for n in 1024 4096 16384 65536; do echo "== members $n"; python tools/profile_run.py $n 4 | grep "run ms" | tail -2 | tr '\n' ' '; echo; done 2>&1 | tee gpurun_out/r02_small_sizes2.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests_p.log 2>&1; tail -4 gpurun_out/r02_gputests_p.log
