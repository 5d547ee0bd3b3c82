python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests_w.log 2>&1; tail -3 gpurun_out/r02_gputests_w.log
python bench.py > gpurun_out/r02_bench_w.json 2> gpurun_out/r02_bench_w.err; tail -c 600 gpurun_out/r02_bench_w.json
