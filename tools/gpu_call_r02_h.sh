set -x
python -m pytest tests -m gpu -q -x -s > gpurun_out/r02_gputests_h.log 2>&1; tail -12 gpurun_out/r02_gputests_h.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tracked-members 0 --biome-members 0 > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; tail -3 gpurun_out/r02_bench_h.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_h.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['multi_scenario_ensemble']['ms_per_step'], d['small_ensemble']['ms_per_step'], d['parity_spot']['ok'])"
