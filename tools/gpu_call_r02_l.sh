set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_l.log 2>&1; tail -4 gpurun_out/r02_gputests_l.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err; tail -3 gpurun_out/r02_bench_l.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_l.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['multi_scenario_ensemble']['ms_per_step'], d['small_ensemble']['ms_per_step'], d['tracked_ensemble']['ms_per_step'], d['biome_ensemble']['ms_per_step'], d['parity_spot']['ok'], d['cpu_baseline'])
print(d['roofline'])"
