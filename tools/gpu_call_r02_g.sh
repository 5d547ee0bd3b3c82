set -x
python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_parity.py tests/test_gpu_constraints.py tests/test_gpu_biomes.py -m gpu -q -x > gpurun_out/r02_gputests_g.log 2>&1; tail -8 gpurun_out/r02_gputests_g.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --small-members 0 --multi-scenario-members 65536 --tracked-members 0 --biome-members 0 > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; tail -3 gpurun_out/r02_bench_g.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_g.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['multi_scenario_ensemble']['ms_per_step'], d['parity_spot']['ok'])"
