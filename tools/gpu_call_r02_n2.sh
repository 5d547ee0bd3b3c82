set -x
python -m pytest tests/test_gpu_exchange.py -q 2>&1 | tail -3
time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 --config4-members 32768 --config5-members 16384 > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err
tail -5 gpurun_out/r02_bench_n2b.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n2b.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print({k: d[k] for k in ("value", "ms_per_step", "exchange", "numa", "failed_members")})
        print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"]); print("kernel_ms", d["roofline"]["kernel_ms"])
        for k in ("config4", "config5"):
            c = d.get(k)
            if c: print(k, {x: c[x] for x in c if x not in ("note", "tracking_summary")})
PY
