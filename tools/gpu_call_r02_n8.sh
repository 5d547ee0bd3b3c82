set -x
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1; nproc >> gpurun_out/topo8.txt
for i in 0 1 2 3 4 5 6 7; do b=$(nvidia-smi -i $i --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | cut -c5-); echo "$i $b numa=$(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) cpus=$(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)" >> gpurun_out/topo8.txt; done
time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -5 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n8.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print({k: d[k] for k in ("value", "ms_per_step", "exchange", "numa", "failed_members")})
        print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"]); print("kernel_ms", d["roofline"]["kernel_ms"])
        for k in ("config4", "config5"):
            c = d[k]; print(k, {x: c[x] for x in c if x not in ("note", "tracking_summary")})
PY
