set -x
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
python -m pytest tests/test_gpu_exchange.py -q > gpurun_out/r02_exchange_tests.log 2>&1; tail -5 gpurun_out/r02_exchange_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --config4-members 32768 --config5-members 16384 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -5 gpurun_out/r02_bench_n2.err; cat gpurun_out/r02_bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --exchange nccl --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_n2_nccl.json 2> gpurun_out/r02_bench_n2_nccl.err; tail -3 gpurun_out/r02_bench_n2_nccl.err; cat gpurun_out/r02_bench_n2_nccl.json
