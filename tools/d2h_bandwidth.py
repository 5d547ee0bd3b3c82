"""Aggregate device-to-host bandwidth of N ranks copying at the same time (the size of one
end-to-end step's output: 582 MB per rank into pinned memory), to tell host ingest limits from
engine effects.   torchrun --nproc-per-node N tools/d2h_bandwidth.py"""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("gloo")
n = 2 * 555 * 65536
src = torch.zeros(n, dtype=torch.float64, device="cuda")
dst = torch.empty(n, dtype=torch.float64).pin_memory()
for _ in range(2):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dst.copy_(src, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
t = torch.tensor([ms])
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("ranks %d: %.1f MB per rank in %.2f ms (slowest rank) = %.1f GB/s per rank, %.1f GB/s aggregate"
          % (world, n * 8 / 1e6, t.item(), n * 8 / t.item() / 1e6, world * n * 8 / t.item() / 1e6))
