"""Bare host<->device copy rate of the box under N concurrent ranks (no kernels): each rank copies a 1 GiB pinned
buffer H2D and another D2H at the same time, repeatedly, for ~2 s.  Launch with torchrun; rank 0 prints one line.
This is the ceiling of bench.py's `e2e` figures: they move 8.6 GB each way per rank per step."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
both(); torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter(); reps = 0
while time.perf_counter() - t0 < 2.0:
    both(); torch.cuda.synchronize(); reps += 1
dt = time.perf_counter() - t0
rate = torch.tensor([reps * n / dt / 1e9], device=dev, dtype=torch.float64)
lo = rate.clone(); tot = rate.clone()
if world > 1:
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
if rank == 0:
    print(f"pcie_sweep ranks={world}: per-rank {float(rate):.1f} GB/s each way (min over ranks {float(lo):.1f}), aggregate {float(tot):.1f} GB/s each way; "
          f"cpus visible {len(os.sched_getaffinity(0))}, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')}")
if world > 1: dist.destroy_process_group()
