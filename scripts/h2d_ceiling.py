#!/usr/bin/env python
"""Host->device copy ceiling of the box: every rank copies a pinned 384 MB block (one bench step's upload: 2 probes x 192 MB)
to its GPU in a loop; prints per-rank and aggregate GB/s.  Run alone and under torchrun --nproc-per-node 8: the e2e arm of
bench.py is bounded by this number, not by the GPUs."""
import os
import time
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 24 * 500 * 2000
host = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
dev = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
for h in host:
    h.zero_()
def step():
    for h, d in zip(host, dev):
        d.copy_(h, non_blocking=True)
for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
K = 20
for _ in range(K):
    step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = K * 2 * n * 8 / dt * 1e-9
t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t)
if rank == 0:
    print("pinned H2D: %d rank(s), rank 0 %.1f GB/s, aggregate %.1f GB/s (%.1f per rank); a bench step uploads 0.384 GB per GPU -> e2e ceiling %.0f evals/s"
          % (world, gbs, float(t.item()), float(t.item()) / world, 2.0 * float(t.item()) / 0.384))
if world > 1:
    dist.destroy_process_group()
