"""Raw pinned-copy microbenchmark: every rank (one per GPU, torchrun) copies a pinned host buffer to its GPU and back,
all ranks at once, and rank 0 prints the per-rank and aggregate GB/s for H2D alone, D2H alone and both directions
together.  This is the ceiling of the host path of bench.py (e2e): 65.6 MB per step on rank 0, 32 MB on the others.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_bench.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
MB = 64
n = MB * 1024 * 1024 // 8
h_in = torch.empty(n, dtype=torch.complex64).pin_memory()
h_out = torch.empty(n, dtype=torch.complex64).pin_memory()
d_a = torch.empty(n, dtype=torch.complex64, device=dev)
d_b = torch.empty(n, dtype=torch.complex64, device=dev)
s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(mode, iters=40):
    barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        if mode in ('h2d', 'both'):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if mode in ('d2h', 'both'):
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    barrier()
    dt = time.perf_counter() - t0
    gbs = iters * MB * 1.048576e-3 / dt * (2 if mode == 'both' else 1)
    t = torch.tensor([gbs], device=dev, dtype=torch.float64)
    mn = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return float(t.item()), float(mn.item())


out = {'ranks': world, 'buffer_MB': MB}
for mode in ('h2d', 'd2h', 'both'):
    run(mode, 5)
    agg, mn = run(mode)
    out[mode] = {'aggregate_GBps': round(agg, 1), 'slowest_rank_GBps': round(mn, 1)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
