"""Configuration 5: 3D 128^3, 32 coils, CG sharded by coil across the GPUs of one node.
Run with torchrun (or plain python for 1 GPU).  Prints per-iteration time and, with --check, compares the
sharded result with an unsharded run on rank 0 (small geometry)."""
import argparse, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import torch.distributed as dist
import pynufft_b200
from pynufft_b200.dist import CoilShardedNUFFT, shard_coils

ap = argparse.ArgumentParser()
ap.add_argument('--coils', type=int, default=32)
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--small', action='store_true')
ap.add_argument('--check', action='store_true')
a = ap.parse_args()
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev) if world > 1 else dist.init_process_group('gloo', init_method='tcp://127.0.0.1:29555', rank=0, world_size=1)
if a.small:
    Nd, Kd, Jd, M = (32,) * 3, (64,) * 3, (6,) * 3, 40000
else:
    Nd, Kd, Jd, M = (128,) * 3, (256,) * 3, (6,) * 3, 2_000_000
rng = numpy.random.default_rng(0)
om = rng.uniform(-numpy.pi, numpy.pi, (M, 3))
sl = shard_coils(a.coils, world, rank)
nloc = sl.stop - sl.start
A = pynufft_b200.NUFFT(dev)
A.plan(om, Nd, Kd, Jd, batch=nloc)
op = CoilShardedNUFFT(A, a.coils)
yall = None
g = torch.Generator(device='cpu').manual_seed(1)
y_all = torch.view_as_complex(torch.randn((M, a.coils, 2), generator=g))
y_loc = y_all[:, sl].contiguous().to(dev)
torch.cuda.synchronize(); dist.barrier()
x = op.solve_cg(y_loc, maxiter=1)      # warm-up
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
x = op.solve_cg(y_loc, maxiter=a.iters)
torch.cuda.synchronize(); dist.barrier()
t1 = time.perf_counter()
res = {'world': world, 'coils': a.coils, 'coils_local': nloc, 'iters': a.iters, 'seconds': t1 - t0,
       'ms_per_iter': (t1 - t0) / (a.iters + 1.5) * 1e3, 'finite': bool(torch.isfinite(torch.view_as_real(x)).all())}
if a.check:
    xs = [torch.empty_like(x) for _ in range(world)] if world > 1 else [x]
    if world > 1:
        dist.all_gather(xs, x)
    if rank == 0:
        A2 = pynufft_b200.NUFFT(dev)
        A2.plan(om, Nd, Kd, Jd, batch=a.coils)
        ref = A2._solve_device(y_all.to(dev), 'cg', maxiter=a.iters)
        got = torch.cat(xs, dim=-1)
        res['rel_err_vs_unsharded'] = float(torch.linalg.norm(got - ref) / torch.linalg.norm(ref))
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
