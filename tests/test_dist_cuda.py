"""
Configuration 5 on real CUDA operators (BASELINE configs[4]: 3-D multi-coil CG sharded by coil, all-reduce of the
many2one image and of the CG dot products).  Two processes share cuda:0 and split the coils (world_size 2, gloo on
CUDA tensors -- NCCL refuses two ranks on one GPU; the multi-GPU NCCL path is what bench.py --gpus N exercises).
Everything goes through pynufft_b200.dist.CoilShardedNUFFT -> pynufft_b200.NUFFT -> libb200nufft.so.
Checked against the UNSHARDED CUDA operator (same library, all coils in one plan) and against the oracle.
"""
import os
import socket
import sys

import numpy
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

ND, KD, JD, B, M, ITERS = (24, 20, 16), (48, 40, 32), (6, 6, 6), 5, 9000, 6


def problem():
    rng = numpy.random.default_rng(17)
    om = rng.uniform(-numpy.pi, numpy.pi, (M, 3))
    sens = (rng.standard_normal(ND + (B,)) + 1j * rng.standard_normal(ND + (B,))).astype(numpy.complex64)
    s = (rng.standard_normal(ND) + 1j * rng.standard_normal(ND)).astype(numpy.complex64)
    y = (rng.standard_normal((M, B)) + 1j * rng.standard_normal((M, B))).astype(numpy.complex64)
    return om, sens, s, y


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import pynufft_b200
        from pynufft_b200.dist import CoilShardedNUFFT, shard_coils
        dev = torch.device('cuda', 0)
        om, sens, s, y = problem()
        sl = shard_coils(B, world, rank)
        A = pynufft_b200.NUFFT(dev)
        A.plan(om, ND, KD, JD, batch=sl.stop - sl.start)
        op = CoilShardedNUFFT(A, B)
        op.set_sense(sens)
        gs = A.to_device(s)
        y_fwd = op.forward_one2many(gs)                              # (M, B_local), no collective
        y_loc = torch.from_numpy(numpy.ascontiguousarray(y[:, sl])).to(dev)
        s_adj = op.adjoint_many2one(y_loc)                           # one all-reduce of the Nd image
        s_self = op.selfadjoint_one2many2one(gs)
        x_cg = op.solve_cg(y_loc, maxiter=ITERS)                     # dot products all-reduced, shared alpha / beta
        torch.cuda.synchronize()
        torch.save({'y_fwd': y_fwd.cpu(), 's_adj': s_adj.cpu(), 's_self': s_self.cpu(), 'x_cg': x_cg.cpu(),
                    'coils': (sl.start, sl.stop), 'launches': int(A._lib.b200nufft_launch_count())},
                   out + '.rank%d' % rank)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        return sk.getsockname()[1]


def rel(a, b):
    a, b = numpy.asarray(a).ravel(), numpy.asarray(b).ravel()
    return numpy.linalg.norm(a - b) / numpy.linalg.norm(b)


@pytest.mark.timeout(600)
def test_coil_sharded_cuda_operators_and_cg(tmp_path):
    assert torch.cuda.is_available()
    world = 2
    out = str(tmp_path / 'res.pt')
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    parts = [torch.load(out + '.rank%d' % r) for r in range(world)]
    assert all(p['launches'] > 0 for p in parts)                      # the CUDA library did the work in both ranks
    assert [p['coils'] for p in parts] == [(0, 3), (3, 5)]
    om, sens, s, y = problem()

    import pynufft_b200
    from oracle import nufft_oracle as orc
    A = pynufft_b200.NUFFT('cuda:0')                                   # unsharded: all coils in one plan
    A.plan(om, ND, KD, JD, batch=B)
    A.set_sense(sens)
    O = orc.NUFFT()
    O.plan(om, ND, KD, JD, batch=B)
    O.set_sense(sens)

    y_cat = numpy.concatenate([p['y_fwd'].numpy() for p in parts], axis=1)
    assert rel(y_cat, O.forward_one2many(s)) < 1e-5
    assert numpy.array_equal(y_cat, A.forward_one2many(s))            # same kernels, same coils: identical
    for p in parts:                                                    # every rank holds the full reduced image
        assert rel(p['s_adj'].numpy(), O.adjoint_many2one(y)) < 1e-5
        assert rel(p['s_self'].numpy(), O.selfadjoint_one2many2one(s)) < 1e-5
    assert rel(parts[0]['s_adj'].numpy(), A.adjoint_many2one(y)) < 1e-6
    # CG: the sharded solver shares alpha / beta over all coils on all ranks (linalg/solve_hsa.py:555, 612-614, 640-643)
    x_cat = numpy.concatenate([p['x_cg'].numpy() for p in parts], axis=-1)
    x_one = A.solve(y, 'cg', maxiter=ITERS)
    assert x_cat.shape == x_one.shape == ND + (B,)
    assert rel(x_cat, x_one) < 2e-5, rel(x_cat, x_one)                # differs only in the order of the dot-product sums
    x64 = orc.solve_cg(O, y, ITERS, dtype=numpy.complex128)
    assert rel(x_cat, x64) < 5 * rel(orc.solve_cg(O, y, ITERS), x64) + 1e-5
