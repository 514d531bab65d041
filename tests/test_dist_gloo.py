"""
CPU-only, world_size = 2, gloo: the multi-GPU host logic of pynufft_b200/dist.py and the CG driver
(coil sharding, where the all-reduces sit, weighting of the many2one mean).  The local operator is an
oracle-backed stand-in with the NUFFT batch API on torch CPU tensors; the vector ops are a torch
restatement of csrc/solver.cu.  Checked against the single-process oracle on all coils.
"""
import os
import socket
import sys

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import nufft_oracle as orc  # noqa: E402

ND, KD, JD, B, M = (12, 10), (24, 20), (4, 4), 5, 300


def problem():
    rng = numpy.random.default_rng(7)
    om = rng.uniform(-numpy.pi, numpy.pi, (M, 2))
    sens = (rng.standard_normal(ND + (B,)) + 1j * rng.standard_normal(ND + (B,))).astype(numpy.complex64)
    s = (rng.standard_normal(ND) + 1j * rng.standard_normal(ND)).astype(numpy.complex64)
    return om, sens, s


class OracleLocal:
    """Stand-in for pynufft_b200.NUFFT (batch API subset) running the oracle on CPU tensors."""

    def __init__(self, om, nb):
        self.O = orc.NUFFT()
        self.O.plan(om, ND, KD, JD, batch=nb)
        self.batch = nb
        self.ndims = len(ND)

    def set_sense(self, sens):
        self.O.set_sense(numpy.asarray(sens))

    def forward_one2many(self, s):
        return torch.from_numpy(self.O.forward_one2many(s.numpy()).astype(numpy.complex64))

    def adjoint_many2one(self, y):
        return torch.from_numpy(self.O.adjoint_many2one(y.numpy()).astype(numpy.complex64))


class TorchVectorOps:
    """torch restatement of the fused CG kernels (csrc/solver.cu) for the CPU test."""

    def scalars(self, device):
        return torch.zeros(6, dtype=torch.float64, device=device)

    @staticmethod
    def _set(out, z):
        out[0], out[1] = float(z.real), float(z.imag)

    def cg_init(self, b, Ax, r, p, rsold):
        r.copy_(b - Ax)
        p.copy_(r)
        self._set(rsold, torch.vdot(r.reshape(-1), r.reshape(-1)))

    def dotc(self, a, b, out):
        self._set(out, torch.vdot(a.reshape(-1), b.reshape(-1)))

    def update_xr(self, x, r, p, Ap, rsold, pAp, rsnew):
        alpha = complex(rsold[0], rsold[1]) / complex(pAp[0], pAp[1])
        x.add_(p, alpha=alpha)
        r.sub_(Ap, alpha=alpha)
        self._set(rsnew, torch.vdot(r.reshape(-1), r.reshape(-1)))

    def update_p(self, p, r, rsnew, rsold):
        beta = complex(rsnew[0], rsnew[1]) / complex(rsold[0], rsold[1])
        p.mul_(beta).add_(r)


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pynufft_b200.dist import CoilShardedNUFFT, shard_coils
        from pynufft_b200.solve import cg_kspace
        om, sens, s = problem()
        sl = shard_coils(B, world, rank)
        local = OracleLocal(om, sl.stop - sl.start)
        op = CoilShardedNUFFT(local, B)
        assert op.coils == sl
        op.set_sense(sens)
        y_loc = op.forward_one2many(torch.from_numpy(s))
        s2 = op.adjoint_many2one(y_loc)
        s3 = op.selfadjoint_one2many2one(torch.from_numpy(s))
        # k-space CG with shared alpha/beta across the shards
        O = local.O
        b = torch.from_numpy(O.y2k(y_loc.numpy()).astype(numpy.complex64))
        G = lambda v: torch.from_numpy(O.y2k(O.k2y(v.numpy())).astype(numpy.complex64))
        x = cg_kspace(G, b, 4, TorchVectorOps(), allreduce=lambda t: dist.all_reduce(t))
        if rank == 0:
            torch.save({'s2': s2, 's3': s3}, out)
        torch.save({'y': y_loc, 'x': x, 'coils': (sl.start, sl.stop)}, out + '.rank%d' % rank)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        return sk.getsockname()[1]


def rel(a, b):
    a, b = numpy.asarray(a).ravel(), numpy.asarray(b).ravel()
    return numpy.linalg.norm(a - b) / numpy.linalg.norm(b)


def test_shard_coils_partition():
    from pynufft_b200.dist import shard_coils
    for total in (1, 5, 8, 32, 33):
        for world in (1, 2, 3, 8):
            sl = [shard_coils(total, world, r) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == total
            assert all(a.stop == b.start for a, b in zip(sl[:-1], sl[1:]))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_coil_sharded_operators_world2(tmp_path):
    world = 2
    out = str(tmp_path / 'res.pt')
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    om, sens, s = problem()
    O = orc.NUFFT()
    O.plan(om, ND, KD, JD, batch=B)
    O.set_sense(sens)
    y_all = O.forward_one2many(s).astype(numpy.complex64)
    r0 = torch.load(out)
    parts = [torch.load(out + '.rank%d' % r) for r in range(world)]
    y_cat = numpy.concatenate([p['y'].numpy() for p in parts], axis=1)
    assert rel(y_cat, y_all) < 1e-6                                       # forward: no collective, blocks tile the coils
    assert rel(r0['s2'].numpy(), O.adjoint_many2one(y_all)) < 1e-5        # one all-reduce, weighted partial means
    assert rel(r0['s3'].numpy(), O.selfadjoint_one2many2one(s)) < 1e-5
    # CG: same iterates as the single-process batched CG (one alpha/beta for all coils)
    c64 = numpy.complex64
    Gall = lambda v: O.y2k(O.k2y(v)).astype(c64)
    b = O.y2k(y_all).astype(c64)
    x = b.copy()
    r = b - Gall(x)
    p = r.copy()
    rs = numpy.vdot(r, r)
    for _ in range(4):
        Ap = Gall(p)
        a = rs / numpy.vdot(p, Ap)
        x = x + a * p
        r = r - a * Ap
        rn = numpy.vdot(r, r)
        p = r + (rn / rs) * p
        rs = rn
    x_cat = numpy.concatenate([p_['x'].numpy() for p_ in parts], axis=-1)
    assert rel(x_cat, x) < 1e-4
