"""
Coil-sharded multi-GPU operators: one process per GPU (torch.distributed, NCCL over
NVLink), every rank plans the same trajectory with batch = its contiguous block of coils.

Collectives exist only where the path has a real exchange:
  * forward_one2many         : none (the image s is replicated)
  * adjoint_many2one         : ONE all-reduce(sum) of the Nd image (mean over coils = sum of the
                               per-rank partial means weighted by B_local / B_total,
                               linalg/nufft_hsa.py:628-656, cAggregate re_subroutine.py:441-497)
  * solve 'cg' (k-space CG)  : two 16-byte all-reduces per iteration (the dot products behind the shared
                               alpha / beta, linalg/solve_hsa.py:555,612-614,640-643); scalars stay on the
                               device
The reference itself has no multi-device support (doc/source/installation/init.rst:11,33).
The functions below only need `local` to expose the NUFFT batch API on torch tensors, so that the
collective placement is testable on CPU with the gloo backend.
"""
import torch
import torch.distributed as dist


def shard_coils(total, world, rank):
    """Contiguous block of coils owned by `rank`: the first `total % world` ranks take one extra."""
    base, extra = divmod(int(total), int(world))
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


class CoilShardedNUFFT:
    """
    local       : an operator planned with batch = number of local coils (pynufft_b200.NUFFT)
    total_batch : coils over all ranks
    group       : torch.distributed process group (None = default group)
    """

    def __init__(self, local, total_batch, group=None):
        self.local = local
        self.total_batch = int(total_batch)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.coils = shard_coils(self.total_batch, self.world, self.rank)
        nloc = self.coils.stop - self.coils.start
        if nloc != local.batch:
            raise ValueError('local operator has batch=%d but this rank owns %d coils' % (local.batch, nloc))

    def set_sense(self, coil_profile_all):
        """coil_profile_all: Nd+(total_batch,) array; each rank keeps its own block."""
        self.local.set_sense(coil_profile_all[..., self.coils])

    def forward_one2many(self, s):
        return self.local.forward_one2many(s)                    # y (M, B_local): no collective

    def adjoint_many2one(self, y_local):
        s = self.local.adjoint_many2one(y_local)                 # mean over the LOCAL coils
        s = s * (self.local.batch / self.total_batch)
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=self.group)
        return s

    def selfadjoint_one2many2one(self, s):
        return self.adjoint_many2one(self.forward_one2many(s))

    def solve_cg(self, y_local, maxiter=30):
        """k-space CG with alpha/beta shared by all coils on all ranks; returns the local Nd+(B_local,) block."""
        from .solve import cg
        return cg(self.local, y_local, maxiter=maxiter, group=self.group if self.group is not None else dist.group.WORLD)
