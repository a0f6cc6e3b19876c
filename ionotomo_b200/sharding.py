"""Ray sharding across GPUs (one process per GPU, torch.distributed).

Rays are independent except for (1) ``tec - tec[i0]`` (needs the reference-antenna ray of
the same (time, direction)) and (2) the sum over rays in the adjoint.  Sharding the TIME
axis keeps antenna ``i0`` local to every shard, so the forward needs no communication and
the adjoint needs exactly one sum of the voxel accumulator across ranks (NCCL allreduce
over NVLink on the GPU box; gloo in the CPU tests).  The density grid is replicated.
"""
import torch
import torch.distributed as dist


def time_shard(Nt, rank, world):
    """Contiguous, balanced block ``[t0, t1)`` of the time axis for ``rank``."""
    base, rem = divmod(Nt, world)
    t0 = rank * base + min(rank, rem)
    return t0, t0 + base + (1 if rank < rem else 0)


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum_(t, group=None):
    """In-place sum over ranks (no-op for a single process). Returns ``t``."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def sharded_misfit(S_local, group=None):
    """Sum of the per-shard misfits as a python float (0-d tensor in, float out)."""
    s = S_local.detach().clone().reshape(1)
    allreduce_sum_(s, group)
    return float(s[0])


def allreduce_sum_async(t, group=None):
    """Start an in-place sum over ranks and return the work handle (``None`` for one process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return None
