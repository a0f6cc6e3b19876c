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


def direction_shard(Nd, rank, world):
    """Contiguous, balanced block ``[d0, d1)`` of the direction axis for ``rank`` (like ``time_shard``)."""
    return time_shard(Nd, rank, world)


def union_index(row_voxels, V, group=None):
    """Common compact numbering of the voxels that ANY rank's rays touch.

    ``row_voxels``: ascending voxel indices of this rank's non-empty operator rows (any integer dtype, on the
    device the collective backend works with).  Returns ``(row_dst, union_voxels, n_union)``: the position of
    every local row in the union (int32), the ascending voxel index of every union position (int32), and the
    union's size.  One MAX-allreduce of a V-byte mask, once per geometry."""
    mask = torch.zeros(int(V), dtype=torch.uint8, device=row_voxels.device)
    mask[row_voxels.long()] = 1
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mask, op=dist.ReduceOp.MAX, group=group)
    union_voxels = torch.nonzero(mask).reshape(-1).to(torch.int32)
    pos = torch.cumsum(mask, 0, dtype=torch.int32) - 1
    row_dst = pos[row_voxels.long()].contiguous()
    return row_dst, union_voxels.contiguous(), int(union_voxels.numel())
