"""Multi-GPU partitioning of a batch of independent OCPs.

Every OCP is a self-contained solve, so the batch is split into contiguous index ranges, one per rank, with NO collective
on the solver path.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only at the edges: scatter of the
initial states / warm starts from rank 0 and gather of the solutions.
"""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous [begin, end) of OCP indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _who(dist, world, rank):
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    return rank, world


def scatter_rows(array_on_rank0, total, row_shape, dist, device="cpu", dtype=None, world=None, rank=None):
    """Rank 0 holds `array_on_rank0` (total x row_shape; a numpy array or a tensor already on `device`); every rank
    receives its shard_range rows.  Point-to-point sends from rank 0 (NCCL on GPUs): no collective, nothing on the
    solver path.  A single process (no process group) gets the whole array."""
    import torch
    rank, world = _who(dist, world, rank)
    dtype = dtype or torch.float64
    b, e = shard_range(total, rank, world)
    out = torch.empty((e - b,) + tuple(row_shape), dtype=dtype, device=device)
    if rank == 0:
        if torch.is_tensor(array_on_rank0):
            src = array_on_rank0.to(device=device, dtype=dtype)
        else:
            src = torch.as_tensor(np.ascontiguousarray(array_on_rank0), dtype=dtype, device=device)
        # one batched group of sends (NCCL: a single ncclGroup, the transfers to all peers run concurrently over NVSwitch
        # instead of one after the other); row slices of a contiguous array are contiguous: no staging copies
        ops = []
        for r in range(1, world):
            rb, re = shard_range(total, r, world)
            if re > rb:
                ops.append(dist.P2POp(dist.isend, src[rb:re], r))
        out.copy_(src[b:e])
        for q in (dist.batch_isend_irecv(ops) if ops else []):
            q.wait()
    elif e > b:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, out, 0)]):
            q.wait()
    return out


def gather_rows(local, total, dist, world=None, rank=None):
    """Inverse of scatter_rows: rank 0 returns the (total x row_shape) tensor, other ranks None."""
    import torch
    rank, world = _who(dist, world, rank)
    if rank == 0:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        b, e = shard_range(total, 0, world)
        out[b:e].copy_(local)
        ops = []   # every peer's shard lands straight in its rows of `out`, all receives in one batched group
        for r in range(1, world):
            rb, re = shard_range(total, r, world)
            if re > rb:
                ops.append(dist.P2POp(dist.irecv, out[rb:re], r))
        for q in (dist.batch_isend_irecv(ops) if ops else []):
            q.wait()
        return out
    if local.shape[0] > 0:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), 0)]):
            q.wait()
    return None
