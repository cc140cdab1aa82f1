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
        reqs = []
        for r in range(1, world):
            rb, re = shard_range(total, r, world)
            if re > rb:
                reqs.append(dist.isend(src[rb:re].contiguous(), dst=r))
        out.copy_(src[b:e])
        for q in reqs:
            q.wait()
    elif e > b:
        dist.recv(out, src=0)
    return out


def gather_rows(local, total, dist, world=None, rank=None):
    """Inverse of scatter_rows: rank 0 returns the (total x row_shape) tensor, other ranks None."""
    import torch
    rank, world = _who(dist, world, rank)
    if rank == 0:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        b, e = shard_range(total, 0, world)
        out[b:e].copy_(local)
        for r in range(1, world):
            rb, re = shard_range(total, r, world)
            if re > rb:
                buf = torch.empty((re - rb,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
                dist.recv(buf, src=r)
                out[rb:re].copy_(buf)
        return out
    if local.shape[0] > 0:
        dist.send(local.contiguous(), dst=0)
    return None
