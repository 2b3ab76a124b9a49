"""One process per GPU over torch.distributed (NCCL on the B200 box, gloo in CPU tests).

Two shardings exist on the hot path:
  * independent KMP problems: a contiguous static split, no collective (``shard_range``);
  * EM over row-partitioned samples: one all-reduce (sum) of the packed sufficient statistics per iteration
    (``allreduce_stats``; mixture.EMEngine calls torch.distributed.all_reduce on the same buffer).
"""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))


def init_process_group(backend=None):
    """Initialises torch.distributed from the torchrun environment (RANK/WORLD_SIZE/MASTER_*)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world == 1:
        return rank, world, local
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        kw = {}
        if backend == 'nccl':
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_items, rank, world):
    """[lo, hi) of a contiguous static split of n_items units over `world` ranks."""
    per, rem = divmod(n_items, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def allreduce_stats(stats, group=None):
    """Sums the packed EM statistics [N_k, sum r x, sum r x x^T (packed), sum log-lik] over the ranks."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def max_over_ranks(value_ms, device):
    """MAX over ranks of a device-timed duration (bench.py's timing rule)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
