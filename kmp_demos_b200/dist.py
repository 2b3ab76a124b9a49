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


def bind_to_gpu_numa_node(local_rank):
    """Pins the calling process to the CPU cores of the NUMA node its GPU hangs off (Linux sysfs), so that pinned host
    buffers allocated afterwards are first-touched on that node: with one process per GPU all writing results to host
    memory, buffers that all live on one node cap the aggregate device-to-host rate at that node's memory bandwidth.
    Returns the node id, or None when the topology cannot be read (nothing is changed then)."""
    try:
        import torch
        bdf = None
        props = torch.cuda.get_device_properties(local_rank)
        if hasattr(props, 'pci_bus_id') and hasattr(props, 'pci_device_id'):
            dom = getattr(props, 'pci_domain_id', 0)
            bdf = f'{dom:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        if bdf is None:
            return None
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f'/sys/devices/system/node/node{node}/cpulist') as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, RuntimeError, AttributeError):
        return None


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
