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


def _cpulist(spec):
    cpus = set()
    for part in spec.strip().split(','):
        if '-' in part:
            a, b = part.split('-')
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(local_rank, local_world=1):
    """Pins the calling process to the CPU cores of one NUMA node, so that pinned host buffers allocated afterwards are
    first-touched there: with one process per GPU all writing results to host memory, buffers that all live on one node
    cap the aggregate device-to-host rate at that node's memory bandwidth.  The node is the one the GPU hangs off
    (sysfs numa_node of its PCI function) when the container exposes it, else the ranks are spread evenly over the
    nodes that are visible.  Returns (node id, how) or (None, reason); nothing is changed in the latter case."""
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir('/sys/devices/system/node') if d.startswith('node') and d[4:].isdigit())
    except OSError:
        return None, 'no /sys/devices/system/node'
    if len(nodes) < 2:
        return None, f'{len(nodes)} NUMA node(s) visible'
    node, how = None, ''
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f'{getattr(props, "pci_domain_id", 0):04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as fh:
            v = int(fh.read().strip())
        if v in nodes:
            node, how = v, f'sysfs numa_node of {bdf}'
    except (OSError, ValueError, RuntimeError, AttributeError):
        pass
    if node is None:
        node, how = nodes[(local_rank * len(nodes)) // max(local_world, 1) % len(nodes)], 'ranks spread over the visible nodes'
    try:
        with open(f'/sys/devices/system/node/node{node}/cpulist') as fh:
            allowed = _cpulist(fh.read()) & set(os.sched_getaffinity(0))
        if not allowed:
            return None, f'no allowed core on node {node}'
        os.sched_setaffinity(0, allowed)
    except (OSError, ValueError):
        return None, 'cpulist unreadable'
    return node, how


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
