"""
Multi-GPU plumbing: chains are independent (the reference runs each as its own OS process,
launch_execute_learning_multiset.sh:60-63), so the chain range is split contiguously over the ranks and the
hot path needs no collective.  The only exchange is one all_gather of a small per-chain statistics record at
checkpoints, so that every rank can write the global summary.

One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

# per-chain record gathered at checkpoints; counters are the reference's trackers (learning_chain.py:308-311, 380)
CHAIN_STATS_FIELDS = ('current_loss', 'best_loss', 'acc_tweak', 'tot_tweak', 'acc_RJ', 'tot_RJ',
                      'last_window_acceptance', 'n_processes')


def rank_world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def init(backend: str | None = None, device: torch.device | None = None) -> tuple[int, int]:
    """Initialise the default process group if WORLD_SIZE > 1 (idempotent).  Returns (rank, world)."""
    rank, world, local = rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kwargs = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kwargs['device_id'] = device if device is not None else torch.device('cuda', local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world


def partition(n_chains: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous chain range [lo, hi) owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_chains, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def partition_sizes(n_chains: int, world: int) -> list[int]:
    return [partition(n_chains, world, r)[1] - partition(n_chains, world, r)[0] for r in range(world)]


def balanced_owners(group_sizes, world: int) -> list[np.ndarray]:
    """
    Cost-balanced ownership of a sweep whose chains come in groups of equal cost (one group per (dataset, configuration,
    number of defects), chains of a group contiguous in the global order): EVERY group is split contiguously over the
    ranks, so each rank gets 1/world of every model size instead of a contiguous slice of the launcher order (which
    gives one rank mostly d = 16 chains and another mostly d = 4 chains: d = 16 costs ~3.5x d = 4).
    Returns, for each rank, the sorted global indices of the chains it owns.
    """
    owners = [[] for _ in range(world)]
    start = 0
    for n in group_sizes:
        for r in range(world):
            lo, hi = partition(int(n), world, r)
            owners[r].append(np.arange(start + lo, start + hi, dtype=np.int64))
        start += int(n)
    return [np.concatenate(o) if o else np.zeros(0, dtype=np.int64) for o in owners]


def gather_chain_stats(local_stats, n_chains: int | None = None, owners=None) -> np.ndarray:
    """
    local_stats: [n_local][len(CHAIN_STATS_FIELDS)] float64 (numpy array or torch tensor on the rank's device).
    Returns the [n_chains][8] table of ALL chains, in global chain order, on every rank -- ONE all_gather.
    owners (optional): for every rank the global indices of its rows (e.g. `balanced_owners`); default = the
    contiguous `partition` of n_chains.
    """
    t = local_stats if isinstance(local_stats, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_stats))
    t = t.to(torch.float64)
    if t.ndim != 2 or t.shape[1] != len(CHAIN_STATS_FIELDS):
        raise RuntimeError('chain statistics must be [n_local][%d]' % len(CHAIN_STATS_FIELDS))
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t.cpu().numpy()
    world, rank = dist.get_world_size(), dist.get_rank()
    if n_chains is None:
        raise RuntimeError('n_chains (global) is required when world_size > 1')
    if owners is not None:
        if len(owners) != world or sum(len(o) for o in owners) != n_chains:
            raise RuntimeError('owners must list every chain exactly once, one index array per rank')
        sizes = [len(o) for o in owners]
    else:
        sizes = partition_sizes(n_chains, world)
    if t.shape[0] != sizes[rank]:
        raise RuntimeError(f'rank {rank} holds {t.shape[0]} chains, partition says {sizes[rank]}')
    width = max(sizes)
    if dist.get_backend() == 'nccl' and not t.is_cuda:
        t = t.cuda()
    padded = torch.zeros((width, t.shape[1]), dtype=torch.float64, device=t.device)
    padded[:t.shape[0]] = t
    out = torch.empty((world * width, t.shape[1]), dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, padded)          # concatenation along dim 0 (works for NCCL and gloo)
    out = out.view(world, width, t.shape[1]).cpu().numpy()
    if owners is None:
        return np.concatenate([out[r, :sizes[r]] for r in range(world)], axis=0)
    table = np.empty((n_chains, t.shape[1]))
    for r in range(world):
        table[np.asarray(owners[r], dtype=np.int64)] = out[r, :sizes[r]]
    return table
