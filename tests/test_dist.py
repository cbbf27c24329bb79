"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: partitioning and the checkpoint all_gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from effective_model_learning_b200 import dist as edist


def test_partition_covers_every_chain_once():
    for n in (0, 1, 7, 8, 4096, 8191):
        for world in (1, 2, 3, 8):
            spans = [edist.partition(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == edist.partition_sizes(n, world)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_chains, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    r, w = edist.init('gloo')
    assert (r, w) == (rank, world)
    lo, hi = edist.partition(n_chains, world, rank)
    chains = np.arange(lo, hi, dtype=np.float64)
    local = np.stack([chains * 10 + f for f in range(len(edist.CHAIN_STATS_FIELDS))], axis=1)
    table = edist.gather_chain_stats(local, n_chains)
    q.put((rank, table))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('n_chains', [9, 16])
def test_gather_chain_stats_two_ranks_gloo(n_chains):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_chains, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.stack([np.arange(n_chains) * 10.0 + f for f in range(len(edist.CHAIN_STATS_FIELDS))], axis=1)
    for r in range(world):
        assert np.array_equal(got[r], want)


def test_single_process_gather_is_identity():
    x = np.random.default_rng(0).normal(size=(5, len(edist.CHAIN_STATS_FIELDS)))
    assert np.array_equal(edist.gather_chain_stats(x), x)
    with pytest.raises(RuntimeError):
        edist.gather_chain_stats(np.zeros((3, 2)))


def test_enumerate_chains_follows_launcher_loop_order():
    from effective_model_learning_b200.multiset import enumerate_chains
    ch = enumerate_chains(2, [11, 0], [1, 2, 3], [2])
    assert len(ch) == 2 * 2 * 3 * 2
    assert [(c.dataset, c.config, c.defects, c.repetition) for c in ch[:5]] == \
        [(0, 11, 1, 1), (0, 11, 1, 2), (0, 11, 2, 1), (0, 11, 2, 2), (0, 11, 3, 1)]
    ch = enumerate_chains(1, [11], [1, 2], [3, 1])
    assert [(c.defects, c.repetition) for c in ch] == [(1, 1), (1, 2), (1, 3), (2, 1)]
    assert len(enumerate_chains(1, [11], [1, 2], [5, 5, 5])) == 6      # malformed repetitions -> default 3 each


def test_balanced_owners_split_every_group_over_the_ranks():
    sizes = [5, 3, 7, 1]
    for world in (1, 2, 3, 8):
        owners = edist.balanced_owners(sizes, world)
        assert len(owners) == world
        allidx = np.sort(np.concatenate(owners))
        assert np.array_equal(allidx, np.arange(sum(sizes)))           # every chain exactly once
        start = 0
        for n in sizes:                                                # each group: sizes differ by at most one
            per_rank = [int(((o >= start) & (o < start + n)).sum()) for o in owners]
            assert max(per_rank) - min(per_rank) <= 1 and sum(per_rank) == n
            start += n
        assert all(np.all(np.diff(o) > 0) for o in owners if len(o) > 1)


def _worker_owners(rank, world, port, sizes, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    edist.init('gloo')
    owners = edist.balanced_owners(sizes, world)
    mine = owners[rank].astype(np.float64)
    local = np.stack([mine * 10 + f for f in range(len(edist.CHAIN_STATS_FIELDS))], axis=1)
    q.put((rank, edist.gather_chain_stats(local, int(sum(sizes)), owners=owners)))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_gather_with_balanced_owners_two_ranks_gloo():
    world, sizes = 2, [5, 3, 7, 1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_owners, args=(r, world, port, sizes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.stack([np.arange(sum(sizes)) * 10.0 + f for f in range(len(edist.CHAIN_STATS_FIELDS))], axis=1)
    for r in range(world):
        assert np.array_equal(got[r], want)
