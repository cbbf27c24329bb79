"""
The multi-rank product path on a single-GPU box: two ranks (gloo rendezvous on 127.0.0.1, both computing on cuda:0)
run a mixed-size sweep with contiguous chain partitioning and one all_gather at the checkpoint; the gathered table
must equal the single-process result bit for bit (chains are independent and draw from per-chain random streams).
NCCL itself is exercised by `bench.py --gpus N` / `tools/run_multiset.py` on multi-GPU boxes.
"""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
OBS3 = ['sigmax', 'sigmay', 'sigmaz']


def _datasets():
    from effective_model_learning_b200 import configs, synthetic
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    ts = np.linspace(0, 2.5, 50)
    out = []
    for D in (1, 2):
        lay = synthetic.LibraryLayout(1, D, hyper)
        ctx = lay.context(ts, OBS3, device=0)
        ex, _ = ctx.evaluate(lay.random_params(np.random.default_rng(20260000 + 10 + D))[None], want_expect=True)
        out.append((ts, ex[0].real + np.random.default_rng(D).normal(0, 0.01, (3, 50)), OBS3))
    return out


def _sweep_table(steps):
    from effective_model_learning_b200.multiset import MultisetSweep
    sweep = MultisetSweep(_datasets(), config_numbers=[11, 7], defects_numbers=[1, 2], repetitions_numbers=[5, 3],
                          seed=3, device=0)
    sweep.run(steps)
    table = sweep.checkpoint()
    n = sweep.n_chains
    sweep.close()
    return n, table


def _worker(rank, world, port, steps, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK='0', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.cuda.set_device(0)
    n, table = _sweep_table(steps)
    q.put((rank, n, table))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    steps = 15
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, n, table = q.get(timeout=300)
        got[rank] = table
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    n, want = _sweep_table(steps)
    assert n == 2 * 2 * (5 + 3) and want.shape == (n, 8)
    assert np.array_equal(got[0], got[1])
    assert np.array_equal(got[0], want)
    assert np.all(want[:, 1] <= want[:, 0] + 1e-15) and np.all(want[:, 3] + want[:, 5] == steps)
