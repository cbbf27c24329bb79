#!/usr/bin/env python
"""
The reference's multiset launcher (launch_execute_learning_multiset.sh: datasets x configurations x defect numbers x
repetitions, one OS process per chain) as ONE batched job on the GPU(s) of a node:

    python examples/execute_learning_multiset.py [experiment_name] [iterations] [repetitions]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/execute_learning_multiset.py ...

Every rank owns a contiguous slice of the chains; at the end (and at every checkpoint) the per-chain statistics are
gathered on all ranks with one all_gather, and each rank writes `<experiment>_<dataset>_conf<C>_D<D>_R<R>_best.pickle`
(+ .txt) for the chains it owns -- the file names of the reference's launcher.  Target data are simulated as in
examples/execute_learning_sim.py.  Needs a GPU.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from effective_model_learning_b200 import configs, definitions, dist, output, synthetic  # noqa: E402
from effective_model_learning_b200.multiset import MultisetSweep  # noqa: E402

experiment_name = sys.argv[1] if len(sys.argv) > 1 else 'multiset-demo'
iterations_number = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
repetitions_number = int(sys.argv[3]) if len(sys.argv) > 3 else 64
defects_numbers = (1, 2, 3)
config_numbers = (11,)
checkpoint_every = 500
measurement_observables = ['sigmax', 'sigmay', 'sigmaz']

rank, world = dist.init()
local = dist.rank_world()[2]
torch.cuda.set_device(local)

# simulated target datasets (the reference digitises them from csv files that are not in its repository)
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
ts = np.linspace(0.0, 2.5, 200)
dataset_names, datasets = [], []
for D_truth in (1, 2):
    layout = synthetic.LibraryLayout(1, D_truth, hyper)
    truth = layout.model_from_params(layout.random_params(np.random.default_rng(20260000 + 10 + D_truth)))
    noise = np.random.default_rng(D_truth)
    data = np.stack([d + noise.normal(0, 0.01, len(ts)) for d in truth.calculate_dynamics(ts, measurement_observables)])
    dataset_names.append(f'sim-D{D_truth}')
    datasets.append((ts, data, measurement_observables))

sweep = MultisetSweep(datasets, config_numbers, defects_numbers, [repetitions_number], seed=2024, device=local,
                      qubit_initial_state=definitions.ops['plus'], defect_initial_state=definitions.ops['mm'])
done = 0
while done < iterations_number:
    n = min(checkpoint_every, iterations_number - done)
    sweep.run(n)
    done += n
    table = sweep.checkpoint()                      # [n_chains][8] on every rank: ONE all_gather
    if rank == 0:
        print(f'{done:7d} proposals/chain: median best loss {np.median(table[:, 1]):.3e}, '
              f'acceptance {(table[:, 2].sum() + table[:, 4].sum()) / max(1.0, table[:, 3].sum() + table[:, 5].sum()):.2f}',
              flush=True)


class Toggles:
    pickle = True
    text = True


for gidx in (int(g) for g in sweep.owned):
    spec = sweep.chains[gidx]
    best = sweep.best_models(gidx)
    filename = f'{experiment_name}_{dataset_names[spec.dataset]}_conf{spec.config}_D{spec.defects}_R{spec.repetition}'
    output.Output(toggles=Toggles, filename=filename, models_to_save=[best], model_names=['best'])
if rank == 0:
    print(f'{sweep.n_chains} chains on {world} rank(s); rank 0 wrote {len(sweep.owned)} best-model pickles')
sweep.close()
if world > 1:
    torch.distributed.destroy_process_group()
