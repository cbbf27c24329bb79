#!/usr/bin/env python
"""
One learning chain on simulated target data, end to end -- the flow of the reference's entry scripts
(execute_learning_full.py:89-226; its execute_learning_sim.py is broken upstream) with this package substituted:

    python examples/execute_learning_sim.py [experiment_name] [defects_count] [repetition] [max_iterations] [config_no]

Target data: a seeded ground-truth model from the chosen process library evaluated at 200 times on [0, 2.5] with
N(0, 0.01^2) noise (the reference loads such data from a pickle that is not in its repository).  Needs a GPU.
"""
import copy
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from effective_model_learning_b200 import (configs, definitions, learning_chain, output, predictive,  # noqa: E402
                                           synthetic)


def arg(i, cast, default):
    try:
        return cast(sys.argv[i])
    except (IndexError, ValueError):
        return default


experiment_name = arg(1, str, time.strftime('%Y_%m_%d_%H%M%S', time.gmtime()))
defects_count = arg(2, int, 1)
repetition_number = arg(3, int, 1)
max_iterations = arg(4, int, 1000)
configuration_number = arg(5, int, 11)

subexperiment_name = list(configs.specific_experiment_chain_hyperparams)[configuration_number]
config = copy.deepcopy(configs.default_chain_hyperparams)
config.update(configs.specific_experiment_chain_hyperparams[subexperiment_name])

measurement_observables = ['sigmax', 'sigmay', 'sigmaz']
ts = np.linspace(0.0, 2.5, 200)
layout = synthetic.LibraryLayout(1, defects_count, config)
truth = layout.model_from_params(layout.random_params(np.random.default_rng(20260000 + 10 + defects_count)))
noise = np.random.default_rng(repetition_number)
measurement_datasets = [d + noise.normal(0, 0.01, len(ts)) for d in truth.calculate_dynamics(ts, measurement_observables)]

filename = f'{experiment_name}_{subexperiment_name}_D{defects_count}_R{repetition_number}'
quest = learning_chain.LearningChain(target_times=ts, target_datasets=measurement_datasets,
                                     target_observables=measurement_observables, initial=(1, defects_count),
                                     qubit_initial_state=definitions.ops['plus'],
                                     defect_initial_state=definitions.ops['mm'],
                                     max_chain_steps=max_iterations, store_all_proposals=True, **config)
t0 = time.time()
with np.errstate(over='ignore'):
    best = quest.run()
print(f'{max_iterations + 1} proposals in {time.time() - t0:.2f} s, best loss {quest.best_loss:.3e}')
evaluation_ts, best_datasets = predictive.dense_grid_dynamics(best, ts, measurement_observables)


class Toggles:
    comparison = True
    loss = True
    log_posterior = True
    acceptance_probability = False
    acceptance = True
    graphs = False
    pickle = True
    text = True
    all_proposals = True
    hyperparams = True


out = output.Output(toggles=Toggles, filename=filename, dynamics_ts=[ts, evaluation_ts],
                    dynamics_datasets=[measurement_datasets, best_datasets],
                    dynamics_datasets_labels=['measurements', 'prediction'], dynamics_formatting=['b+', 'r-'],
                    observable_labels=measurement_observables, loss=quest.explored_loss, best_loss=quest.best_loss,
                    acceptance_probability=quest.explored_acceptance_probability,
                    overall_acceptance={'parameters tweak': quest.acc_tweak_steps / max(quest.tot_tweak_steps, 1),
                                        'reversible jump': quest.acc_RJ_steps / max(quest.tot_RJ_steps, 1)},
                    acceptance=quest.chain_windows_acceptance_log, models_to_save=[best], model_names=['best'],
                    chain_hyperparams=quest.get_init_hyperparams(), all_proposals=quest.all_proposals)
print('wrote', len(out.written), 'files, e.g.', os.path.basename(out.written[1]))
