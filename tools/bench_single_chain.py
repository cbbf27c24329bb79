"""C1 (BASELINE.json configs[0]): one chain, 1 qubit + D defects, configs.py defaults, ~1000 RJMCMC steps, wall clock."""
import cProfile
import json
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import effective_model_learning_b200 as eml  # noqa: E402
from effective_model_learning_b200 import configs, definitions, synthetic  # noqa: E402

OBS3 = ['sigmax', 'sigmay', 'sigmaz']
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfg = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
cfg['iterations_till_progress_update'] = False
ts = np.linspace(0, 2.5, 200)
lay = synthetic.LibraryLayout(1, D, cfg)
ctx = lay.context(ts, OBS3)
ex, _ = ctx.evaluate(lay.random_params(np.random.default_rng(20260000 + 10 + D))[None], want_expect=True)
targets = [ex[0, m].real + np.random.default_rng(m).normal(0, 0.01, 200) for m in range(3)]
np.random.seed(1000)
chain = eml.LearningChain(target_times=ts, target_datasets=targets, target_observables=OBS3, initial=(1, D),
                          qubit_initial_state=definitions.ops['plus'], defect_initial_state=definitions.ops['mm'],
                          max_chain_steps=steps, **cfg)
with np.errstate(over='ignore'):
    chain.run(20)
profile = len(sys.argv) > 3 and sys.argv[3] == 'profile'      # cProfile slows the host loop: the steps/s figure is taken without it
prof = cProfile.Profile()
t0 = time.perf_counter()
if profile:
    prof.enable()
with np.errstate(over='ignore'):
    chain.run(steps)
if profile:
    prof.disable()
dt = time.perf_counter() - t0
print(json.dumps({'config': f'C1 single chain, 1 qubit + {D} defects', 'steps': steps + 1, 'seconds': dt,
                  'steps_per_s': (steps + 1) / dt, 'best_loss': chain.best_loss, 'profiled': profile}))
if profile:
    pstats.Stats(prof).sort_stats('cumulative').print_stats(14)
