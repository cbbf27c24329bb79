"""
One-off diagnostic behind tests/test_gpu_chain_statistics.py: a larger host population (LearningChain, np.random) against the
device population (BatchedChains), with bootstrap errors of the host quantiles.    python tools/chain_statistics.py [n_host]
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from effective_model_learning_b200 import BatchedChains, definitions  # noqa: E402
from effective_model_learning_b200.chains import MOVES  # noqa: E402
from tests.test_gpu_chain_statistics import _host_chain, _problem  # noqa: E402
from tests.helpers import OBS3  # noqa: E402

n_host = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D, steps, n_dev = 1, 200, 8192
cfg, ts, targets = _problem(D)
cfg['jump_length_rescaling_factor'] = 1.2
cfg['acceptance_window'] = 10
cfg['temperature_proposal'] = 0.005
cfg['chain_step_options'] = {m: (6 if m == MOVES[0] else 1) for m in MOVES}
h_best, h_cur = [], []
for c in range(n_host):
    np.random.seed(4000 + c)
    chain = _host_chain(cfg, ts, targets, D, steps - 1)
    with np.errstate(over='ignore'):
        chain.run()
    h_best.append(chain.best_loss); h_cur.append(chain.current_loss)
eng = BatchedChains(ts, targets, OBS3, n_chains=n_dev, n_defects=D, seed=9, defect_initial_state=definitions.ops['mm'], **cfg)
eng.run(steps)
st = eng.read()
rng = np.random.default_rng(0)
out = {}
for name, h, d in (('best', np.log10(h_best), np.log10(st['best_loss'])), ('current', np.log10(h_cur), np.log10(st['current_loss']))):
    for q in (0.25, 0.5, 0.75):
        boot = [np.quantile(rng.choice(h, len(h)), q) for _ in range(400)]
        boot64 = [np.quantile(rng.choice(h, 64), q) for _ in range(400)]
        out[f'{name}_q{q}'] = {'host': float(np.quantile(h, q)), 'host_boot_sd': float(np.std(boot)), 'sd_for_64_chains': float(np.std(boot64)),
                               'device': float(np.quantile(d, q))}
print(json.dumps(out, indent=1))
