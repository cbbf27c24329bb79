"""RJMCMC steps/s of BatchedChains for several (D, chains) shapes, chains started from random library models."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from effective_model_learning_b200 import BatchedChains, configs, definitions, synthetic
OBS3 = ['sigmax', 'sigmay', 'sigmaz']
cfg = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
ts = np.linspace(0, 2.5, 200)
import os
SHAPES = ((3, 4096, 100),) if os.environ.get("EML_CHAIN_GROUPS") else ((1, 1024, 200), (1, 8192, 100), (2, 1024, 200), (2, 8192, 100), (3, 512, 100), (3, 4096, 100))
if len(sys.argv) > 1:       # shapes on the command line: D:chains:steps ...
    SHAPES = tuple(tuple(int(x) for x in a.split(':')) for a in sys.argv[1:])
for D, C, steps in SHAPES:
    lay = synthetic.LibraryLayout(1, D, cfg)
    ctx = lay.context(ts, OBS3)
    ex, _ = ctx.evaluate(lay.random_params(np.random.default_rng(20260000 + 10 + D))[None], want_expect=True)
    targets = ex[0].real + np.random.default_rng(1).normal(0, 0.01, (3, 200))
    eng = BatchedChains(ts, targets, OBS3, n_chains=C, n_defects=D, seed=5, initial_params=lay.batch_params(1000 + np.arange(C)),
                        defect_initial_state=definitions.ops['mm'], **cfg)
    eng.run(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); eng.run(steps); e1.record(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    print(json.dumps({'D': D, 'chains': C, 'steps_per_s': C * steps / (ms * 1e-3), 'ms_per_step': ms / steps,
                      'host_enqueue_ms_per_step': 1e3 * wall / steps}))
    eng.close()
