"""Tiny end-to-end case for compute-sanitizer: every kernel once, small sizes."""
import sys
import numpy as np
sys.path.insert(0, '.')
from effective_model_learning_b200 import _native, configs, definitions, synthetic, BatchedChains

OBS3 = ['sigmax', 'sigmay', 'sigmaz']
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
ts = np.linspace(0, 1.0, 12)
for Q, D, kernel, B in ((1, 1, _native.KERNEL_AUTO, 5), (1, 2, _native.KERNEL_AUTO, 5), (1, 3, _native.KERNEL_AUTO, 9),
                        (1, 3, _native.KERNEL_GENERIC, 3), (2, 3, _native.KERNEL_AUTO, 2)):
    lay = synthetic.LibraryLayout(Q, D, hyper)
    ctx = lay.context(ts, OBS3, kernel=kernel)
    ctx.set_targets(np.zeros((3, len(ts))))
    ex, loss = ctx.evaluate(lay.batch_params(1000 + np.arange(B)), want_expect=True, want_loss=True)
    print(Q, D, ctx.plan.kernel_name, float(loss.sum()))
cfg = dict(hyper)
cfg['shock_anneal_at'] = 2
eng = BatchedChains(ts, np.zeros((3, len(ts))), OBS3, n_chains=8, n_defects=3, seed=1, defect_initial_state=definitions.ops['mm'], **cfg)
eng.run(4)
print(eng.read()['stats'][:, :2].sum())
eng.close()
