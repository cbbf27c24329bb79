"""Worst deviation of the CUDA path from the frozen exact oracle observables (tests/golden/oracle_observables.npz)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from effective_model_learning_b200 import configs, synthetic
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
z = np.load('tests/golden/oracle_observables.npz')
ts = z['times']
for Q, D in ((1, 1), (1, 2), (1, 3), (2, 3)):
    lay = synthetic.LibraryLayout(Q, D, hyper)
    ctx = lay.context(ts, ['sigmax', 'sigmay', 'sigmaz'])
    ex, _ = ctx.evaluate(lay.batch_params(z[f'seeds_Q{Q}_D{D}']), want_expect=True)
    print(f'd={2 ** (Q + D)} kernel={ctx.plan.kernel_name} models={len(ex)} max|GPU-exact|={np.abs(ex.real - z[f"expect_Q{Q}_D{D}"]).max():.2e} '
          f'max|imag|={np.abs(ex.imag).max():.1e}')
