"""Does the batched engine actually learn?  Chains on simulated data (sigma = 0.01) should reach MSE ~ 1e-4."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from effective_model_learning_b200 import BatchedChains, configs, definitions, synthetic
OBS3 = ['sigmax', 'sigmay', 'sigmaz']
D = int(sys.argv[1]) if len(sys.argv) > 1 else 2
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
cfg = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
cfg['shock_anneal_at'] = int(0.7 * steps)
ts = np.linspace(0, 2.5, 200)
lay = synthetic.LibraryLayout(1, D, cfg)
truth = lay.random_params(np.random.default_rng(20260000 + 10 + D), p_present=0.3)
ctx = lay.context(ts, OBS3)
ex, _ = ctx.evaluate(truth[None], want_expect=True)
targets = ex[0].real + np.random.default_rng(1).normal(0, 0.01, (3, 200))
eng = BatchedChains(ts, targets, OBS3, n_chains=C, n_defects=D, seed=11, defect_initial_state=definitions.ops['mm'], **cfg)
t0 = time.time()
for chunk in range(5):
    eng.run(steps // 5)
    st = eng.read()
    print(f'{(chunk + 1) * steps // 5:6d} steps: best loss min {st["best_loss"].min():.3e} median {np.median(st["best_loss"]):.3e} '
          f'({time.time() - t0:.1f} s)', flush=True)
best = int(np.argmin(st['best_loss']))
print('noise floor (sigma^2) = 1.0e-04; truth processes:', int((truth[:lay.n_library] != 0).sum()) - D,
      '; best model processes:', int(st['stats'][best, 7]))
print('truth ', np.round(truth[:lay.n_library], 3))
print('best  ', np.round(st['best_params'][best][:lay.n_library], 3))
