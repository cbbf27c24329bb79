import sys, numpy as np, torch
sys.path.insert(0, '.')
from effective_model_learning_b200 import configs, synthetic
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
ts = np.linspace(0, 2.5, 200)
lay = synthetic.LibraryLayout(2, 3, hyper)
ctx = lay.context(ts, ['sigmax', 'sigmay', 'sigmaz'])
ctx.set_targets(np.zeros((3, 200)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
p = lay.batch_params(1000 + np.arange(B))
for _ in range(2):
    ctx.evaluate(p, want_expect=False, want_loss=True)
print('ok', ctx.plan.kernel_name)
