"""
d = 4 evaluator: two models per warp (eval_pack2_kernel, launches with more models than warp slots) against one model per warp
(EML_PACK2=0).  Run once per setting; writes observables / losses / status to gpurun_out/pack2_<tag>.npz and prints the timing.
Usage: [EML_PACK2=0] python tools/pack2_check.py [--chains 5001] [--multi-step]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from effective_model_learning_b200 import configs, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--chains', type=int, default=5001)
ap.add_argument('--span', type=float, default=2.5)
ap.add_argument('--reps', type=int, default=20)
a = ap.parse_args()
tag = ('one' if os.environ.get('EML_PACK2') == '0' else 'two') + f'_{a.span:g}'
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
ts = np.linspace(0, a.span, 200)
lay = synthetic.LibraryLayout(1, 1, hyper)
ctx = lay.context(ts, ['sigmax', 'sigmay', 'sigmaz'])
rng = np.random.default_rng(3)
ctx.set_targets(rng.normal(0, 0.3, (3, 200)))
params = lay.batch_params(1000 + np.arange(a.chains))
params[7] = np.nan          # a model that cannot be evaluated must not disturb its partner
plan = ctx.plan
p = torch.from_numpy(params).cuda()
lo = torch.zeros(a.chains, dtype=torch.float64, device='cuda')
st = torch.zeros(a.chains, dtype=torch.int32, device='cuda')
exd = torch.zeros((a.chains, 3, 200, 2), dtype=torch.float64, device='cuda')
s = torch.cuda.current_stream().cuda_stream
plan.eval_device(a.chains, p.data_ptr(), 0, exd.data_ptr(), lo.data_ptr(), st.data_ptr(), s)
torch.cuda.synchronize()
os.makedirs('gpurun_out', exist_ok=True)
np.savez(f'gpurun_out/pack2_{tag}.npz', expect=exd.cpu().numpy(), loss=lo.cpu().numpy(), status=st.cpu().numpy())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    plan.eval_device(a.chains, p.data_ptr(), 0, 0, lo.data_ptr(), st.data_ptr(), s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(json.dumps({'tag': tag, 'kernel': plan.kernel_name, 'models': a.chains, 'ms': ms, 'evals_per_s': a.chains / ms * 1e3,
                  'status_min': int(st.min()), 'bad': int((st < 0).sum()), 'mean_apps': float(st[st >= 0].float().mean())}))
