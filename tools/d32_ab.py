"""
d = 32 evaluator A/B: the orbit kernel (eml_cta_orbit.cu, default) against the first-generation tile-row kernel
(EML_D32_TILEROW=1), same batch, device-resident parameters, CUDA events.  Run once per variant (the switch is read once per
process); with --check the losses / statuses are written to gpurun_out/d32_<variant>.npz for comparison.
Usage: [EML_D32_TILEROW=1] python tools/d32_ab.py [--chains 512] [--reps 5]
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
ap.add_argument('--chains', type=int, nargs='+', default=[512, 592, 2048])
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--check', action='store_true')
args = ap.parse_args()
variant = 'tilerow' if os.environ.get('EML_D32_TILEROW') == '1' else 'orbit'
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
ts = np.linspace(0, 2.5, 200)
lay = synthetic.LibraryLayout(2, 3, hyper)
ctx = lay.context(ts, ['sigmax', 'sigmay', 'sigmaz'])
ctx.set_targets(np.zeros((3, 200)))
plan = ctx.plan
for B in args.chains:
    p = torch.from_numpy(lay.batch_params(1000 + np.arange(B))).cuda()
    loss = torch.zeros(B, dtype=torch.float64, device='cuda')
    st = torch.zeros(B, dtype=torch.int32, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print(json.dumps({'variant': variant, 'kernel': plan.kernel_name, 'models': B, 'ms_per_batch': ms, 'evals_per_s': B / ms * 1e3,
                      'applications_per_eval': float(st.float().mean()), 'loss_sum': float(loss.sum())}))
    if args.check:
        os.makedirs('gpurun_out', exist_ok=True)
        np.savez(f'gpurun_out/d32_{variant}_{B}.npz', loss=loss.cpu().numpy(), status=st.cpu().numpy())
