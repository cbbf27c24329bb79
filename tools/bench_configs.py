"""
Evaluation throughput of the other BASELINE.json configurations (parity-test cases, not the bench line):
device-resident parameters, CUDA-event timed, through eml_eval_batch.  Usage: python tools/bench_configs.py
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from effective_model_learning_b200 import _native, configs, synthetic  # noqa: E402

OBS3 = ['sigmax', 'sigmay', 'sigmaz']
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
ts = np.linspace(0, 2.5, 200)
rows = []
for name, Q, D, B, reps in (('C1 d=4, 1 chain', 1, 1, 1, 200), ('d=4, 4096 chains', 1, 1, 4096, 20),
                            ('d=4, 8192 chains (two models per warp)', 1, 1, 8192, 20), ('d=4, 65536 chains (two models per warp)', 1, 1, 65536, 5),
                            ('C2 d=8, 1024 chains', 1, 2, 1024, 50), ('d=8, 8192 chains', 1, 2, 8192, 20),
                            ('C3 d=16, 4096 chains', 1, 3, 4096, 20), ('d=16, 16384 chains', 1, 3, 16384, 10), ('d=16, 65536 chains', 1, 3, 65536, 3),
                            ('C4 d=32, 512 chains', 2, 3, 512, 3), ('d=32, 2048 chains', 2, 3, 2048, 2)):
    lay = synthetic.LibraryLayout(Q, D, hyper)
    ctx = lay.context(ts, OBS3)
    ctx.set_targets(np.zeros((3, 200)))
    plan = ctx.plan
    p = torch.from_numpy(lay.batch_params(1000 + np.arange(B))).cuda()
    loss = torch.zeros(B, dtype=torch.float64, device='cuda')
    st = torch.zeros(B, dtype=torch.int32, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rows.append({'config': name, 'kernel': plan.kernel_name, 'models': B, 'ms_per_batch': ms, 'evals_per_s': B / ms * 1e3,
                 'applications_per_eval': float(st.float().mean())})
    print(json.dumps(rows[-1]))
