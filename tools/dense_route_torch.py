"""
The north star's DENSE route at library speed, for comparison with the kernels of this repository:
column-stacked d^2 x d^2 Liouvillian -> expm(L dt) (torch.linalg.matrix_exp = cuBLAS ZGEMM scaling-and-squaring Pade)
-> T-1 mat-vecs (cuBLAS batched GEMV) -> traces.  Only the GPU part (expm + stepping + traces) is timed; the
Liouvillians are assembled on the host with the package's own model classes.  Also a GPU-side cross-check: the two
routes must agree to 1e-8.

    python tools/dense_route_torch.py [n_defects] [batch]
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from effective_model_learning_b200 import configs, synthetic  # noqa: E402

OBS3 = ['sigmax', 'sigmay', 'sigmaz']
D = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
lay = synthetic.LibraryLayout(1, D, hyper)
ts = np.linspace(0, 2.5, 200)
params = lay.batch_params(1000 + np.arange(B))

# ---- this repository's route
ctx = lay.context(ts, OBS3)
ours, _ = ctx.evaluate(params, want_expect=True)
ctx.set_targets(np.zeros((3, len(ts))))
pd = torch.from_numpy(params).cuda()
loss = torch.zeros(B, dtype=torch.float64, device='cuda')
st = torch.zeros(B, dtype=torch.int32, device='cuda')
stream = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    ctx.plan.eval_device(B, pd.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), stream)
torch.cuda.synchronize()
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record()
for _ in range(5):
    ctx.plan.eval_device(B, pd.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), stream)
a1.record()
torch.cuda.synchronize()
t_ours = a0.elapsed_time(a1) * 1e-3 / 5

# ---- dense route: host assembly (not timed), GPU expm + stepping (timed with CUDA events)
d = 2 ** (1 + D)
Ls = np.empty((B, d * d, d * d), dtype=np.complex128)
for b in range(B):
    Ls[b] = lay.model_from_params(params[b]).build_Liouvillian()
m0 = lay.model_from_params(params[0])
rho0 = m0.initial_DM.ravel(order='F')
W = np.stack([np.kron(np.eye(d // 2), np.zeros((2, 2)))] * 3).astype(np.complex128) * 0
from effective_model_learning_b200 import definitions  # noqa: E402
for m, name in enumerate(OBS3):
    O = np.kron(definitions.ops[name], np.eye(d // 2))
    W[m] = 0
    W_row = O.T.ravel(order='F')
    if m == 0:
        Wrows = np.zeros((3, d * d), dtype=np.complex128)
    Wrows[m] = W_row
Lg = torch.from_numpy(Ls).cuda()
v0 = torch.from_numpy(np.tile(rho0, (B, 1))).cuda().unsqueeze(-1)
Wg = torch.from_numpy(Wrows).cuda()
dt = float(ts[1] - ts[0])
for rep in range(2):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    P = torch.linalg.matrix_exp(Lg * dt)
    e1.record()
    v = v0.clone()
    out = torch.empty((B, 3, len(ts)), dtype=torch.complex128, device='cuda')
    for k in range(len(ts)):
        if k:
            v = torch.bmm(P, v)
        out[:, :, k] = torch.matmul(Wg, v.squeeze(-1).T).T
    e2.record()
    torch.cuda.synchronize()
dense = out.cpu().numpy()
ms_expm, ms_step = e0.elapsed_time(e1), e1.elapsed_time(e2)
print(json.dumps({'d': d, 'batch': B, 'max_abs_diff_between_routes': float(np.abs(dense - ours).max()),
                  'dense_route_expm_ms': ms_expm, 'dense_route_stepping_ms': ms_step,
                  'dense_route_evals_per_s': B / ((ms_expm + ms_step) * 1e-3),
                  'this_repo_evals_per_s_same_batch': B / t_ours}))
