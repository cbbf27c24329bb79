"""
A/B of the model order of the d = 16 Chebyshev kernel in ONE process (eml_set_order_pack_mode): the packed list must
evaluate every model exactly once (result buffers are poisoned before each run and compared bit for bit with the
longest-first run), then both are timed with CUDA events.  Usage: python tools/order_pack_ab.py [chains] [reps]
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from effective_model_learning_b200 import _native, configs, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
lay = synthetic.LibraryLayout(1, 3, hyper)
ctx = lay.context(np.linspace(0, 2.5, 200), ['sigmax', 'sigmay', 'sigmaz'])
ctx.set_targets(np.random.default_rng(5).normal(0, 0.3, (3, 200)))
plan = ctx.plan
p = torch.from_numpy(lay.batch_params(1000 + np.arange(B))).cuda()
loss = torch.empty(B, dtype=torch.float64, device='cuda')
st = torch.empty(B, dtype=torch.int32, device='cuda')
s = torch.cuda.current_stream().cuda_stream
out = {'chains': B, 'kernel': plan.kernel_name}
ref = None
for mode in (0, 1):
    _native.set_order_pack_mode(mode)
    loss.fill_(float('nan'))
    st.fill_(-7)
    plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
    got = (loss.cpu().numpy().copy(), st.cpu().numpy().copy())
    assert np.isfinite(got[0]).all() and (got[1] > 0).all(), f'mode {mode}: models left unevaluated'
    if ref is None:
        ref = got
    else:
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), 'packed order changed results'
    for _ in range(3):
        plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    out[f'ms_per_batch_mode{mode}'] = e0.elapsed_time(e1) / reps
_native.set_order_pack_mode(-1)
out['identical_results'] = True
# the pre-pass's predicted application counts against the evaluator's own (status)
pred = torch.empty(B, dtype=torch.float32, device='cuda')
plan.predict_applications_device(B, p.data_ptr(), pred.data_ptr(), s)
torch.cuda.synchronize()
dev = pred.cpu().numpy() - ref[1]
out['predicted_applications'] = {'mean': float(pred.mean()), 'actual_mean': float(ref[1].mean()), 'max_abs_dev': float(np.abs(dev).max()),
                                 'models_off_by_more_than_2': int((np.abs(dev) > 2).sum())}
print(json.dumps(out))
