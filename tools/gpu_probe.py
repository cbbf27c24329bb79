"""One-off GPU probe: FP64 peaks (our microbenchmarks + cuBLAS DGEMM/ZGEMM through torch) and a first throughput number."""
import json, time, sys
import numpy as np
import torch
sys.path.insert(0, '.')
from effective_model_learning_b200 import _native, engine
from oracle import lindblad_oracle as lo
from tests.helpers import model_from_spec, OBS3

out = {}
out['dfma_tflops'] = _native.measure_fp64_peak(0, False, 1.0) / 1e12
out['dmma_tflops'] = _native.measure_fp64_peak(0, True, 1.0) / 1e12
for dt, name in ((torch.float64, 'dgemm'), (torch.complex128, 'zgemm')):
    n = 4096
    a = torch.randn(n, n, dtype=dt, device='cuda'); b = torch.randn(n, n, dtype=dt, device='cuda')
    for _ in range(2): (a @ b)
    torch.cuda.synchronize()
    best = 0
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        fl = (2 if dt == torch.float64 else 8) * n ** 3
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3))
    out[name + '_tflops'] = best / 1e12
print(json.dumps(out))
ts = np.linspace(0, 2.5, 200)
for D, B in ((1, 2048), (2, 2048), (3, 2048), (4, 256)):
    Q = 1 if D < 4 else 2
    Dd = D if D < 4 else 3
    models = [model_from_spec(lo.random_library_spec(np.random.default_rng(1000 + b), Q, Dd)) for b in range(64)]
    ctx = engine.context_for(models[0], OBS3, ts)
    params = ctx.params_of([engine.model_processes(m) for m in models])
    params = np.tile(params, (B // 64, 1))
    ctx.set_targets(np.zeros((3, 200)))
    ctx.evaluate(params[:64], want_expect=False, want_loss=True)
    t = time.time(); ctx.evaluate(params, want_expect=False, want_loss=True); dt_ = time.time() - t
    print(f'n_tls={Q+Dd} B={B} kernel={ctx.plan.kernel_name} {B/dt_:.0f} evals/s (host API, wall)')
