"""
Where the warps of the d <= 16 Chebyshev evaluator spend their cycles (profiling build only):
    tools/build_variant.sh prof -DEML_PROFILE_PHASES
    EML_B200_LIB=build/variants/libeml_prof.so python tools/phase_profile.py [--defects 3] [--chains 4096 65536]
Prints warp-cycles per phase (clock64 deltas summed over warps) and the idle share of the warp-slot time
(slots x launch duration - busy cycles).
"""
import argparse
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from effective_model_learning_b200 import _native, configs, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--defects', type=int, default=3)
ap.add_argument('--chains', type=int, nargs='+', default=[4096, 65536])
ap.add_argument('--slots-per-sm', type=int, default=16)
ap.add_argument('--flags', type=int, default=0)
a = ap.parse_args()
OBS3 = ['sigmax', 'sigmay', 'sigmaz']
hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
ts = np.linspace(0, 2.5, 200)
lib = _native.load()
lib.eml_debug_phase_clocks.argtypes = [C.POINTER(C.c_uint64), C.c_int]
props = torch.cuda.get_device_properties(0)
sm_hz = props.clock_rate * 1e3 if hasattr(props, 'clock_rate') else 1.965e9
for B in a.chains:
    lay = synthetic.LibraryLayout(1, a.defects, hyper)
    ctx = lay.context(ts, OBS3, flags=a.flags)
    ctx.set_targets(np.zeros((3, 200)))
    plan = ctx.plan
    p = torch.from_numpy(lay.batch_params(1000 + np.arange(B))).cuda()
    loss = torch.zeros(B, dtype=torch.float64, device='cuda')
    st = torch.zeros(B, dtype=torch.int32, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    torch.cuda.synchronize()
    out = (C.c_uint64 * 8)()
    _native.check(lib.eml_debug_phase_clocks(out, 1), 'phase clocks')
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    _native.check(lib.eml_debug_phase_clocks(out, 1), 'phase clocks')
    clk = np.array(list(out)[:5], dtype=np.float64) / reps
    slots = props.multi_processor_count * a.slots_per_sm
    total = slots * ms * 1e-3 * sm_hz
    napp = float(st.float().mean())
    print(json.dumps({'kernel': plan.kernel_name, 'models': B, 'ms': ms, 'evals_per_s': B / ms * 1e3,
                      'applications_per_eval': napp,
                      'busy_share_of_slot_time': clk.sum() / total,
                      'phase_share_of_busy': dict(zip(('assembly+bounds', 'plan+first', 'recurrence', 'outputs', 'epilogue'),
                                                      (clk / clk.sum()).round(4).tolist())),
                      'warp_clk_per_application': clk[2] / (B * napp),
                      'warp_clk_per_model': dict(zip(('assembly+bounds', 'plan+first', 'recurrence', 'outputs', 'epilogue'),
                                                     (clk / B).round(0).tolist())), 'sm_hz_assumed': sm_hz}))
