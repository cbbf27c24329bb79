"""
GPU tests of the launch structure of the d <= 16 Chebyshev path (csrc/eml_warp_mma.cu: launch_warp_mma): the longest-first
order from the prepare kernel's cost-bin lists, the fetch workspace that the evaluator's last warp zeroes for the next
launch, and the evaluator as a programmatic dependent launch.  None of them may change a single bit of a result, so the same
batches are evaluated (a) repeatedly and in alternating sizes through one plan and (b) in child processes with the
switches off (EML_ORDER_LISTS=0 EML_PDL=0: ordering kernel, ordinary launch) and with two d = 4 models per warp forced on
for launches that use the bin lists (EML_PACK2=2) -- the environment is read once per process.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TS = np.linspace(0.0, 2.5, 200)
OBS3 = ['sigmax', 'sigmay', 'sigmaz']
CASES = [(3, 4096), (3, 3000), (1, 3001), (2, 2500)]      # (defects, models): all above the warp slots of a B200, at most 4096

CHILD = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tests.test_gpu_launch_switches import evaluate_cases
np.savez(sys.argv[2], **evaluate_cases())
'''


def _layout(D):
    from effective_model_learning_b200 import configs, synthetic
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    return synthetic.LibraryLayout(1, D, hyper)


def evaluate_cases():
    out = {}
    targets = np.random.default_rng(3).normal(0, 0.3, (3, len(TS)))
    for D, B in CASES:
        lay = _layout(D)
        ctx = lay.context(TS, OBS3)
        ctx.set_targets(targets)
        _, loss = ctx.evaluate(lay.batch_params(1000 + np.arange(B)), want_expect=False, want_loss=True)
        out[f'loss_{D}_{B}'] = loss
    return out


def _child(tmp_path, name, env):
    path = str(tmp_path / f'{name}.npz')
    full = dict(os.environ, **env)
    subprocess.run([sys.executable, '-c', CHILD, ROOT, path], check=True, env=full, cwd=ROOT, timeout=600)
    return np.load(path)


def test_results_do_not_depend_on_the_launch_switches(tmp_path):
    here = evaluate_cases()
    plain = _child(tmp_path, 'plain', {'EML_ORDER_LISTS': '0', 'EML_PDL': '0'})
    for key, loss in here.items():
        assert np.all(np.isfinite(loss))
        assert np.array_equal(plain[key], loss), key      # one model per warp either way: the same bits
    packed = _child(tmp_path, 'pack2', {'EML_PACK2': '2'})
    for key, loss in here.items():
        if key.startswith('loss_1_'):      # d = 4: pairs share a recurrence (the longer series of the two), last bits differ
            np.testing.assert_allclose(packed[key], loss, rtol=1e-9, atol=1e-15)
            assert not np.array_equal(packed[key], loss)      # (the pair kernel did run)
        else:
            assert np.array_equal(packed[key], loss), key


def test_fetch_workspace_is_clean_after_every_launch():
    """Alternating batch sizes through ONE plan (bin lists <-> no order <-> ordering kernel), with rows that cannot be
    evaluated in between: every launch must find the counter and the bin counts zero, i.e. return what the first one did."""
    import torch
    lay = _layout(3)
    ctx = lay.context(TS, OBS3)
    ctx.set_targets(np.zeros((3, len(TS))))
    params = lay.batch_params(1000 + np.arange(6000))
    _, ref = ctx.evaluate(params, want_expect=False, want_loss=True)      # 6000 models: ordering kernel
    for B in (4096, 100, 3000, 4096, 6000, 2369, 1, 4096):
        _, loss = ctx.evaluate(params[:B], want_expect=False, want_loss=True)
        assert np.array_equal(loss, ref[:B]), B
    bad = params[:4096].copy()
    bad[[0, 17, 4095]] = np.nan
    p = torch.from_numpy(bad).cuda()
    loss = torch.zeros(4096, dtype=torch.float64, device='cuda')
    st = torch.zeros(4096, dtype=torch.int32, device='cuda')
    for _ in range(2):
        ctx.plan.eval_device(4096, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got, status = loss.cpu().numpy(), st.cpu().numpy()
        ok = np.ones(4096, bool)
        ok[[0, 17, 4095]] = False
        assert np.all(status[~ok] == -1) and np.all(status[ok] > 0)
        assert np.array_equal(got[ok], ref[:4096][ok])
    _, loss = ctx.evaluate(params[:4096], want_expect=False, want_loss=True)
    assert np.array_equal(loss, ref[:4096])
