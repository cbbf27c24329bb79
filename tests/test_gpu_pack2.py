"""
GPU tests of the two-models-per-warp evaluator for d = 4 and d = 2 (eval_pack2_kernel): it runs when a launch has more than
2.5 models per warp slot (about 5 900 on a B200), so the batches here are large; a model's result must agree with what the
one-model-per-warp kernel returns for it in a small batch, with the exact oracle, and must not depend on its partner
(not even on a partner that cannot be evaluated).
"""
import numpy as np
import pytest

from oracle import lindblad_oracle as lo
from tests.helpers import spec_from_context_row

pytestmark = pytest.mark.gpu
OBS3 = ['sigmax', 'sigmay', 'sigmaz']


def _layout(Q, D):
    from effective_model_learning_b200 import configs, synthetic
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    return synthetic.LibraryLayout(Q, D, hyper)


@pytest.mark.parametrize('D,B,span', [(1, 6501, 2.5), (0, 6200, 2.5), (1, 6500, 70.0)])
def test_packed_pairs_equal_single_models(D, B, span):
    """span 70: several macro steps (the pair follows the shorter step, the coefficient tables are per model)."""
    lay = _layout(1, D)
    ts = np.linspace(0.0, span, 200)
    params = lay.batch_params(1000 + np.arange(B))
    rng = np.random.default_rng(11)
    targets = rng.normal(0, 0.3, (3, len(ts)))
    ctx = lay.context(ts, OBS3)
    ctx.set_targets(targets)
    ex, loss = ctx.evaluate(params, want_expect=True, want_loss=True)
    assert np.abs(ex.imag).max() < 1e-13
    ex = ex.real
    assert (ex ** 2).sum(axis=1).max() < 1 + 1e-11
    want = ((ex - targets[None]) ** 2).sum(axis=(1, 2)) / (3 * len(ts))
    np.testing.assert_allclose(loss, want, rtol=1e-11, atol=1e-15)
    # the same models in a small batch run one per warp
    sub = np.concatenate([[0, 1, B - 2, B - 1], rng.choice(B, 60, replace=False)])
    ex_s, loss_s = ctx.evaluate(params[sub], want_expect=True, want_loss=True)
    tol = 1e-12 if span < 10 else 1e-10
    assert np.abs(ex_s.real - ex[sub]).max() < tol
    np.testing.assert_allclose(loss_s, loss[sub], rtol=1e-9, atol=1e-15)
    # exact oracle, a few models (both slots of a pair among them)
    for b in (0, 1, int(sub[7])):
        spec = spec_from_context_row(ctx, params[b])
        want_b = np.array([np.asarray(e) for e in lo.expect_exact(spec, ts, OBS3)])
        assert np.abs(ex[b] - want_b.real).max() < 1e-8, b


def test_a_partner_that_cannot_be_evaluated_does_not_disturb_the_pair():
    import torch
    lay = _layout(1, 1)
    ts = np.linspace(0.0, 2.5, 200)
    B = 6400
    params = lay.batch_params(1000 + np.arange(B))
    ctx = lay.context(ts, OBS3)
    ctx.set_targets(np.zeros((3, len(ts))))
    _, good = ctx.evaluate(params, want_loss=True)
    bad = params.copy()
    bad[[5, 6, 4001]] = np.nan
    p = torch.from_numpy(bad).cuda()
    loss = torch.zeros(B, dtype=torch.float64, device='cuda')
    st = torch.zeros(B, dtype=torch.int32, device='cuda')
    ctx.plan.eval_device(B, p.data_ptr(), 0, 0, loss.data_ptr(), st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    loss, st = loss.cpu().numpy(), st.cpu().numpy()
    ok = np.ones(B, bool)
    ok[[5, 6, 4001]] = False
    assert np.all(st[~ok] == -1) and np.all(np.isnan(loss[~ok]))
    assert np.all(st[ok] > 0)
    np.testing.assert_allclose(loss[ok], good[ok], rtol=1e-11, atol=1e-16)
