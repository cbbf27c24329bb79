"""
GPU tests at the FULL sizes of BASELINE.json's configurations (C2: 1024 x d=8, C3: 4096 x d=16, C4: 512 x d=32; 200 times,
3 observables), through properties that do not need the CPU oracle at that size: trace preservation, positivity of the
reduced state, consistency of the fused loss with the returned observables, independence of the result from the
position of a model in the batch (longest-first ordering, persistent warps) and from the time grid it is embedded in
(macro-step selection), plus an oracle spot check of a few models of the same batch.
"""
import numpy as np
import pytest

from oracle import lindblad_oracle as lo
from tests.helpers import spec_from_context_row

pytestmark = pytest.mark.gpu
TS = np.linspace(0.0, 2.5, 200)


def _layout(Q, D):
    from effective_model_learning_b200 import configs, synthetic
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    return synthetic.LibraryLayout(Q, D, hyper)


@pytest.mark.parametrize('Q,D,B', [(1, 2, 1024), (1, 3, 4096), (2, 3, 512)])
def test_full_size_batch_properties(Q, D, B):
    lay = _layout(Q, D)
    params = lay.batch_params(1000 + np.arange(B))
    rng = np.random.default_rng(5)

    # --- trace preservation and positivity: <exc> + <gnd> = Tr rho = 1, <exc> - <gnd> = <sigmaz>
    ctx = lay.context(TS, ['exc', 'gnd', 'sigmaz'])
    ex, _ = ctx.evaluate(params, want_expect=True)
    assert ex.shape == (B, 3, len(TS))
    assert np.abs(ex.imag).max() < 1e-13
    ex = ex.real
    assert np.abs(ex[:, 0] + ex[:, 1] - 1.0).max() < 1e-12
    assert np.abs(ex[:, 0] - ex[:, 1] - ex[:, 2]).max() < 1e-12
    assert ex[:, :2].min() > -1e-12 and ex[:, :2].max() < 1 + 1e-12

    # --- Bloch vector inside the unit ball; fused loss == loss recomputed from the returned observables
    ctx = lay.context(TS, ['sigmax', 'sigmay', 'sigmaz'])
    targets = rng.normal(0, 0.3, (3, len(TS)))
    ctx.set_targets(targets)
    ex, loss = ctx.evaluate(params, want_expect=True, want_loss=True)
    ex = ex.real
    assert (ex ** 2).sum(axis=1).max() < 1 + 1e-12
    assert np.abs(ex[:, :, 0] - np.array([1.0, 0.0, 0.0])).max() < 1e-14          # rho0: qubit |+><+|
    want = ((ex - targets[None]) ** 2).sum(axis=(1, 2)) / (3 * len(TS))
    np.testing.assert_allclose(loss, want, rtol=1e-12, atol=1e-15)

    # --- a model's result does not depend on where it sits in the batch, nor on the batch it is in
    perm = rng.permutation(B)
    ex_p, loss_p = ctx.evaluate(params[perm], want_expect=True, want_loss=True)
    assert np.array_equal(ex_p.real, ex[perm]) and np.array_equal(loss_p, loss[perm])
    sub = rng.choice(B, 37, replace=False)
    ex_s, loss_s = ctx.evaluate(params[sub], want_expect=True, want_loss=True)
    assert np.array_equal(ex_s.real, ex[sub]) and np.array_equal(loss_s, loss[sub])

    # --- nor on the time grid around the common points: every other time, and the grid extended by a far point
    ctx2 = lay.context(TS[::2], ['sigmax', 'sigmay', 'sigmaz'])
    ex2, _ = ctx2.evaluate(params[:256], want_expect=True)
    assert np.abs(ex2.real - ex[:256, :, ::2]).max() < 1e-12
    ctx3 = lay.context(np.concatenate([TS, [9.0]]), ['sigmax', 'sigmay', 'sigmaz'])
    ex3, _ = ctx3.evaluate(params[:256], want_expect=True)
    assert np.abs(ex3.real[:, :, :-1] - ex[:256]).max() < 1e-12

    # --- oracle spot check on models of this very batch
    for b in rng.choice(B, 3 if D == 3 and Q == 2 else 5, replace=False):
        spec = spec_from_context_row(ctx, params[b])
        want_b = np.array([np.asarray(e) for e in lo.expect_exact(spec, TS, ['sigmax', 'sigmay', 'sigmaz'])])
        assert np.abs(ex[b] - want_b.real).max() < 1e-11, b


def test_page_locked_caller_buffers_are_used_directly_and_give_the_same_results():
    """eml_eval_batch_host DMAs straight from / into page-locked caller arrays (no staging copy): same bits as pageable."""
    import torch
    lay = _layout(1, 2)
    params = lay.batch_params(2000 + np.arange(300))
    ctx = lay.context(TS, ['sigmax', 'sigmay', 'sigmaz'])
    ctx.set_targets(np.zeros((3, len(TS))))
    ex, loss = ctx.evaluate(params, want_expect=True, want_loss=True)
    p_pin = torch.from_numpy(params).pin_memory().numpy()
    loss_pin = torch.empty(300, dtype=torch.float64).pin_memory().numpy()
    ex_pin = torch.empty((300, 3, len(TS)), dtype=torch.complex128).pin_memory().numpy()
    ex2, loss2 = ctx.evaluate(p_pin, want_expect=True, want_loss=True, loss_out=loss_pin, expect_out=ex_pin)
    assert ex2 is ex_pin and loss2 is loss_pin
    assert np.array_equal(ex2, ex) and np.array_equal(loss2, loss)
    with pytest.raises(RuntimeError, match='loss_out'):
        ctx.evaluate(params, want_expect=False, want_loss=True, loss_out=np.empty(7))
