"""GPU tests of the batched RJMCMC engine against the per-chain Python restatement (tests/chain_reference.py)."""
import numpy as np
import pytest
import torch

from effective_model_learning_b200 import BatchedChains, configs, definitions, engine
from oracle import lindblad_oracle as lo
from tests.chain_reference import ReferenceChain, MOVES
from tests.helpers import OBS3

pytestmark = pytest.mark.gpu


def _setup(D, n_chains, steps_cfg=None, temperature=0.002):
    cfg = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    cfg['shock_anneal_at'] = 40
    cfg['jump_length_rescaling_factor'] = 1.3
    cfg['temperature_proposal'] = temperature
    if steps_cfg:
        cfg['chain_step_options'] = steps_cfg
    ts = np.linspace(0, 2.5, 60)
    truth = lo.random_library_spec(np.random.default_rng(20260000 + 10 + D), 1, D)
    targets = np.array([np.asarray(e) for e in lo.expect_exact(truth, ts, OBS3)])
    targets = targets + np.random.default_rng(5).normal(0, 0.01, targets.shape)
    eng = BatchedChains(ts, targets, OBS3, n_chains=n_chains, n_defects=D, seed=1234, defect_initial_state=definitions.ops['mm'], **cfg)
    return cfg, ts, targets, eng


@pytest.mark.parametrize('D,temperature', [(2, 0.002), (3, 0.002), (1, (2.0, 0.001)), (2, (0.7, 0.004))])
def test_engine_matches_python_reference_step_by_step(D, temperature):
    """Fixed temperature and gamma-sampled temperature (shape >= 1 and the shape < 1 boost branch)."""
    n_chains, steps = 6, 90
    cfg, ts, targets, eng = _setup(D, n_chains, {m: (6 if m == MOVES[0] else 1) for m in MOVES}, temperature)
    state = eng.read()
    hyper = eng.hyperparameters
    refs = [ReferenceChain(eng.layout, hyper, cfg, state['current_params'][c], state['current_loss'][c], 1234, c)
            for c in range(n_chains)]
    hist = {'loss': torch.zeros((1, n_chains), dtype=torch.float64, device='cuda'),
            'aprob': torch.zeros((1, n_chains), dtype=torch.float64, device='cuda'),
            'code': torch.zeros((1, n_chains), dtype=torch.int32, device='cuda'),
            'params': torch.zeros((1, n_chains, eng.layout.n_params), dtype=torch.float64, device='cuda')}
    seen = set()
    for s in range(steps):
        props = np.stack([r.propose() for r in refs])
        _, ref_loss = eng.ctx.evaluate(props, want_expect=False, want_loss=True)
        eng.run(1, history=hist)
        state = eng.read()
        code = hist['code'].cpu().numpy()[0]
        assert np.array_equal(hist['params'].cpu().numpy()[0], state['current_params'])
        np.testing.assert_allclose(hist['loss'].cpu().numpy()[0], ref_loss, rtol=1e-9, atol=1e-14,
                                   err_msg=f'step {s}: the engine evaluated a different proposal')
        for c, r in enumerate(refs):
            a, accept = r.decide(float(ref_loss[c]))
            assert code[c] >> 1 == r.mv, (s, c)
            assert bool(code[c] & 1) == accept, (s, c, a)
            assert hist['aprob'].cpu().numpy()[0][c] == pytest.approx(a, rel=1e-7, abs=1e-300)
            np.testing.assert_allclose(state['current_params'][c], r.cur, rtol=1e-12, atol=1e-15)
            seen.add(r.mv)
        np.testing.assert_allclose(state['best_loss'], [r.best_loss for r in refs], rtol=1e-9)
    assert len(seen) >= 6, 'the run should exercise most move types'
    st = state['stats']
    for c, r in enumerate(refs):
        assert [st[c, 5], st[c, 4], st[c, 3], st[c, 2]] == r.counters
        assert st[c, 0] == pytest.approx(r.cur_loss, rel=1e-9) and st[c, 7] == r.ways(r.cur)[1]
    eng.close()


def test_many_chains_reduce_the_loss_and_stay_in_bounds():
    cfg, ts, targets, eng = _setup(2, 512)
    start = eng.read()['current_loss'].copy()
    eng.run(300)
    state = eng.read()
    assert np.median(state['best_loss']) < 0.25 * np.median(start)
    assert np.all(state['best_loss'] <= state['current_loss'] + 1e-15)
    lay = eng.layout
    b = cfg['params_bounds']
    for c, cls in enumerate(lay.classes):
        col = state['current_params'][:, c]
        if cls == 'qubit_energy':
            assert np.all(col == 1.0)
        elif cls == 'energies':
            assert np.all((col >= b['energies'][0]) & (col <= b['energies'][1]))
        else:
            nz = col[col != 0]
            assert np.all((nz >= b[cls][0]) & (nz <= b[cls][1]))
    # the reported losses are the losses of the reported models
    _, check = eng.ctx.evaluate(state['best_params'], want_expect=False, want_loss=True)
    np.testing.assert_allclose(check, state['best_loss'], rtol=1e-9, atol=1e-14)
    m = eng.best_models([0])[0]
    assert m.final_loss == state['best_loss'][0]
    table = eng.checkpoint()
    assert table.shape == (512, 8) and np.array_equal(table[:, 1], state['best_loss'])
    eng.close()


def test_results_do_not_depend_on_how_chains_are_split_over_engines():
    """first_chain makes chain c of a sweep use the same random stream whichever engine / GPU owns it."""
    cfg, ts, targets, whole = _setup(2, 64)
    whole.run(25)
    a = whole.read()
    whole.close()
    parts = []
    for lo_, hi_ in ((0, 24), (24, 64)):
        eng = BatchedChains(ts, targets, OBS3, n_chains=hi_ - lo_, n_defects=2, seed=1234, first_chain=lo_, defect_initial_state=definitions.ops['mm'], **cfg)
        eng.run(25)
        parts.append(eng.read())
        eng.close()
    for key in ('current_params', 'best_loss', 'stats'):
        assert np.array_equal(a[key], np.concatenate([p[key] for p in parts]))


def test_two_group_pipelining_gives_the_same_chains():
    """>= 2048 chains advance as two groups on two streams; the result must equal two single-group engines."""
    cfg, ts, targets, whole = _setup(1, 2304)
    whole.run(12)
    a = whole.read()
    whole.close()
    parts = []
    for lo_, hi_ in ((0, 1000), (1000, 2304)):
        eng = BatchedChains(ts, targets, OBS3, n_chains=hi_ - lo_, n_defects=1, seed=1234, first_chain=lo_, defect_initial_state=definitions.ops['mm'], **cfg)
        eng.run(12)
        parts.append(eng.read())
        eng.close()
    for key in ('current_params', 'current_loss', 'best_loss', 'stats'):
        assert np.array_equal(a[key], np.concatenate([p[key] for p in parts])), key


def test_checkpoint_resume_continues_exactly():
    cfg, ts, targets, a = _setup(2, 40)
    a.run(30)
    blob = a.export_state()
    a.run(30)
    want = a.read()
    a.close()
    b = BatchedChains(ts, targets, OBS3, n_chains=40, n_defects=2, seed=1234, defect_initial_state=definitions.ops['mm'], **cfg)
    b.import_state(blob)
    b.run(30)
    got = b.read()
    b.close()
    for key in want:
        assert np.array_equal(want[key], got[key]), key
    with pytest.raises(RuntimeError):
        c = BatchedChains(ts, targets, OBS3, n_chains=41, n_defects=2, seed=1234, defect_initial_state=definitions.ops['mm'], **cfg)
        c.import_state(blob)


def test_batched_chains_write_the_reference_output_files(tmp_path):
    """
    Scope row N3 from the batched engine: best models of device chains go through output.Output (the reference's file set,
    output.py:79-128) under the reference's module names, and come back as models whose parameter vector under the
    library (vectorise_under_library, basic_model.py:575-765) is the device's best row and whose dynamics reproduce its loss.
    """
    import json
    import pickle
    from effective_model_learning_b200 import compat, output
    cfg, ts, targets, eng = _setup(2, 32)
    eng.run(60)
    state = eng.read()

    class Toggles:
        pickle = True
        text = True
        hyperparams = True

    which = [0, 7, 31]
    for c, best in zip(which, eng.best_models(which)):
        st = state['stats'][c]
        output.Output(toggles=Toggles, filename=str(tmp_path / f'sweep_D2_R{c + 1}'), models_to_save=[best], model_names=['best'],
                      overall_acceptance={'RJ': st[4] / max(st[5], 1.0), 'tweak': st[2] / max(st[3], 1.0)},
                      chain_hyperparams=cfg, chain_name='chain', reference_module_names=True)
    compat.install_reference_aliases()
    try:
        for c in which:
            base = str(tmp_path / f'sweep_D2_R{c + 1}')
            with open(base + '_best.pickle', 'rb') as f:
                m = pickle.load(f)
            assert type(m).__module__.endswith('learning_model') and m.final_loss == state['best_loss'][c]
            vals, labels, _ = m.vectorise_under_library(cfg)
            want = [v for v, cls in zip(state['best_params'][c], eng.layout.classes) if cls != 'qubit_energy']
            np.testing.assert_allclose(vals, want, rtol=0, atol=0)
            loss = engine.evaluate_models([m], ts, OBS3, targets=targets)[1][0]
            assert loss == pytest.approx(state['best_loss'][c], rel=1e-9)
            acc = json.load(open(base + '_overall_acceptance.json'))
            assert 0.0 <= acc['tweak'] <= 1.0
            assert pickle.load(open(base + '_chain_hyperparameters.pickle', 'rb'))['complexity_factor'] == cfg['complexity_factor']
            assert 'TLS' in open(base + '_best.txt').read() or len(open(base + '_best.txt').read()) > 20
    finally:
        import sys
        for name in compat.FLAT_MODULES:
            if getattr(sys.modules.get(name), '__name__', '').startswith('effective_model_learning_b200.'):
                sys.modules.pop(name, None)
    eng.close()
