"""CPU tests of the host-side mirror of the reference API (no GPU: the device is replaced by the oracle)."""
import copy
import json
import os
import re

import numpy as np
import pytest
import scipy.linalg

import effective_model_learning_b200 as eml
from effective_model_learning_b200 import _native, configs, definitions, engine
from oracle import lindblad_oracle as lo
from tests.helpers import OBS3, model_from_spec
from tests.test_oracle import spec_from_json

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
ROOT = os.path.dirname(os.path.dirname(__file__))


def _jsonable(x):
    if isinstance(x, dict):
        return {(k if isinstance(k, str) else repr(k)): _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    return x


def test_configs_equal_reference_dump():
    ref = json.load(open(os.path.join(GOLD, 'reference_configs.json')))
    assert _jsonable(configs.default_chain_hyperparams) == ref['default_chain_hyperparams']
    assert list(configs.specific_experiment_chain_hyperparams) == ref['order']
    assert _jsonable(configs.specific_experiment_chain_hyperparams) == ref['specific_experiment_chain_hyperparams']
    h = configs.get_hyperparams(ref['order'][0])
    assert list(h['qubit_Ls_library']) == ['sigmaz'] and h['acceptance_window'] == 10


def test_ops_table_is_the_oracles():
    assert list(definitions.ops) == list(lo.OPS)
    for k in lo.OPS:
        assert np.array_equal(definitions.ops[k], lo.OPS[k]), k
    assert definitions.T(1, definitions.ops['id2']) is definitions.ops['id2']
    assert definitions.T(definitions.ops['sigmax'], definitions.ops['sigmaz']).shape == (4, 4)


def test_model_operators_match_reference_golden():
    z = np.load(os.path.join(GOLD, 'reference_operators.npz'))
    meta = json.load(open(os.path.join(GOLD, 'reference_operators_specs.json')))
    for i, js in enumerate(meta['specs']):
        m = model_from_spec(spec_from_json(js))
        assert np.array_equal(m.H, z[f'H_{i}']), i
        assert len(m.Ls) == z[f'Ls_{i}'].shape[0]
        for a, b in zip(m.Ls, z[f'Ls_{i}']):
            assert np.array_equal(a, b)
        assert np.array_equal(m.initial_DM, z[f'rho0_{i}'])


def test_calculate_dynamics_contract(oracle_backend):
    meta = json.load(open(os.path.join(GOLD, 'reference_operators_specs.json')))
    z = np.load(os.path.join(GOLD, 'reference_operators.npz'))
    ts = np.array(meta['times'])
    for i in (0, 3, 7, 13, 17, 20, 23):
        m = model_from_spec(spec_from_json(meta['specs'][i]))
        got = m.calculate_dynamics(ts, OBS3)
        want = z[f'expect_{i}']
        assert len(got) == 3
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, atol=1e-12)
    # default observable, string wrapping, custom function, legacy method names, errors
    m = model_from_spec(lo.random_library_spec(np.random.default_rng(0), 1, 1))
    a = m.calculate_dynamics(ts)
    b = m.calculate_dynamics(ts, 'exc', dynamics_method='qutip')
    assert len(a) == 1 and np.array_equal(a[0], b[0]) and a[0].dtype == np.float64
    assert m.calculate_dynamics(ts, ['sigmap'])[0].dtype == np.complex128
    assert m.calculate_dynamics(ts, 'sigmaz', custom_function_on_return=lambda r: 'wrapped') == 'wrapped'
    with pytest.raises(RuntimeError):
        m.calculate_dynamics(ts, [1])
    with pytest.raises(RuntimeError):
        eml.BasicModel().calculate_dynamics(ts)


def test_no_cpu_fallback_without_device():
    if _native.device_count() > 0:
        pytest.skip('a GPU is present')
    m = model_from_spec(lo.random_library_spec(np.random.default_rng(0), 1, 1))
    engine.clear_contexts()
    with pytest.raises(RuntimeError, match='no CUDA device|eml_plan_create'):
        m.calculate_dynamics(np.linspace(0, 1, 5))


def test_c_abi_exports_every_declared_symbol():
    lib = _native.load()
    header = open(os.path.join(ROOT, 'include', 'eml_b200.h')).read()
    declared = set(re.findall(r'\b(eml_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.eml_version() == 1


def test_add_TLS_input_normalisation():
    m = eml.BasicModel()
    q = m.add_TLS('qubit', True, 1.0)
    d = m.add_TLS('d', False, 0.2, couplings={'qubit': (0.5, ('sigmax', 'sigmax'))})
    assert d.couplings == {q: [(0.5, [('sigmax', 'sigmax')])]}
    with pytest.raises(RuntimeError):
        m.add_TLS('bad', False, 0.2, couplings={'qubit': [(0.5, 'sigmax', 'sigmax')]})
    assert 'Qubit' in m.model_description_str()
    assert m.model_description_dict()['1']['couplings:'] == {'0': [(0.5, [('sigmax', 'sigmax')])]}


def test_vectorise_matches_reference_and_round_trips():
    meta = json.load(open(os.path.join(GOLD, 'reference_operators_specs.json')))
    vec = json.load(open(os.path.join(GOLD, 'reference_vectorise.json')))
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    for i in range(12):
        spec = spec_from_json(meta['specs'][i])
        m = model_from_spec(spec, eml.LearningModel)
        vals, labels, latex = m.vectorise_under_library(hyper)
        assert labels == vec[i]['labels']
        assert [float(v) for v in vals] == vec[i]['values']
        D = len(spec['tls']) - 1
        back = eml.LearningModel().configure_to_params_vector((vals, labels, latex), D=D, default_qubit_energy=1.0,
                                                              qubit_initial_state=definitions.ops['plus'],
                                                              defect_initial_state=definitions.ops['mm'])
        assert back.vectorise_under_library(hyper)[0] == vals
        np.testing.assert_allclose(back.H, m.H, atol=1e-15)


def test_marshalling_canonical_processes():
    m = eml.BasicModel()
    q = m.add_TLS('qubit', True, 1.0, Ls={'sigmaz': 0.1})
    d = m.add_TLS('d', False, 0.2, couplings={'qubit': [(0.5, [('sigmax', 'sigmay')]), (0.7, [('sigmax', 'sigmay')])]})
    m.build_operators()
    procs = dict(engine.model_processes(m))
    # held by the defect with (op_defect, op_qubit): canonical form puts the qubit (site 0) first and swaps the pair
    assert procs[('C', 0, 1, (('sigmay', 'sigmax'),), 0)] == 0.5
    assert procs[('C', 0, 1, (('sigmay', 'sigmax'),), 1)] == 0.7     # repeated type keeps its own column (s^2 does not add)
    assert procs[('L', 0, 'sigmaz')] == 0.1 and procs[('E', 1)] == 0.2


def test_deepcopy_is_independent():
    m = model_from_spec(lo.random_library_spec(np.random.default_rng(3), 1, 2), eml.LearningModel)
    c = copy.deepcopy(m)
    assert c.vectorise_under_library(configs.default_chain_hyperparams)[0] == m.vectorise_under_library(configs.default_chain_hyperparams)[0]
    for t, u in zip(m.TLSs, c.TLSs):
        assert t is not u and u.model is c
        for p in u.couplings:
            assert p in c.TLSs
    c.TLSs[1].Ls['sigmax'] = 0.123
    c.TLSs[1].energy = 9.0
    assert m.TLSs[1].Ls.get('sigmax') != 0.123 and m.TLSs[1].energy != 9.0


def test_one_pass_move_counts_and_vectorised_normalisers_equal_the_per_move_path(oracle_backend):
    """
    The chain's hot loop counts all twelve jump moves in one pass (ProcessHandler.count_moves) and evaluates the
    truncated-Gaussian normalisers in one vectorised call: both must give what the reference-shaped methods give one
    move / one parameter at a time (process_handling.py update=False counts; learning_chain.py:979-994 xi), bit for bit.
    """
    from effective_model_learning_b200.learning_chain import _JUMP_METHODS
    cfg = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    cfg['iterations_till_progress_update'] = False
    ts = np.linspace(0, 1, 5)
    np.random.seed(7)
    chain = eml.LearningChain(target_times=ts, target_datasets=[np.zeros(5)] * 3, target_observables=OBS3, initial=(1, 3),
                              qubit_initial_state=definitions.ops['plus'], defect_initial_state=definitions.ops['mm'], **cfg)
    # the helpers create the handlers on first use, as step() does (tests/test_gpu_chain_statistics.py calls them on a chain
    # that never ran)
    model = copy.deepcopy(chain.current)
    assert not chain.params_handler
    assert chain._xi_product(model) == chain.xi(0.1, chain.params_handler.jump_lengths['energies'], *chain.params_bounds['energies']) ** 3
    chain.params_handler = False
    assert len(chain._move_weights(model)) == len(chain.next_step_labels) and chain.params_handler
    moves = [m for m in chain.next_step_labels if m != 'tweak all parameters']
    seen_nonzero = set()
    for it in range(300):
        ways = chain.process_handler.count_moves(model)
        for label in moves:
            assert ways[_JUMP_METHODS[label]] == chain.step(model, label, update=False)[1], (it, label)
            if ways[_JUMP_METHODS[label]]:
                seen_nonzero.add(label)
        assert chain._move_weights(model) == [chain.next_step_priorities_dict[x] * chain.step(model, x, update=False)[1]
                                              for x in chain.next_step_labels]
        # the per-parameter product, as the reference writes it
        jl, bounds = chain.params_handler.jump_lengths, chain.params_bounds
        want = 1
        for tls in model.TLSs:
            if not tls.is_qubit:
                want *= chain.xi(tls.energy, jl['energies'], *bounds['energies'])
            for rate in tls.Ls.values():
                want *= chain.xi(rate, jl['Ls'], *bounds['Ls'])
            for partner in tls.couplings:
                for coupling in tls.couplings[partner]:
                    want *= chain.xi(coupling[0], jl['couplings'], *bounds['couplings'])
        assert chain._xi_product(model) == want
        prior_want = np.exp(-(chain.process_handler.count_Ls(model) + chain.process_handler.count_defect2defect_couplings(model)
                              + chain.process_handler.count_qubit2defect_couplings(model)) / chain.complexity_factor)
        assert chain.prior(model) == prior_want
        # random walk through model space: a random available jump move, now and then a tweak
        label = moves[np.random.choice(len(moves))]
        model, _ = chain.step(model, label, update=True)
        if it % 5 == 0:
            model, _ = chain.step(model, 'tweak all parameters', update=True)
    assert seen_nonzero == set(moves)
    # a missing library: the per-move path (and its error) is kept
    chain.process_handler.qubit_Ls_library = False
    assert chain.process_handler.count_moves(model) is None
    with pytest.raises(RuntimeError, match='library not specified'):
        chain._move_weights(model)


@pytest.mark.parametrize('bounded', [True, False])
def test_run_equals_the_plain_loop_bit_for_bit(oracle_backend, bounded):
    """
    LearningChain.run against tests/plain_chain_loop.PlainLoopChain (everything recomputed every step, np.random.choice)
    on the same seed: acceptance windows that rescale the jump lengths, a shock anneal in the middle, bounded and
    unbounded parameters (learning_chain.py:454-487), two consecutive run() calls.
    """
    from tests.plain_chain_loop import PlainLoopChain
    cfg = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    cfg.update(iterations_till_progress_update=False, jump_length_rescaling_factor=1.3, acceptance_window=10,
               shock_anneal_at=40, temperature_proposal=0.01)
    if not bounded:
        cfg['params_bounds'] = False
    ts = np.linspace(0, 2.5, 6)
    truth = lo.random_library_spec(np.random.default_rng(20260012), 1, 2)
    targets = [np.asarray(e).real for e in lo.expect_exact(truth, ts, OBS3)]
    out = []
    for cls in (PlainLoopChain, eml.LearningChain):
        np.random.seed(77)
        chain = cls(target_times=ts, target_datasets=targets, target_observables=OBS3, initial=(1, 2),
                    qubit_initial_state=definitions.ops['plus'], defect_initial_state=definitions.ops['mm'],
                    max_chain_steps=400, **copy.deepcopy(cfg))
        with np.errstate(all='ignore'):
            chain.run(60)
            first = (list(chain.run_acceptance_tracker), list(chain.run_step_type_tracker))      # (reset by every run())
            chain.run(35)
        out.append((chain.explored_loss, first[0] + chain.run_acceptance_tracker, first[1] + chain.run_step_type_tracker,
                    chain.explored_acceptance_probability, chain.chain_windows_acceptance_log,
                    dict(chain.params_handler.jump_lengths), chain.best_loss, np.random.uniform()))
    plain, fast = out
    assert len(plain[0]) == 98 and 'jump to best' in plain[2] and len(set(plain[2])) >= 6 and 5 < sum(plain[1]) < 95
    for a, b in zip(plain, fast):
        assert a == b


def test_weighted_choice_is_numpys_choice_draw_for_draw():
    """learning_chain._weighted_choice against np.random.choice(labels, p=p) (reference learning_chain.py:446-449) on one stream."""
    from effective_model_learning_b200.learning_chain import _weighted_choice
    gen = np.random.default_rng(11)
    labels = [f'move {i}' for i in range(13)]
    cases = []
    for _ in range(2000):
        w = gen.integers(0, 12, size=13).astype(float) * gen.choice([0.0, 1.0, 10.0], size=13)
        if w.sum() > 0:
            cases.append([x / w.sum() for x in w.tolist()])
    np.random.seed(5)
    want = [str(np.random.choice(labels, p=p)) for p in cases]
    tail_want = np.random.uniform()
    np.random.seed(5)
    got = [labels[_weighted_choice(p)] for p in cases]
    assert got == want and np.random.uniform() == tail_want


@pytest.mark.parametrize('D', [1, 2, 3])
def test_chain_replays_reference_trajectory(oracle_backend, D):
    """
    Same seed, same hyperparameters as the REAL reference chain recorded in make_golden.py: the port must
    propose the same moves, compute the same losses and take the same accept/reject decisions.
    """
    g = json.load(open(os.path.join(GOLD, f'reference_chain_D{D}.json')))
    config = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[g['config_index']])
    config['iterations_till_progress_update'] = False
    np.random.seed(g['seed'])
    chain = eml.LearningChain(target_times=np.array(g['target_times']),
                              target_datasets=[np.array(t) for t in g['targets']], target_observables=OBS3,
                              initial=(1, D), qubit_initial_state=definitions.ops['plus'],
                              defect_initial_state=definitions.ops['mm'], max_chain_steps=g['steps'],
                              store_all_proposals=True, **config)
    replay = iter(g['coupling_holder_is_first'])
    chain.process_handler.coupling_holder_choice = lambda first, second: next(replay)
    with np.errstate(over='ignore'):
        best = chain.run()
    assert chain.run_step_type_tracker == g['step_types']
    assert [bool(x) for x in chain.run_acceptance_tracker] == g['acceptance']
    np.testing.assert_allclose(chain.explored_loss, g['loss'], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(chain.explored_acceptance_probability, g['acceptance_probability'], rtol=1e-5)
    assert [chain.tot_RJ_steps, chain.acc_RJ_steps, chain.tot_tweak_steps, chain.acc_tweak_steps] == g['counters']
    vals, labels, _ = best.vectorise_under_library(config)
    assert labels == g['best_labels']
    np.testing.assert_allclose(vals, g['best_values'], rtol=1e-12)
    assert best.final_loss == pytest.approx(g['best_loss'], rel=1e-9)
    assert set(chain.all_proposals) == {'proposals', 'loss', 'acceptance', 'log_posterior', 'log_likelihood_prior',
                                        'acceptance_probability', 'annealed', 'step_types'}
    assert len(chain.explored_proposals) == 1 + sum(g['acceptance'])


def test_output_writes_the_reference_file_set(oracle_backend, tmp_path):
    """Scope row N3: data files of reference output.py, loadable without qutip."""
    import pickle
    from effective_model_learning_b200 import output
    g = json.load(open(os.path.join(GOLD, 'reference_chain_D1.json')))
    config = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    config['iterations_till_progress_update'] = False
    config['shock_anneal_at'] = 15
    np.random.seed(3)
    chain = eml.LearningChain(target_times=np.array(g['target_times'])[:30],
                              target_datasets=[np.array(t)[:30] for t in g['targets']], target_observables=OBS3,
                              initial=(1, 1), qubit_initial_state=definitions.ops['plus'],
                              defect_initial_state=definitions.ops['mm'], max_chain_steps=30,
                              store_all_proposals=True, **config)
    with np.errstate(over='ignore'):
        best = chain.run()

    class Toggles:
        comparison = False
        loss = True
        log_posterior = True
        acceptance_probability = True
        acceptance = True
        graphs = False
        pickle = True
        text = True
        all_proposals = True
        hyperparams = True

    base = str(tmp_path / 'run_D1_R1')
    out = output.Output(toggles=Toggles, filename=base, loss=chain.explored_loss, best_loss=chain.best_loss,
                        acceptance_probability=chain.explored_acceptance_probability,
                        overall_acceptance={'parameters tweak': chain.acc_tweak_steps / max(chain.tot_tweak_steps, 1),
                                            'reversible jump': chain.acc_RJ_steps / max(chain.tot_RJ_steps, 1)},
                        acceptance=chain.chain_windows_acceptance_log, models_to_save=[best], model_names=['best'],
                        chain_hyperparams=chain.get_init_hyperparams(), all_proposals=chain.all_proposals)
    names = {os.path.basename(p)[len('run_D1_R1'):] for p in out.written}
    assert {'_overall_acceptance.json', '_best.pickle', '_best.txt', '_proposals.pickle', '_hyperparameters.txt',
            '_hyperparameters.pickle', '_loss.pickle', '_accepted_loss.pickle', '_unannealed_accepted_loss.pickle',
            '_annealed_accepted_loss.pickle', '_accepted_log_posterior.pickle', '_AP.pickle',
            '_accepted_AP.pickle'} <= names
    loaded = pickle.load(open(base + '_best.pickle', 'rb'))
    assert loaded.vectorise_under_library(config)[0] == best.vectorise_under_library(config)[0]
    assert loaded.final_loss == chain.best_loss
    props = pickle.load(open(base + '_proposals.pickle', 'rb'))
    assert props['step_types'] == chain.run_step_type_tracker and 'jump to best' in props['step_types']
    assert len(pickle.load(open(base + '_accepted_loss.pickle', 'rb'))) == sum(chain.run_acceptance_tracker)
    assert pickle.load(open(base + '_loss.pickle', 'rb')) == chain.explored_loss


def test_total_dev_paths_agree_and_wrap_single_dataset(oracle_backend):
    """Fused device loss vs the host path taken when custom_function_on_dynamics_return is set; str/ndarray wrapping."""
    ts = np.linspace(0, 2.0, 25)
    spec = lo.random_library_spec(np.random.default_rng(12), 1, 2)
    truth = model_from_spec(spec, eml.LearningModel)
    data = [np.asarray(d) + 0.01 for d in lo.expect_exact(spec, ts, OBS3)]
    config = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    config['iterations_till_progress_update'] = False
    kw = dict(initial=(1, 2), qubit_initial_state=definitions.ops['plus'], defect_initial_state=definitions.ops['mm'])
    fused = eml.LearningChain(ts, data, OBS3, **kw, **config)
    config2 = dict(config, custom_function_on_dynamics_return=lambda r: r)
    hosted = eml.LearningChain(ts, data, OBS3, **kw, **config2)
    assert fused.total_dev(truth) == pytest.approx(hosted.total_dev(truth), rel=1e-12)
    assert fused.total_dev(truth) == pytest.approx(1e-4, rel=1e-9)          # every point is off by exactly 0.01
    single = eml.LearningChain(ts, data[0], 'sigmax', **kw, **config)
    assert isinstance(single.target_datasets, list) and single.target_observables == ['sigmax']
    assert single.total_dev(truth) == pytest.approx(1e-4, rel=1e-9)
    with pytest.raises(RuntimeWarning):
        eml.LearningChain(ts, data, OBS3, initial='nonsense', **config)
    with pytest.raises(RuntimeError):
        fused.step(truth, 'no such move')
    assert fused.complementary_step('add qubit L') == 'remove qubit L'
    assert fused.acceptance_probability((truth, 0.2), (truth, 0.1), 1.0, 1.0) > 1.0


def test_build_Liouvillian_is_the_corrected_column_stacked_superoperator():
    """
    Scope row a10.  BasicModel.build_Liouvillian returns L = -i(I (x) H - H^T (x) I) + sum_c [conj(c) (x) c - 1/2 I (x) c^dag c
    - 1/2 (c^dag c)^T (x) I] for vec with order='F' -- the superoperator qutip.liouvillian builds and mesolve integrates.
    DELIBERATE DEVIATION from reference basic_model.py:285-310, which uses L.T @ L where the column-stacking convention
    needs (L^dag L)^T and therefore is wrong for complex jump operators such as sigmay: for those the
    reference's formula differs from the oracle, this one does not.
    """
    from oracle import lindblad_oracle as lo
    from tests.helpers import model_from_spec, random_general_spec
    for seed in range(6):
        spec = random_general_spec(np.random.default_rng(700 + seed), 1 + seed % 3)
        m = model_from_spec(spec)
        L = m.build_Liouvillian()
        want = lo.liouvillian(lo.build_H(spec), lo.build_c_ops(spec))
        assert L.shape == want.shape and np.allclose(L, want, rtol=0, atol=1e-14)
        assert m.Liouvillian is L
        # it generates the dynamics the evaluation path reproduces: d vec(rho)/dt = L vec(rho)
        rho0 = lo.build_rho0(spec)
        v = scipy.linalg.expm(L * 0.3) @ rho0.ravel(order='F')
        assert abs(np.trace(v.reshape(rho0.shape, order='F')) - np.trace(rho0)) < 1e-12
    # the reference's own formula on a sigmay dissipator (what NOT to reproduce)
    spec = {'tls': [{'is_qubit': True, 'energy': 1.0, 'Ls': {'sigmay': 0.3}, 'initial_state': None, 'couplings': []}]}
    m = model_from_spec(spec)
    H, c = m.H, m.Ls[0]
    eye = np.eye(2)
    ref_formula = (-1j * (np.kron(eye, H) - np.kron(H.T, eye)) + np.kron(c.conj(), c)
                   - 0.5 * np.kron(eye, c.T @ c) - 0.5 * np.kron((c.T @ c).T, eye))       # L.T @ L as at basic_model.py:304-307
    assert np.abs(ref_formula - m.build_Liouvillian()).max() > 0.1
    empty = eml.BasicModel()
    with pytest.raises(RuntimeError):
        empty.build_Liouvillian()


def test_reference_analysis_scripts_can_load_what_output_writes(oracle_backend, tmp_path):
    """
    Scope row N3, consumer side.  The reference's clustering scripts `import learning_model`, unpickle *_best.pickle /
    *_proposals.pickle and call build_Liouvillian / vectorise_under_library / final_loss on the models
    (find_clusters.py:71-90, cluster_accepted.py:42-58, models_params_corr_mat.py:38-56, collate.py:122-128).  Files written
    with reference_module_names=True carry the reference's flat module names; a fresh interpreter that only installs the
    aliases -- no import of the package's classes by their package path -- runs the consumers' statements on them.
    """
    import pickle
    import subprocess
    import sys
    from effective_model_learning_b200 import output
    config = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    config['iterations_till_progress_update'] = False
    ts = np.linspace(0, 1.0, 12)
    targets = [np.cos(2 * ts), np.sin(2 * ts), np.zeros_like(ts)]
    np.random.seed(3)
    chain = eml.LearningChain(target_times=ts, target_datasets=targets, target_observables=OBS3, initial=(1, 1),
                              qubit_initial_state=definitions.ops['plus'], defect_initial_state=definitions.ops['mm'],
                              max_chain_steps=25, store_all_proposals=True, **config)
    best = chain.run()

    class Toggles:
        pickle = True
        all_proposals = True

    base = str(tmp_path / 'run_D1_R1')
    output.Output(toggles=Toggles, filename=base, models_to_save=[best], model_names=['best'],
                  all_proposals=chain.all_proposals, reference_module_names=True)
    raw = open(base + '_best.pickle', 'rb').read()
    assert b'learning_model' in raw and b'effective_model_learning_b200' not in raw
    assert type(best).__module__ == 'effective_model_learning_b200.learning_model'      # the renaming did not leak
    consumer = '''
import sys, pickle, json
import numpy as np
sys.path.insert(0, %r)
from effective_model_learning_b200 import compat
compat.install_reference_aliases()
import learning_model, configs                      # what the reference's scripts import (find_clusters.py:13-20)
with open(%r + '_best.pickle', 'rb') as filestream:
    new_model = pickle.load(filestream)             # find_clusters.py:71-73
assert type(new_model) is learning_model.LearningModel
Liouvillian_mat = new_model.build_Liouvillian()     # find_clusters.py:84-87
vect = np.concatenate((Liouvillian_mat.ravel().real, Liouvillian_mat.ravel().imag))
hyperparams = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
with open(%r + '_proposals.pickle', 'rb') as filestream:
    proposals = pickle.load(filestream)['proposals']    # models_params_corr_mat.py:38-39
vectors = [p.vectorise_under_library(hyperparameters=hyperparams)[0] for p in proposals]
print(json.dumps({'n': len(vect), 'loss': new_model.final_loss, 'proposals': len(vectors), 'width': len(vectors[0])}))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), base, base)
    out = subprocess.run([sys.executable, '-c', consumer], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.strip().splitlines()[-1])
    assert got['n'] == 2 * 16 * 16 and got['loss'] == pytest.approx(best.final_loss)
    assert got['proposals'] == len(chain.all_proposals['proposals']) and got['width'] == 10
    # and the default writer keeps the package path (loads wherever the package is importable, no aliases needed)
    output.Output(toggles=Toggles, filename=base + '_pkg', models_to_save=[best], model_names=['best'])
    with open(base + '_pkg_best.pickle', 'rb') as f:
        again = pickle.load(f)
    assert type(again) is type(best) and again.final_loss == best.final_loss
