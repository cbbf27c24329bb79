"""GPU parity: CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from oracle import lindblad_oracle as lo
from tests.helpers import OBS3, model_from_spec, assert_close, max_err

pytestmark = pytest.mark.gpu

TOL = 1e-8   # north_star: observables within 1e-8 (relative, floor 1) of the exact solution


@pytest.mark.parametrize('n_defects', [0, 1, 2, 3])
def test_random_library_models_match_oracle(n_defects):
    ts = np.linspace(0.0, 2.5, 200)
    worst = 0.0
    for b in range(6):
        spec = lo.random_library_spec(np.random.default_rng(1000 + b), 1, n_defects)
        want = lo.expect_exact(spec, ts, OBS3)
        got = model_from_spec(spec).calculate_dynamics(ts, OBS3)
        assert_close(got, want, TOL, f'D={n_defects} b={b}')
        assert all(g.dtype == np.float64 for g in got)
        worst = max(worst, max_err(got, want))
    print(f'd={2 ** (1 + n_defects)} worst abs err vs exact oracle: {worst:.2e}')


def test_two_qubits_three_defects_d32():
    ts = np.linspace(0.0, 2.5, 60)
    spec = lo.random_library_spec(np.random.default_rng(7), 2, 3)
    want = lo.expect_exact(spec, ts, OBS3)
    got = model_from_spec(spec).calculate_dynamics(ts, OBS3)
    assert_close(got, want, TOL, 'd=32')


def test_default_observable_and_nonhermitian_observable():
    ts = np.linspace(0.0, 1.0, 33)
    spec = lo.random_library_spec(np.random.default_rng(3), 1, 2)
    m = model_from_spec(spec)
    got = m.calculate_dynamics(ts)                       # default 'exc'
    assert_close(got, lo.expect_exact(spec, ts, False), TOL, 'exc')
    got = m.calculate_dynamics(ts, ['sigmap', 'sigmam', 'sigmaz'])
    want = lo.expect_exact(spec, ts, ['sigmap', 'sigmam', 'sigmaz'])
    assert got[0].dtype == np.complex128 and got[2].dtype == np.float64
    assert_close(got, want, TOL, 'non-hermitian obs')


def test_nonuniform_grid_and_repeated_times():
    rng = np.random.default_rng(11)
    ts = np.sort(np.concatenate([[0.3], 0.3 + np.cumsum(rng.uniform(1e-4, 0.2, 70))]))
    ts[10] = ts[9]                                        # a repeated time point (zero-length interval)
    spec = lo.random_library_spec(np.random.default_rng(5), 1, 2)
    got = model_from_spec(spec).calculate_dynamics(ts, OBS3)
    assert_close(got, lo.expect_exact(spec, ts, OBS3), TOL, 'non-uniform')


def test_long_interval_needs_substeps():
    ts = np.array([0.0, 5.0, 5.5, 30.0])
    spec = lo.random_library_spec(np.random.default_rng(9), 1, 2)
    got = model_from_spec(spec).calculate_dynamics(ts, OBS3)
    assert_close(got, lo.expect_zvode(spec, ts, OBS3), 1e-7, 'long intervals (vs tight ZVODE)')
    assert_close(got, lo.expect_exact(spec, ts, OBS3), TOL, 'long intervals')


def test_general_ops_nonhermitian_generator():
    """sigmam/sigmap jump operators, asymmetric op pairs (non-Hermitian H), non-default initial states."""
    spec = {'tls': [
        {'is_qubit': True, 'energy': 1.0, 'Ls': {'sigmam': 0.2, 'sigmaz': 0.05}, 'initial_state': lo.OPS['2e+1g'],
         'couplings': [(1, 0.8, [('sigmap', 'sigmam'), ('sigmam', 'sigmap')]), (2, 0.5, [('sigmap', 'sigmam')])]},
        {'is_qubit': False, 'energy': 0.3, 'Ls': {'sigmap': 0.1, 'sigmax': 0.07}, 'initial_state': None, 'couplings': []},
        {'is_qubit': False, 'energy': 0.2, 'Ls': {'exc': 0.3}, 'initial_state': lo.OPS['mm'],
         'couplings': [(1, 1.1, [('sigmax', 'sigmaz')])]},
    ]}
    ts = np.linspace(0, 3, 50)
    got = model_from_spec(spec).calculate_dynamics(ts, ['sigmax', 'sigmap', 'exc'])
    assert_close(got, lo.expect_exact(spec, ts, ['sigmax', 'sigmap', 'exc']), TOL, 'general ops')


def test_single_time_and_single_qubit():
    spec = {'tls': [{'is_qubit': True, 'energy': 1.0, 'Ls': {'sigmaz': 0.1}, 'initial_state': None, 'couplings': []}]}
    m = model_from_spec(spec)
    got = m.calculate_dynamics(np.array([0.7]), OBS3)
    assert_close(got, lo.expect_exact(spec, np.array([0.7]), OBS3), TOL, 'single time')
    ts = np.linspace(0, 4, 101)
    got = m.calculate_dynamics(ts, OBS3)
    # closed form: plus state under E sigmaz with sigmaz dephasing gamma (SURVEY.md section 4 iii, iv)
    np.testing.assert_allclose(got[0], np.exp(-2 * 0.1 * ts) * np.cos(2 * ts), atol=1e-10)
    np.testing.assert_allclose(got[1], np.exp(-2 * 0.1 * ts) * np.sin(2 * ts), atol=1e-10)
    np.testing.assert_allclose(got[2], 0 * ts, atol=1e-10)


def test_errors_match_reference_contract():
    from effective_model_learning_b200 import BasicModel
    m = BasicModel()
    m.add_TLS('qubit', True, 1.0)
    with pytest.raises(RuntimeError):
        m.calculate_dynamics(np.linspace(0, 1, 5))       # build_operators() not called
    m.build_operators()
    with pytest.raises(RuntimeError):
        m.calculate_dynamics(np.linspace(0, 1, 5), observable_ops=[1, 2])


@pytest.mark.parametrize('D', [1, 2, 3])      # D = 3: the headline size (d = 16, sector kernel)
def test_chain_on_gpu_replays_reference_accept_reject_sequence(D):
    """north_star: for fixed seeds the accept/reject sequence matches the reference's (losses to 1e-8)."""
    import json
    import os
    import effective_model_learning_b200 as eml
    from effective_model_learning_b200 import configs, definitions
    g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', f'reference_chain_D{D}.json')))
    config = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[g['config_index']])
    config['iterations_till_progress_update'] = False
    np.random.seed(g['seed'])
    chain = eml.LearningChain(target_times=np.array(g['target_times']),
                              target_datasets=[np.array(t) for t in g['targets']], target_observables=OBS3,
                              initial=(1, D), qubit_initial_state=definitions.ops['plus'],
                              defect_initial_state=definitions.ops['mm'], max_chain_steps=g['steps'], **config)
    replay = iter(g['coupling_holder_is_first'])
    chain.process_handler.coupling_holder_choice = lambda first, second: next(replay)
    with np.errstate(over='ignore'):
        chain.run()
    assert chain.run_step_type_tracker == g['step_types']
    assert [bool(x) for x in chain.run_acceptance_tracker] == g['acceptance']
    np.testing.assert_allclose(chain.explored_loss, g['loss'], rtol=1e-8, atol=1e-12)


def test_frozen_golden_observables_d4_to_d32():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'oracle_observables.npz'))
    ts = z['times']
    from effective_model_learning_b200 import engine
    for (Q, D) in ((1, 1), (1, 2), (1, 3), (2, 3)):
        seeds = z[f'seeds_Q{Q}_D{D}']
        models = [model_from_spec(lo.random_library_spec(np.random.default_rng(int(s)), Q, D)) for s in seeds]
        expect, _ = engine.evaluate_models(models, ts, OBS3)
        err = np.abs(expect.real - z[f'expect_Q{Q}_D{D}']).max()
        assert err < TOL, (Q, D, err)
        assert np.abs(expect.imag).max() < 1e-12


# ---------------------------------------------------------------------------------------------
# d = 16 specialised kernel (one warp per model, FP64 DMMA) against the oracle and the generic kernel
@pytest.fixture
def use_kernel():
    from effective_model_learning_b200 import engine, _native
    saved = engine.default_kernel

    def select(name):
        engine.default_kernel = {'auto': _native.KERNEL_AUTO, 'generic': _native.KERNEL_GENERIC,
                                 'warp_mma': _native.KERNEL_WARP_MMA, 'cta_mma': _native.KERNEL_CTA_MMA,
                                 'warp_cheb': _native.KERNEL_WARP_CHEB, 'cta_cheb': _native.KERNEL_CTA_CHEB}[name]
        engine.clear_contexts()
    yield select
    engine.default_kernel = saved
    engine.clear_contexts()


def _d16_specs():
    specs = [lo.random_library_spec(np.random.default_rng(1000 + b), 1, 3) for b in range(12)]
    # qubit NOT first (observable site = 2), amplitude-damping style jumps, projector jumps, default initial states
    specs.append({'tls': [
        {'is_qubit': False, 'energy': 0.3, 'Ls': {'sigmam': 0.2, 'sigmax': 0.05}, 'initial_state': None, 'couplings': []},
        {'is_qubit': False, 'energy': 0.12, 'Ls': {'exc': 0.3, 'sigmap': 0.11}, 'initial_state': lo.OPS['mm'],
         'couplings': [(0, 1.4, [('sigmax', 'sigmax')]), (2, 0.9, [('sigmay', 'sigmay'), ('sigmaz', 'sigmaz')])]},
        {'is_qubit': True, 'energy': 1.0, 'Ls': {'sigmaz': 0.07, 'sigmam': 0.13, 'gnd': 0.2},
         'initial_state': lo.OPS['2e+1g'], 'couplings': [(3, 2.2, [('sigmax', 'sigmax')])]},
        {'is_qubit': True, 'energy': 0.8, 'Ls': {'sigmay': 0.3}, 'initial_state': None,
         'couplings': [(0, 0.6, [('sigmaz', 'sigmax')])]},
    ]})
    return specs


@pytest.mark.parametrize('kernel', ['warp_cheb', 'warp_mma', 'generic'])
def test_d16_kernels_match_oracle(use_kernel, kernel):
    from effective_model_learning_b200 import engine
    use_kernel(kernel)
    ts = np.linspace(0.0, 2.5, 200)
    worst = 0.0
    for i, spec in enumerate(_d16_specs()):
        m = model_from_spec(spec)
        got = m.calculate_dynamics(ts, OBS3)
        assert engine.context_for(m, OBS3, ts).plan.kernel_name.startswith(kernel)
        want = lo.expect_exact(spec, ts, OBS3)
        assert_close(got, want, TOL, f'{kernel} model {i}')
        worst = max(worst, max_err(got, want))
    print(f'{kernel}: worst abs err {worst:.2e}')


@pytest.mark.parametrize('kernel', ['warp_cheb', 'warp_mma'])
def test_warp_kernels_nonuniform_grid_long_steps_and_complex_observables(use_kernel, kernel):
    use_kernel(kernel)
    rng = np.random.default_rng(4)
    ts = np.sort(np.concatenate([[0.1], 0.1 + np.cumsum(rng.uniform(1e-4, 0.3, 60)), [40.0]]))
    ts[20] = ts[19]
    for spec in _d16_specs()[10:]:
        m = model_from_spec(spec)
        obs = ['sigmap', 'exc', 'sigmay']
        assert_close(m.calculate_dynamics(ts, obs), lo.expect_exact(spec, ts, obs), TOL, f'{kernel} non-uniform')


def test_warp_mma_loss_and_batch_equals_generic(use_kernel):
    from effective_model_learning_b200 import engine
    ts = np.linspace(0.0, 2.5, 200)
    specs = [lo.random_library_spec(np.random.default_rng(2000 + b), 1, 3) for b in range(300)]
    models = [model_from_spec(s) for s in specs]
    targets = np.array([np.asarray(e) for e in lo.expect_exact(specs[0], ts, OBS3)]) + 0.01
    out = {}
    for kernel in ('generic', 'warp_mma', 'warp_cheb'):
        use_kernel(kernel)
        out[kernel] = engine.evaluate_models(models, ts, OBS3, targets=targets)
    for kernel in ('warp_mma', 'warp_cheb'):
        assert np.abs(out['generic'][0] - out[kernel][0]).max() < 1e-10, kernel
        np.testing.assert_allclose(out['generic'][1], out[kernel][1], rtol=1e-9, atol=1e-14)
    want = np.array([lo.total_dev(lo.expect_exact(s, ts, OBS3), list(targets), len(ts)) for s in specs[:8]])
    np.testing.assert_allclose(out['warp_cheb'][1][:8], want, rtol=1e-8, atol=1e-14)


@pytest.mark.parametrize('kernel', ['warp_cheb', 'warp_mma', 'generic'])
@pytest.mark.parametrize('QD', [(1, 0), (1, 1), (1, 2)])
def test_small_models_both_kernels(use_kernel, kernel, QD):
    """d = 2, 4 (embedded into 3 sites by the warp kernel) and d = 8, incl. qubit-not-first and sigma-/+ jumps."""
    from effective_model_learning_b200 import engine
    use_kernel(kernel)
    ts = np.linspace(0.0, 2.5, 200)
    specs = [lo.random_library_spec(np.random.default_rng(3000 + b), *QD) for b in range(6)]
    if QD == (1, 2):
        specs.append({'tls': [
            {'is_qubit': False, 'energy': 0.3, 'Ls': {'sigmam': 0.2, 'sigmax': 0.05}, 'initial_state': None, 'couplings': []},
            {'is_qubit': True, 'energy': 1.0, 'Ls': {'sigmaz': 0.07, 'sigmap': 0.13, 'gnd': 0.2},
             'initial_state': lo.OPS['2e+1g'], 'couplings': [(0, 1.4, [('sigmax', 'sigmax')]), (2, 2.2, [('sigmay', 'sigmay')])]},
            {'is_qubit': False, 'energy': 0.12, 'Ls': {'exc': 0.3, 'sigmay': 0.11}, 'initial_state': lo.OPS['mm'],
             'couplings': [(0, 0.6, [('sigmaz', 'sigmax')])]},
        ]})
    for i, spec in enumerate(specs):
        m = model_from_spec(spec)
        obs = ['sigmax', 'sigmap', 'sigmaz']
        got = m.calculate_dynamics(ts, obs)
        assert engine.context_for(m, obs, ts).plan.kernel_name.startswith(kernel)
        assert_close(got, lo.expect_exact(spec, ts, obs), TOL, f'{kernel} {QD} model {i}')


def test_d8_batch_loss_equals_generic(use_kernel):
    from effective_model_learning_b200 import engine
    ts = np.sort(np.random.default_rng(0).uniform(0, 3.0, 150))
    specs = [lo.random_library_spec(np.random.default_rng(4000 + b), 1, 2) for b in range(400)]
    models = [model_from_spec(s) for s in specs]
    targets = np.array([np.asarray(e) for e in lo.expect_exact(specs[0], ts, OBS3)]) + 0.01
    out = {}
    for kernel in ('generic', 'warp_mma', 'warp_cheb'):
        use_kernel(kernel)
        out[kernel] = engine.evaluate_models(models, ts, OBS3, targets=targets)
    for kernel in ('warp_mma', 'warp_cheb'):
        assert np.abs(out['generic'][0] - out[kernel][0]).max() < 1e-10, kernel
        np.testing.assert_allclose(out['generic'][1], out[kernel][1], rtol=1e-9, atol=1e-14)


def test_auto_dispatch_falls_back_to_generic_when_structure_is_not_supported(use_kernel):
    from effective_model_learning_b200 import engine
    use_kernel('auto')
    ts = np.linspace(0, 1, 20)
    spec = lo.random_library_spec(np.random.default_rng(1), 1, 3)
    m = model_from_spec(spec)
    m.calculate_dynamics(ts, OBS3)
    assert engine.context_for(m, OBS3, ts).plan.kernel_name == 'warp_cheb_parity_sector_realH'
    spec['tls'][0]['couplings'].append((1, 0.7, [('sigmap', 'sigmam')]))      # non-Hermitian H
    m = model_from_spec(spec)
    got = m.calculate_dynamics(ts, OBS3)
    assert engine.context_for(m, OBS3, ts).plan.kernel_name == 'generic'
    assert_close(got, lo.expect_exact(spec, ts, OBS3), TOL, 'fallback')
    spec = lo.random_library_spec(np.random.default_rng(1), 1, 3)
    spec['tls'][1]['Ls']['plus'] = 0.2                                       # dense real jump operator
    m = model_from_spec(spec)
    got = m.calculate_dynamics(ts, OBS3)
    assert engine.context_for(m, OBS3, ts).plan.kernel_name == 'generic'
    assert_close(got, lo.expect_exact(spec, ts, OBS3), TOL, 'fallback 2')


def test_posterior_predictive_and_dense_grid_callers():
    """Scope row N2: batched re-evaluation of sampled parameter vectors and the dense post-run grid."""
    from effective_model_learning_b200 import configs, predictive, LearningModel, definitions
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    D = 2
    ts = np.linspace(0.0, 2.5, 120)
    vectors, specs = [], []
    for b in range(40):
        spec = lo.random_library_spec(np.random.default_rng(6000 + b), 1, D)
        m = model_from_spec(spec, LearningModel)
        vectors.append(m.vectorise_under_library(hyper)[0])
        specs.append(spec)
    means, stds, evaluated = predictive.posterior_predictive(vectors, hyper, ts, OBS3, D=D,
                                                             qubit_initial_state=definitions.ops['plus'],
                                                             defect_initial_state=definitions.ops['mm'])
    want = np.array([[np.asarray(e) for e in lo.expect_exact(s, ts, OBS3)] for s in specs])
    for m_, name in enumerate(OBS3):
        assert np.abs(evaluated[name] - want[:, m_]).max() < TOL
        np.testing.assert_allclose(means[name], want[:, m_].mean(axis=0), atol=1e-9)
        np.testing.assert_allclose(stds[name], want[:, m_].std(axis=0), atol=1e-9)
    grid, data = predictive.dense_grid_dynamics(model_from_spec(specs[0]), ts, OBS3)
    assert len(grid) == 1200 and grid[0] == ts[0] and grid[-1] == ts[-1]
    assert_close(data, lo.expect_exact(specs[0], grid, OBS3), TOL, 'dense grid')


# ---------------------------------------------------------------------------------------------
# d = 32 kernel (four warps per model)
def _d32_specs():
    specs = [lo.random_library_spec(np.random.default_rng(7000 + b), 2, 3) for b in range(5)]
    # first qubit NOT at site 0, sigma-/+ and projector jumps, a complex Hermitian coupling (sigmax (x) sigmay)
    specs.append({'tls': [
        {'is_qubit': False, 'energy': 0.3, 'Ls': {'sigmam': 0.2, 'sigmax': 0.05}, 'initial_state': None, 'couplings': []},
        {'is_qubit': False, 'energy': 0.12, 'Ls': {'exc': 0.3, 'sigmap': 0.11}, 'initial_state': lo.OPS['mm'],
         'couplings': [(0, 1.4, [('sigmax', 'sigmax')]), (2, 0.9, [('sigmay', 'sigmay'), ('sigmaz', 'sigmaz')])]},
        {'is_qubit': True, 'energy': 1.0, 'Ls': {'sigmaz': 0.07, 'sigmam': 0.13, 'gnd': 0.2},
         'initial_state': lo.OPS['2e+1g'], 'couplings': [(3, 2.2, [('sigmax', 'sigmax')]), (4, 0.7, [('sigmax', 'sigmay')])]},
        {'is_qubit': True, 'energy': 0.8, 'Ls': {'sigmay': 0.3}, 'initial_state': None,
         'couplings': [(0, 0.6, [('sigmaz', 'sigmax')])]},
        {'is_qubit': False, 'energy': 0.21, 'Ls': {'sigmaz': 0.15, 'sigmay': 0.02}, 'initial_state': lo.OPS['mm'],
         'couplings': [(1, 1.1, [('sigmaz', 'sigmaz')])]},
    ]})
    return specs


@pytest.mark.parametrize('kernel', ['cta_cheb', 'cta_mma', 'generic'])
def test_d32_kernels_match_oracle(use_kernel, kernel):
    from effective_model_learning_b200 import engine
    use_kernel(kernel)
    rng = np.random.default_rng(8)
    ts = np.sort(np.concatenate([np.linspace(0.0, 2.5, 70), [2.5 + 9.0], rng.uniform(0, 2.5, 10)]))
    worst = 0.0
    for i, spec in enumerate(_d32_specs()):
        m = model_from_spec(spec)
        obs = ['sigmax', 'sigmap', 'sigmaz']
        got = m.calculate_dynamics(ts, obs)
        assert engine.context_for(m, obs, ts).plan.kernel_name.startswith(kernel)
        want = lo.expect_exact(spec, ts, obs)
        assert_close(got, want, TOL, f'{kernel} d=32 model {i}')
        worst = max(worst, max_err(got, want))
    print(f'{kernel} d=32: worst abs err {worst:.2e}')


def test_d32_batch_loss_equals_generic(use_kernel):
    from effective_model_learning_b200 import engine
    ts = np.linspace(0.0, 2.5, 40)
    specs = [lo.random_library_spec(np.random.default_rng(9000 + b), 2, 3) for b in range(330)]
    models = [model_from_spec(s) for s in specs]
    targets = np.array([np.asarray(e) for e in lo.expect_exact(specs[0], ts, OBS3)]) + 0.01
    out = {}
    for kernel in ('generic', 'cta_mma', 'cta_cheb'):
        use_kernel(kernel)
        out[kernel] = engine.evaluate_models(models, ts, OBS3, targets=targets)
    for kernel in ('cta_mma', 'cta_cheb'):
        assert np.abs(out['generic'][0] - out[kernel][0]).max() < 1e-10, kernel
        np.testing.assert_allclose(out['generic'][1], out[kernel][1], rtol=1e-9, atol=1e-14)


def test_c_abi_edge_cases_and_error_codes():
    """Empty batch, eight observables, a single time, bad descriptors -> RuntimeError with the library's message."""
    from effective_model_learning_b200 import _native, definitions, engine
    ops = definitions.ops
    op_table = np.stack([ops['sigmaz'], ops['sigmax']])
    rho0 = np.kron(ops['plus'], ops['mm'])
    terms = [(0, _native.TERM_ENERGY, 0, -1, 0, -1), (1, _native.TERM_COUPLING, 0, 1, 1, 1), (2, _native.TERM_LINDBLAD, 1, -1, 0, -1)]
    ts = np.linspace(0, 1, 9)
    names = ['sigmax', 'sigmay', 'sigmaz', 'sigmap', 'sigmam', 'exc', 'gnd', 'id2']
    obs = np.stack([ops[n] for n in names])
    plan = _native.Plan(2, 3, terms, op_table, rho0, 0, obs, ts)
    ex, loss = plan.eval_host(np.zeros((0, 3)), want_expect=True, want_loss=False)
    assert ex.shape == (0, 8, 9)
    ex, _ = plan.eval_host(np.array([[1.0, 0.9, 0.2]]), want_expect=True, want_loss=False)
    spec = {'tls': [{'is_qubit': True, 'energy': 1.0, 'Ls': {}, 'initial_state': lo.OPS['plus'],
                     'couplings': [(1, 0.9, [('sigmax', 'sigmax')])]},
                    {'is_qubit': False, 'energy': 0.0, 'Ls': {'sigmaz': 0.2}, 'initial_state': lo.OPS['mm'], 'couplings': []}]}
    want = lo.expect_exact(spec, ts, names)
    for m in range(8):
        assert np.abs(ex[0, m] - want[m]).max() < TOL
    assert abs(ex[0, 7] - 1).max() < 1e-12                    # Tr(rho) stays 1
    # loss requested without targets is defined as 0; targets of the wrong shape are refused
    _, loss = plan.eval_host(np.array([[1.0, 0.9, 0.2]]), want_expect=False, want_loss=True)
    assert loss[0] == 0.0
    with pytest.raises(RuntimeError):
        plan.set_targets(np.zeros((3, 9)))
    plan.close()
    with pytest.raises(RuntimeError, match='non-decreasing'):
        _native.Plan(2, 3, terms, op_table, rho0, 0, obs, ts[::-1].copy())
    with pytest.raises(RuntimeError, match='malformed'):
        _native.Plan(2, 3, [(5, 0, 0, -1, 0, -1)], op_table, rho0, 0, obs, ts)
    with pytest.raises(RuntimeError, match='n_tls'):
        _native.Plan(6, 0, [], op_table, np.eye(64), 0, obs, ts)
    with pytest.raises(RuntimeError, match='n_obs'):
        _native.Plan(2, 3, terms, op_table, rho0, 0, np.stack([ops['id2']] * 9), ts)
    with pytest.raises(RuntimeError, match='warp kernels'):
        _native.Plan(5, 0, [], op_table, np.eye(32) / 32, 0, obs, ts, kernel=_native.KERNEL_WARP_MMA)


@pytest.mark.parametrize('QD', [(1, 1), (1, 2), (1, 3), (2, 3)])
def test_trailing_and_isolated_repeated_times(QD):
    """Zero-length macro steps (the last times equal; a repeat followed by a long interval) must not produce NaNs."""
    ts = np.array([0.0, 0.0, 0.4, 0.4, 0.4, 30.0, 30.0, 30.5, 30.5])
    spec = lo.random_library_spec(np.random.default_rng(77), *QD)
    got = model_from_spec(spec).calculate_dynamics(ts, OBS3)
    assert all(np.all(np.isfinite(g)) for g in got)
    assert_close(got, lo.expect_exact(spec, ts, OBS3), TOL, f'repeated times {QD}')


@pytest.mark.parametrize('kernel,QD', [('auto', (1, 2)), ('auto', (1, 3)), ('generic', (1, 3)), ('auto', (2, 3))])
def test_non_finite_parameters_fail_loudly_instead_of_hanging(use_kernel, kernel, QD):
    from effective_model_learning_b200 import configs, synthetic
    use_kernel(kernel)
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    lay = synthetic.LibraryLayout(*QD, hyper)
    ctx = lay.context(np.linspace(0, 2.5, 30), OBS3)
    good = lay.batch_params(1000 + np.arange(3))
    bad = good.copy()
    bad[1, 2] = np.nan
    with pytest.raises(RuntimeError, match='Taylor cap|not converge|hit'):
        ctx.evaluate(bad, want_expect=True)
    huge = good.copy()
    huge[0, lay.n_library - 1] = 1e9          # a rate of 1e9: nu * span far beyond anything sensible
    with pytest.raises(RuntimeError):
        ctx.evaluate(huge, want_expect=True)
    ex, _ = ctx.evaluate(good, want_expect=True)      # the plan is still usable afterwards
    assert np.all(np.isfinite(ex))


@pytest.mark.parametrize('n', [2, 3, 4, 5])
def test_randomised_general_models_dmma_vs_generic_vs_oracle(use_kernel, n):
    """Stress: 40 random general models per size; DMMA kernels == generic kernel (1e-10), a few also vs the oracle."""
    from effective_model_learning_b200 import engine
    from tests.helpers import random_general_spec
    rng = np.random.default_rng(100 + n)
    ts = np.sort(np.concatenate([[0.0], rng.uniform(0, 3.0, 45), [3.0, 3.0, 9.0]]))
    obs = ['sigmax', 'sigmam', 'exc']
    specs = [random_general_spec(np.random.default_rng(31000 + 97 * n + i), n) for i in range(40)]
    results = {}
    names = set()
    for kernel in ('auto', 'generic'):
        use_kernel(kernel)
        out = []
        for spec in specs:
            m = model_from_spec(spec)
            out.append(m.calculate_dynamics(ts, obs))
            names.add((kernel, engine.context_for(m, obs, ts).plan.kernel_name))
        results[kernel] = out
    assert any(k == 'auto' and name != 'generic' for k, name in names), names
    for a, b in zip(results['auto'], results['generic']):
        for x, y in zip(a, b):
            assert np.abs(np.asarray(x) - np.asarray(y)).max() < 1e-10
    for i in (0, 7, 19):
        assert_close(results['auto'][i], lo.expect_exact(specs[i], ts, obs), TOL, f'n={n} model {i}')


def test_small_host_batches_replayed_from_a_graph_equal_the_eager_path():
    """
    eml_eval_batch_host replays a CUDA graph of the whole call for batches of <= 8 models from the third such call on
    (a single chain's total_dev is one model per call).  Same results as the eager path, also after the targets change
    (they are kernel arguments of the captured graph) and when the call shape alternates.
    """
    from effective_model_learning_b200 import configs, synthetic
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    ts = np.linspace(0.0, 2.5, 40)
    for D in (1, 3):
        lay = synthetic.LibraryLayout(1, D, hyper)
        ctx = lay.context(ts, OBS3)
        rows = lay.batch_params(1000 + np.arange(12))
        tg1 = np.random.default_rng(1).normal(0, 0.3, (3, len(ts)))
        tg2 = np.random.default_rng(2).normal(0, 0.3, (3, len(ts)))
        ctx.set_targets(tg1)
        ex_all, loss1 = ctx.evaluate(rows, want_expect=True, want_loss=True)            # 12 models: eager
        for i in range(12):                                                             # one model per call: graph from call 3
            ex, loss = ctx.evaluate(rows[i:i + 1], want_expect=True, want_loss=True)
            assert np.array_equal(ex[0], ex_all[i]) and loss[0] == loss1[i], (D, i)
            _, only_loss = ctx.evaluate(rows[i:i + 1], want_expect=False, want_loss=True)   # another call shape
            assert only_loss[0] == loss1[i]
        for i in range(0, 12, 3):
            ex, loss = ctx.evaluate(rows[i:i + 3], want_expect=True, want_loss=True)
            assert np.array_equal(ex, ex_all[i:i + 3]) and np.array_equal(loss, loss1[i:i + 3])
        ctx.set_targets(tg2)
        _, loss2 = ctx.evaluate(rows, want_expect=False, want_loss=True)
        assert not np.allclose(loss1, loss2)
        for i in range(4):
            _, loss = ctx.evaluate(rows[i:i + 1], want_expect=False, want_loss=True)
            assert loss[0] == loss2[i]
