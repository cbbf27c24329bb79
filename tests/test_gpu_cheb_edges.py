"""
GPU tests of the Chebyshev-series kernels (warp_cheb, cta_cheb) on the cases that stress the spectral enclosure and
the macro-step logic -- the same battery tests/test_cheb_model.py runs through the numpy model of the algorithm --
checked against the exact oracle at a much tighter tolerance than the 1e-8 of the parity tests.
"""
import numpy as np
import pytest

from oracle import lindblad_oracle as lo
from tests.helpers import OBS3, model_from_spec, max_err
from tests.test_cheb_model import EDGE, _spec

pytestmark = pytest.mark.gpu
TIGHT = 2e-12


@pytest.mark.parametrize('name', list(EDGE))
def test_edge_cases_on_the_gpu(name):
    from effective_model_learning_b200 import engine
    spec, times = EDGE[name]
    m = model_from_spec(spec)
    got = m.calculate_dynamics(times, OBS3)
    assert 'cheb' in engine.context_for(m, OBS3, times).plan.kernel_name
    want = lo.expect_exact(spec, times, OBS3)
    assert max_err(got, want) < TIGHT, (name, max_err(got, want))


def test_wide_spectrum_splits_into_macro_steps_at_every_size():
    """Couplings at the upper bound and a long grid: more terms than the trace tables hold (512 / 256 / 1024), so the
    end-of-step state path runs several times; d = 8, 16 and 32."""
    from effective_model_learning_b200 import engine
    for n_def, n_q in ((2, 1), (3, 1), (3, 2)):
        n = n_q + n_def
        couplings = [(0, k, 4.0, [('sigmax', 'sigmax')]) for k in range(n_q, n)] + \
                    [(i, j, 3.5, [('sigmay', 'sigmay')]) for i in range(n_q, n) for j in range(i + 1, n)]
        tls = [{'is_qubit': True, 'energy': 1.0, 'Ls': {'sigmaz': 0.05}, 'couplings': [], 'initial_state': lo.OPS['plus'].copy()}
               for _ in range(n_q)]
        tls += [{'is_qubit': False, 'energy': 0.2 + 0.05 * v, 'Ls': {'sigmax': 0.1}, 'couplings': [], 'initial_state': lo.OPS['mm'].copy()}
                for v in range(n_def)]
        for (i, j, s, pairs) in couplings:
            tls[i]['couplings'].append((j, s, pairs))
        spec = {'tls': tls}
        times = np.linspace(0.0, 30.0, 40)
        m = model_from_spec(spec)
        ctx = engine.context_for(m, OBS3, times)
        got = m.calculate_dynamics(times, OBS3)
        assert 'cheb' in ctx.plan.kernel_name
        want = lo.expect_exact(spec, times, OBS3)
        assert max_err(got, want) < 1e-11, (n, max_err(got, want))


def test_chebyshev_and_taylor_kernels_agree_on_a_batch():
    """The two series are independent implementations of the propagator action: 256 library models, d = 16."""
    from effective_model_learning_b200 import _native, configs, synthetic
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    lay = synthetic.LibraryLayout(1, 3, hyper)
    ts = np.linspace(0.0, 2.5, 200)
    p = lay.batch_params(5000 + np.arange(256))
    out = {}
    for name, kid in (('cheb', _native.KERNEL_WARP_CHEB), ('taylor', _native.KERNEL_WARP_MMA)):
        ctx = lay.context(ts, OBS3, kernel=kid)
        out[name], _ = ctx.evaluate(p, want_expect=True)
    assert np.abs(out['cheb'] - out['taylor']).max() < 1e-10


def test_parity_blocked_and_general_variants_of_the_d16_kernel():
    """Same-Pauli couplings conserve the parity of the basis index (H block diagonal, half the DMMAs); one sigma_z (x)
    sigma_x coupling breaks it and selects the general variant.  Both against the exact oracle, qubit not first."""
    from effective_model_learning_b200 import engine
    ts = np.linspace(0.0, 2.5, 120)
    base = lo.random_library_spec(np.random.default_rng(77), 1, 3)
    base['tls'] = [base['tls'][1], base['tls'][0], base['tls'][2], base['tls'][3]]      # observed site = 1
    remap = {0: 1, 1: 0, 2: 2, 3: 3}
    for t in base['tls']:
        t['couplings'] = [(remap[j], s, pairs) for (j, s, pairs) in t['couplings']]
    base['tls'][3]['Ls']['sigmam'] = 0.2
    base['tls'][0]['couplings'].append((2, 1.3, [('sigmay', 'sigmax')]))                 # flip (x) flip, complex H
    names = []
    for extra in (None, (1, 3, 0.9, [('sigmaz', 'sigmax')])):
        spec = {'tls': [dict(t, couplings=list(t['couplings']), Ls=dict(t['Ls'])) for t in base['tls']]}
        if extra:
            spec['tls'][extra[0]]['couplings'].append(extra[1:])
        m = model_from_spec(spec)
        got = m.calculate_dynamics(ts, OBS3)
        names.append(engine.context_for(m, OBS3, ts).plan.kernel_name)
        assert max_err(got, lo.expect_exact(spec, ts, OBS3)) < TIGHT, names
    assert names == ['warp_cheb_parity', 'warp_cheb'], names


def test_parity_blocked_d32_kernel_with_complex_hermitian_coupling():
    """d = 32, observed qubit at site 2, sigma-/+ jumps, one sigma_x (x) sigma_y coupling (flip (x) flip: parity is still
    conserved, H is complex) -> cta_cheb_parity; against the exact oracle."""
    from effective_model_learning_b200 import engine
    spec = lo.random_library_spec(np.random.default_rng(91), 2, 3)
    order = [2, 3, 0, 1, 4]                                  # new site k is old site order[k]
    inv = {old: new for new, old in enumerate(order)}
    tls = [spec['tls'][o] for o in order]
    for t in tls:
        t['couplings'] = [(inv[j], s, pairs) for (j, s, pairs) in t['couplings']]
    tls[0]['couplings'].append((4, 1.7, [('sigmax', 'sigmay')]))
    tls[1]['Ls']['sigmam'] = 0.25
    tls[2]['Ls']['sigmap'] = 0.15
    spec = {'tls': tls}
    ts = np.sort(np.concatenate([np.linspace(0.0, 2.5, 60), [7.0]]))
    m = model_from_spec(spec)
    got = m.calculate_dynamics(ts, OBS3)
    assert engine.context_for(m, OBS3, ts).plan.kernel_name == 'cta_cheb_parity'
    assert max_err(got, lo.expect_exact(spec, ts, OBS3)) < TIGHT


def test_steady_parity_sector_reduction_equals_the_full_evolution():
    """Production initial state (qubit |+><+|, defects maximally mixed) and a Pauli library: the parity-even half of rho
    is a steady state and the d = 16 kernel evolves the odd tiles only.  Same results as with the reduction disabled,
    also on a long non-uniform grid that needs several macro steps; other initial states / non-unital jump operators
    do not select it."""
    from effective_model_learning_b200 import _native, configs, definitions, synthetic
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    lay = synthetic.LibraryLayout(1, 3, hyper)
    p = lay.batch_params(7000 + np.arange(200))
    rng = np.random.default_rng(3)
    for ts in (np.linspace(0.0, 2.5, 200), np.sort(np.concatenate([[0.0, 0.0], rng.uniform(0, 40.0, 50), [40.0, 40.0]]))):
        out = {}
        for name, flags in (('sector', 0), ('full', _native.FLAG_NO_SECTOR_REDUCTION)):
            ctx = lay.context(ts, ['sigmax', 'sigmap', 'sigmaz'], flags=flags)
            ctx.set_targets(np.zeros((3, len(ts))))
            out[name] = ctx.evaluate(p, want_expect=True, want_loss=True)
            assert ('sector' in ctx.plan.kernel_name) == (flags == 0), ctx.plan.kernel_name
        assert np.abs(out['sector'][0] - out['full'][0]).max() < 1e-12
        np.testing.assert_allclose(out['sector'][1], out['full'][1], rtol=1e-11, atol=1e-15)
    spec_row = lay.batch_params([7000])[0]
    want = lo.expect_exact(__import__('tests.helpers', fromlist=['x']).spec_from_context_row(
        lay.context(np.linspace(0.0, 2.5, 200), OBS3), spec_row), np.linspace(0.0, 2.5, 200), OBS3)
    got, _ = lay.context(np.linspace(0.0, 2.5, 200), OBS3).evaluate(spec_row[None], want_expect=True)
    assert max_err(list(got[0].real), want) < TIGHT
    # a polarised defect state: the even sector is not a multiple of the identity
    ctx = lay.context(np.linspace(0.0, 2.5, 50), OBS3, defect_state=definitions.ops['gnd'])
    ctx.evaluate(p[:4], want_expect=True)
    assert 'sector' not in ctx.plan.kernel_name and 'parity' in ctx.plan.kernel_name
