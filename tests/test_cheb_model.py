"""
CPU tests of the numpy model of the Chebyshev-series algorithm the warp_cheb / cta_cheb kernels run
(oracle/chebyshev_action_model.py): accuracy against the exact oracle on library models and on the edge cases that
stress the spectral enclosure, validity of the term-count formula, step splitting.
"""
import numpy as np
import pytest
import scipy.special as sp

from oracle import chebyshev_action_model as cm
from oracle import lindblad_oracle as lo
from oracle import taylor_action_model as tm

OBS3 = ['sigmax', 'sigmay', 'sigmaz']


def _run(spec, times, obs=OBS3, **kw):
    H, c_ops, rho0 = lo.build_H(spec), lo.build_c_ops(spec), lo.build_rho0(spec)
    out, napp, info = cm.evaluate(H, c_ops, rho0, lo.build_observables(spec, obs), times, **kw)
    exact = np.array([np.asarray(e) for e in lo.expect_exact(spec, times, obs)])
    return np.abs(out - exact).max(), napp, info, (H, c_ops, rho0)


def _spec(n_def, energies, couplings, Ls, qubit_energy=1.0):
    tls = [{'is_qubit': True, 'energy': qubit_energy, 'Ls': {}, 'couplings': [], 'initial_state': lo.OPS['plus'].copy()}]
    for v in range(n_def):
        tls.append({'is_qubit': False, 'energy': energies[v], 'Ls': {}, 'couplings': [], 'initial_state': lo.OPS['mm'].copy()})
    for (i, j, s, pairs) in couplings:
        tls[i]['couplings'].append((j, s, pairs))
    for (k, op, g) in Ls:
        tls[k]['Ls'][op] = g
    return {'tls': tls}


def test_term_count_formula_bounds_the_bessel_tail():
    """|J_k(z)| < 1e-14 for every k >= cheb_terms(z), over the range of z the kernels use."""
    for z in np.concatenate([np.logspace(-8, 0, 17), np.linspace(1, 1000, 120)]):
        K = cm.cheb_terms(z)
        assert np.abs(sp.jv(np.arange(K, K + 40), z)).max() < 1e-14, z


def test_zcap_fits_the_caps_of_the_kernels():
    for cap in (256, 512, 1024):
        z = cm.zcap_for(cap)
        assert cm.cheb_terms(z) <= cap < cm.cheb_terms(z + 1.0)


def test_miller_recurrence_matches_scipy():
    for z in (1e-6, 0.3, 7.5, 210.0, 900.0):
        K = cm.cheb_terms(z)       # the start index the kernels use
        got = cm.miller(K, z)
        assert np.abs(got - sp.jv(np.arange(K + 1), z)).max() < 2e-14, z


@pytest.mark.parametrize('n_defects,n_models', [(1, 12), (2, 6), (3, 3)])
def test_library_models_match_exact_and_need_fewer_applications_than_taylor(n_defects, n_models):
    times = np.linspace(0.0, 2.5, 200)
    for s in range(1000, 1000 + n_models):
        spec = lo.random_library_spec(np.random.default_rng(s), 1, n_defects)
        err, napp, info, (H, c_ops, rho0) = _run(spec, times)
        assert err < 1e-12, (s, err)
        assert info['steps'] == 1
        _, ntaylor = tm.evaluate(H, c_ops, rho0, lo.build_observables(spec, OBS3), times)
        assert napp < 0.6 * ntaylor, (s, napp, ntaylor)


def test_spread_bound_is_an_upper_bound_and_tighter_than_gershgorin():
    ratios = []
    for s in range(1000, 1040):
        H = lo.build_H(lo.random_library_spec(np.random.default_rng(s), 1, 3))
        ev = np.linalg.eigvalsh(H)
        b = cm.spread_bound(H)
        assert b >= (ev[-1] - ev[0]) * (1 - 1e-12)
        ratios.append(b / (ev[-1] - ev[0]))
    assert np.mean(ratios) < 1.15


EDGE = {
    'amplitude damping (non-normal dissipator)': (_spec(1, [0.3], [(0, 1, 2.0, [('sigmax', 'sigmax')])],
                                                        [(0, 'sigmam', 0.4), (1, 'sigmap', 0.3), (1, 'sigmaz', 0.2)]), np.linspace(0, 2.5, 120)),
    'pure dissipation, H = 0': (_spec(1, [0.0], [], [(0, 'sigmax', 0.4), (1, 'sigmam', 0.3)], qubit_energy=0.0), np.linspace(0, 5, 40)),
    'no dissipation': (_spec(2, [0.3, 0.2], [(0, 1, 3.0, [('sigmax', 'sigmax')]), (1, 2, 2.0, [('sigmay', 'sigmay')])], []), np.linspace(0, 2.5, 100)),
    'long span, several macro steps': (_spec(2, [0.3, 0.2], [(0, 1, 3.0, [('sigmax', 'sigmax')]), (1, 2, 2.0, [('sigmaz', 'sigmaz')])],
                                             [(0, 'sigmaz', 0.1), (2, 'sigmax', 0.05)]), np.linspace(0, 60, 25)),
    'non-uniform grid with repeats': (_spec(2, [0.3, 0.2], [(0, 2, 3.5, [('sigmax', 'sigmax')])], [(0, 'sigmax', 0.3), (1, 'sigmay', 0.4)]),
                                      np.array([0.5, 0.5, 0.51, 0.7, 0.7, 1.9, 2.0, 2.0001, 9.0])),
    'dissipation far above the library bounds': (_spec(1, [0.01], [], [(0, 'sigmax', 40.0), (0, 'sigmaz', 25.0), (1, 'sigmay', 30.0)]), np.linspace(0, 2.5, 60)),
    'diagonal H (Gershgorin exact) with transverse dissipators': (_spec(2, [0.4, 0.4], [(0, 1, 4.0, [('sigmaz', 'sigmaz')])],
                                                                          [(0, 'sigmax', 0.4), (1, 'sigmay', 0.4), (2, 'sigmax', 0.4)]), np.linspace(0, 2.5, 100)),
    'complex Hermitian H': (_spec(2, [0.3, 0.2], [(0, 1, 3.0, [('sigmax', 'sigmay')]), (1, 2, 2.0, [('sigmay', 'sigmax')])], [(0, 'sigmaz', 0.1)]), np.linspace(0, 2.5, 100)),
}


@pytest.mark.parametrize('name', list(EDGE))
def test_edge_cases_match_exact(name):
    spec, times = EDGE[name]
    err, napp, info, _ = _run(spec, times)
    assert err < 1e-12, (name, err, info)


def test_small_term_cap_splits_into_macro_steps_with_the_same_result():
    spec = lo.random_library_spec(np.random.default_rng(1003), 1, 2)
    times = np.linspace(0.0, 2.5, 80)
    e1, n1, i1, _ = _run(spec, times, kcap=640)
    e2, n2, i2, _ = _run(spec, times, kcap=48)
    assert i1['steps'] == 1 and i2['steps'] > 3
    assert e1 < 1e-12 and e2 < 1e-12
