"""CPU tests that pin the oracle: closed forms, two independent evaluators, and the reference's own operators."""
import json
import os

import numpy as np
import pytest

from oracle import lindblad_oracle as lo
from oracle import taylor_action_model as tm
from tests.helpers import OBS3

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def spec_from_json(js):
    tls = []
    for t in js['tls']:
        init = t['initial_state']
        tls.append({'is_qubit': t['is_qubit'], 'energy': t['energy'], 'Ls': dict(t['Ls']),
                    'initial_state': None if init is None else np.array([[complex(*z) for z in row] for row in init]),
                    'couplings': [(j, s, [tuple(p) for p in pairs]) for (j, s, pairs) in t['couplings']]})
    return {'tls': tls}


def qubit(energy=1.0, Ls=None, init=None):
    return {'is_qubit': True, 'energy': energy, 'Ls': Ls or {}, 'initial_state': init, 'couplings': []}


def defect(energy=0.1, Ls=None, init=None, couplings=None):
    return {'is_qubit': False, 'energy': energy, 'Ls': Ls or {}, 'initial_state': init, 'couplings': couplings or []}


def test_closed_form_free_precession_and_dephasing():
    # SURVEY.md section 4 (iii),(iv): H = E sigmaz, 'plus' state; sigmaz dissipator gamma
    ts = np.linspace(0.3, 4.0, 57)          # initial state sits at ts[0], not at 0
    E, g = 0.8, 0.07
    ex = lo.expect_exact({'tls': [qubit(E)]}, ts, OBS3)
    np.testing.assert_allclose(ex[0], np.cos(2 * E * (ts - ts[0])), atol=1e-12)
    np.testing.assert_allclose(ex[1], np.sin(2 * E * (ts - ts[0])), atol=1e-12)
    np.testing.assert_allclose(ex[2], 0, atol=1e-12)
    ex = lo.expect_exact({'tls': [qubit(E, {'sigmaz': g})]}, ts, OBS3)
    np.testing.assert_allclose(ex[0], np.exp(-2 * g * (ts - ts[0])) * np.cos(2 * E * (ts - ts[0])), atol=1e-12)


def test_closed_form_xx_coupling_to_maximally_mixed_defect():
    # H = E sz (x) 1 + s^2 sx (x) sx with the defect maximally mixed and at zero energy: the defect's sx is
    # conserved (+-1 with prob 1/2 each), so the qubit precesses about (+-s^2, 0, E):
    # <sz>(t) = -(E s^2 / W^2)(1 - cos 2Wt) for sx=+1 and +(...) for sx=-1 -> averages to 0; <sx>(t) = 1 - (E^2/W^2)(1 - cos 2Wt)
    E, s = 1.0, 1.3
    W = np.hypot(E, s * s)
    ts = np.linspace(0, 3, 80)
    spec = {'tls': [dict(qubit(E), couplings=[(1, s, [('sigmax', 'sigmax')])]), defect(0.0, init=lo.OPS['mm'])]}
    ex = lo.expect_exact(spec, ts, OBS3)
    np.testing.assert_allclose(ex[0], 1 - (E * E / W ** 2) * (1 - np.cos(2 * W * ts)), atol=1e-12)
    np.testing.assert_allclose(ex[2], 0, atol=1e-12)


def test_coupling_strength_enters_squared():
    # SURVEY.md F7 (basic_model.py:245,247)
    spec = {'tls': [dict(qubit(), couplings=[(1, 1.7, [('sigmax', 'sigmay')])]), defect()]}
    H = lo.build_H(spec)
    H0 = lo.build_H({'tls': [qubit(), defect()]})
    np.testing.assert_allclose(H - H0, 1.7 ** 2 * np.kron(lo.OPS['sigmax'], lo.OPS['sigmay']), atol=1e-15)


@pytest.mark.parametrize('QD', [(1, 1), (1, 2), (1, 3)])
def test_expm_and_tight_zvode_agree(QD):
    ts = np.linspace(0, 2.5, 40)
    spec = lo.random_library_spec(np.random.default_rng(42), *QD)
    a = lo.expect_exact(spec, ts, OBS3)
    b = lo.expect_zvode(spec, ts, OBS3)
    assert max(np.abs(x - y).max() for x, y in zip(a, b)) < 2e-9


def test_qutip_emulation_is_only_1e_minus_5_accurate():
    # the reference's own integrator tolerance
    ts = np.linspace(0, 2.5, 200)
    spec = lo.random_library_spec(np.random.default_rng(1001), 1, 2)
    a = lo.expect_exact(spec, ts, OBS3)
    b = lo.expect_qutip_emulation(spec, ts, OBS3)
    err = max(np.abs(x - y).max() for x, y in zip(a, b))
    assert 1e-9 < err < 1e-4


def test_trace_and_hermiticity_preserved():
    spec = lo.random_library_spec(np.random.default_rng(5), 1, 2)
    L = lo.liouvillian(lo.build_H(spec), lo.build_c_ops(spec))
    d = 8
    tr_row = np.eye(d).ravel(order='F')
    np.testing.assert_allclose(tr_row @ L, 0, atol=1e-13)          # d/dt Tr(rho) = 0
    import scipy.linalg
    rho = (scipy.linalg.expm(L * 0.7) @ lo.build_rho0(spec).ravel(order='F')).reshape(d, d, order='F')
    np.testing.assert_allclose(rho, rho.conj().T, atol=1e-13)
    assert abs(np.trace(rho) - 1) < 1e-13


def test_operators_match_what_the_reference_builds():
    """H, c_ops, rho0, e_ops and observables recorded from the REAL reference code (make_golden.py)."""
    z = np.load(os.path.join(GOLD, 'reference_operators.npz'))
    meta = json.load(open(os.path.join(GOLD, 'reference_operators_specs.json')))
    ts = np.array(meta['times'])
    for i, js in enumerate(meta['specs']):
        spec = spec_from_json(js)
        assert np.array_equal(lo.build_H(spec), z[f'H_{i}']), i
        c = lo.build_c_ops(spec)
        assert len(c) == z[f'Ls_{i}'].shape[0]
        for a, b in zip(c, z[f'Ls_{i}']):
            assert np.array_equal(a, b)
        assert np.array_equal(lo.build_rho0(spec), z[f'rho0_{i}'])
        for a, b in zip(lo.build_observables(spec, OBS3), z[f'obs_{i}']):
            assert np.array_equal(a, b)
        got = np.array([np.asarray(e, dtype=complex) for e in lo.expect_exact(spec, ts, OBS3)])
        np.testing.assert_allclose(got, z[f'expect_{i}'], atol=1e-13, rtol=0)


def test_frozen_observables_reproduce():
    z = np.load(os.path.join(GOLD, 'oracle_observables.npz'))
    ts = z['times']
    for (Q, D) in ((1, 1), (1, 2)):
        for n in (0, 17, 49):
            spec = lo.random_library_spec(np.random.default_rng(int(z[f'seeds_Q{Q}_D{D}'][n])), Q, D)
            np.testing.assert_allclose(np.array(lo.expect_exact(spec, ts, OBS3)), z[f'expect_Q{Q}_D{D}'][n], atol=1e-12)


def test_total_dev_formula():
    a = [np.array([1.0, 2.0, 3.0]), np.array([0.0, 0.0, 1j])]
    b = [np.array([1.0, 1.0, 1.0]), np.array([0.0, 0.0, 0.0])]
    assert lo.total_dev(a, b, 3) == pytest.approx((0 + 1 + 4 + 1) / 6)


@pytest.mark.parametrize('QD', [(1, 1), (1, 2), (1, 3)])
def test_kernel_algorithm_model_matches_exact(QD):
    """The Taylor-action/dense-output algorithm the CUDA kernels run, with their constants, on the CPU."""
    ts = np.linspace(0, 2.5, 200)
    spec = lo.random_library_spec(np.random.default_rng(1003), *QD)
    H, cs, r0 = lo.build_H(spec), lo.build_c_ops(spec), lo.build_rho0(spec)
    obs = lo.build_observables(spec, OBS3)
    out, napp = tm.evaluate(H, cs, r0, obs, ts, theta_target=16.0, tol=1e-11, q_max=16, m_cap=80)
    ex = np.array(lo.expect_exact(spec, ts, OBS3))
    assert np.abs(out.real - ex).max() < 1e-10
    assert napp < 200 * 10
