"""
The reference-side binding (integration/eml_binding.py, the stub INTEGRATION.md tells a maintainer to add at
basic_model.py:381) on CPU: what it marshals from the REAL reference's objects is pinned by a committed fixture, and the
fixture describes exactly the operators the reference itself built (H, c_ops, rho0, e_ops of reference_operators.npz).
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, 'integration'))

import eml_binding  # noqa: E402
from effective_model_learning_b200 import definitions  # noqa: E402
from tests.helpers import OBS3, model_from_spec  # noqa: E402

GOLD = os.path.join(HERE, 'golden')


def _fixture():
    z = np.load(os.path.join(GOLD, 'reference_binding.npz'))
    n = sum(1 for k in z.files if k.startswith('terms_'))
    return [{'terms': z[f'terms_{i}'], 'params': z[f'params_{i}'], 'rho0': z[f'rho0_{i}'], 'obs': z[f'obs_{i}'],
             'times': z[f'times_{i}'], 'n_tls': int(z[f'shape_{i}'][0]), 'obs_site': int(z[f'shape_{i}'][1]),
             'op_table': z['op_table']} for i in range(n)]


def _specs():
    with open(os.path.join(GOLD, 'reference_operators_specs.json')) as f:
        g = json.load(f)
    specs = []
    for s in g['specs']:
        tls = []
        for t in s['tls']:
            init = None if t['initial_state'] is None else np.array([[complex(*zz) for zz in row] for row in t['initial_state']])
            tls.append({'is_qubit': t['is_qubit'], 'energy': t['energy'], 'Ls': t['Ls'], 'initial_state': init,
                        'couplings': [(j, s_, [tuple(p) for p in pairs]) for (j, s_, pairs) in t['couplings']]})
        specs.append({'tls': tls})
    return specs, np.array(g['times'])


def _same(a: dict, b: dict):
    assert a['n_tls'] == b['n_tls'] and a['obs_site'] == b['obs_site']
    for key in ('terms', 'params', 'rho0', 'obs', 'times', 'op_table'):
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key


def _kron_site(op, site, n):
    out = np.eye(1, dtype=complex)
    for k in range(n):
        out = np.kron(out, op if k == site else np.eye(2))
    return out


def test_descriptor_rebuilds_the_operators_the_reference_built():
    """descriptor -> H, c_ops, rho0, e_ops with nothing but numpy == the Qobj data of the real reference (bit for bit up to
    summation order), including the squared coupling strength (basic_model.py:245,247)."""
    ref = np.load(os.path.join(GOLD, 'reference_operators.npz'))
    for i, d in enumerate(_fixture()):
        n, ops = d['n_tls'], d['op_table']
        H = np.zeros((2 ** n, 2 ** n), dtype=complex)
        c_ops = []
        for (param, kind, sa, sb, oa, ob) in d['terms']:
            v = d['params'][param]
            if kind == eml_binding.TERM_ENERGY:
                H += v * _kron_site(ops[oa], sa, n)
            elif kind == eml_binding.TERM_COUPLING:
                H += v * v * (_kron_site(ops[oa], sa, n) @ _kron_site(ops[ob], sb, n))
            else:
                c_ops.append(np.sqrt(v) * _kron_site(ops[oa], sa, n))
        assert np.allclose(H, ref[f'H_{i}'], rtol=0, atol=1e-13), i
        want = ref[f'Ls_{i}']
        assert len(c_ops) == len(want)
        for a, b in zip(c_ops, want):
            assert np.allclose(a, b, rtol=0, atol=1e-15), i
        assert np.array_equal(d['rho0'], ref[f'rho0_{i}']), i
        for m in range(3):
            assert np.array_equal(_kron_site(d['obs'][m], d['obs_site'], n), ref[f'obs_{i}'][m]), i


def test_package_models_marshal_exactly_like_the_reference_models():
    """The numpy-based classes of this package are duck-type compatible: same descriptor as from the reference's objects."""
    specs, ts = _specs()
    fx = _fixture()
    assert len(specs) == len(fx) == 24
    for spec, want in zip(specs, fx):
        _same(eml_binding.marshal(model_from_spec(spec), ts, OBS3, definitions.ops), want)


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='the reference exists only in the authoring container')
def test_fixture_is_what_the_real_reference_objects_marshal_to():
    import subprocess
    code = ('import sys; sys.path.insert(0, %r); import numpy as np, make_binding_golden as g; d = g.reference_descriptors(); '
            'np.savez(sys.argv[1], **{f"{k}_{i}": np.asarray(v) for i, x in enumerate(d) for k, v in x.items()})' % GOLD)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, 'fresh.npz')
        env = dict(os.environ, PYTHONHASHSEED='0')
        subprocess.run([sys.executable, '-c', code, out], check=True, env=env, cwd=ROOT)
        z = np.load(out)
        for i, want in enumerate(_fixture()):
            got = {k: z[f'{k}_{i}'] for k in ('terms', 'params', 'rho0', 'obs', 'times', 'op_table')}
            got['n_tls'], got['obs_site'] = int(z[f'n_tls_{i}']), int(z[f'obs_site_{i}'])
            _same(got, want)


def test_binding_needs_the_library_and_a_device():
    """No CPU fallback on this path either: without a GPU the C ABI reports EML_ERR_NO_DEVICE and the stub raises."""
    from effective_model_learning_b200 import _native
    if _native.device_count() > 0:
        pytest.skip('a CUDA device is present')
    with pytest.raises(RuntimeError):
        eml_binding.evaluate(_fixture()[0])
