"""Shared helpers for the tests: building package models from oracle specs and comparing results."""
import numpy as np

from oracle import lindblad_oracle as lo

OBS3 = ['sigmax', 'sigmay', 'sigmaz']


def model_from_spec(spec, cls=None):
    """Build an effective_model_learning_b200 model from an oracle spec (plain dict)."""
    from effective_model_learning_b200 import basic_model
    cls = cls or basic_model.BasicModel
    m = cls()
    made = []
    for t in spec['tls']:
        made.append(m.add_TLS(is_qubit=t['is_qubit'], energy=t['energy'], couplings={}, Ls=dict(t['Ls']),
                              initial_state=(t['initial_state'] if t['initial_state'] is not None else False)))
    for i, t in enumerate(spec['tls']):
        for (j, strength, op_pairs) in t['couplings']:
            made[i].couplings.setdefault(made[j], []).append((strength, list(op_pairs)))
    m.build_operators()
    return m


def assert_close(got, want, tol=1e-8, what=''):
    """|a-b| <= tol * max(1, max|b|) on every observable (relative with a floor)."""
    assert len(got) == len(want)
    for m, (a, b) in enumerate(zip(got, want)):
        a = np.asarray(a)
        b = np.asarray(b)
        assert a.shape == b.shape, (what, m, a.shape, b.shape)
        err = np.abs(a - b).max() if a.size else 0.0
        lim = tol * max(1.0, np.abs(b).max() if b.size else 0.0)
        assert err <= lim, f'{what} observable {m}: max abs err {err:.3e} > {lim:.3e}'
    return True


def max_err(got, want):
    return max(float(np.abs(np.asarray(a) - np.asarray(b)).max()) for a, b in zip(got, want))
