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


# ---------------------------------------------------------------------------------------------
# CPU stand-in for the device, for host-logic tests ONLY (tests may use the oracle; the product
# never does).  It receives exactly what the C ABI would receive -- the context's library and one
# parameter row per model -- so it also checks the marshalling.
def spec_from_context_row(ctx, row):
    tls = [{'is_qubit': (k == ctx.obs_site), 'energy': 0.0, 'Ls': {}, 'couplings': [],
            'initial_state': np.array(ctx.site_states[k])} for k in range(ctx.n_tls)]
    for key, col in ctx.columns.items():
        v = float(row[col])
        if v == 0.0:
            continue
        if key[0] == 'E':
            tls[key[1]]['energy'] = v
        elif key[0] == 'C':
            tls[key[1]]['couplings'].append((key[2], v, list(key[3])))
        else:
            tls[key[1]]['Ls'][key[2]] = v
    return {'tls': tls}


def oracle_evaluate(ctx, params, target_set=None, want_expect=True, want_loss=False):
    params = np.atleast_2d(params)
    B, K, T = params.shape[0], len(ctx.obs_names), len(ctx.times)
    expect = np.zeros((B, K, T), dtype=np.complex128)
    for b in range(B):
        ex = lo.expect_exact(spec_from_context_row(ctx, params[b]), ctx.times, list(ctx.obs_names))
        expect[b] = np.array([np.asarray(e, dtype=complex) for e in ex])
    loss = None
    if want_loss:
        tg = ctx.targets if ctx.targets.ndim == 3 else ctx.targets[None]
        sets = np.zeros(B, dtype=int) if target_set is None else np.asarray(target_set)
        loss = np.array([np.sum(np.abs(expect[b] - tg[sets[b]]) ** 2) / (T * K) for b in range(B)])
    return (expect if want_expect else None), loss


def random_general_spec(rng, n, p_coupling=0.5):
    """
    Random model exercising everything the DMMA kernels support beyond the config libraries: any qubit position,
    all initial states, generalized-permutation jump operators, Hermitian couplings with mixed Pauli pairs
    (sigmax (x) sigmay makes H complex), repeated couplings on a pair.
    """
    paulis = ['sigmax', 'sigmay', 'sigmaz']
    jumps = ['sigmax', 'sigmay', 'sigmaz', 'sigmap', 'sigmam', 'exc', 'gnd']
    states = [None, lo.OPS['plus'], lo.OPS['mm'], lo.OPS['2e+1g'], lo.OPS['gnd'], lo.OPS['exc']]
    qubit_at = int(rng.integers(0, n))
    tls = []
    for k in range(n):
        init = states[int(rng.integers(0, len(states)))]
        Ls = {str(op): float(rng.uniform(0.01, 0.4)) for op in rng.permutation(jumps)[:int(rng.integers(0, 4))]}
        tls.append({'is_qubit': k == qubit_at or bool(rng.uniform() < 0.2), 'energy': float(rng.uniform(0.05, 1.2)),
                    'Ls': Ls, 'initial_state': None if init is None else init.copy(), 'couplings': []})
    for i in range(n):
        for j in range(n):
            if i != j and rng.uniform() < p_coupling / max(1, n - 2):
                pairs = [(str(rng.choice(paulis)), str(rng.choice(paulis))) for _ in range(int(rng.integers(1, 3)))]
                tls[i]['couplings'].append((j, float(rng.uniform(0.1, 2.5)), pairs))
    return {'tls': tls}
