"""
CPU oracle for the Lindblad model-evaluation hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference
arm may import this module.  The product package (`effective_model_learning_b200`)
never does; it fails loudly when its CUDA library is missing.

PARITY STATUS: **parity unpinned by the reference's own tests** -- the reference
ships no tests, golden vectors or data files and its arithmetic lives
in the un-vendored third-party `qutip` 4.7.5 (`qutip.mesolve`, reference
`basic_model.py:381-388`; pin at `README.txt:24`), which is not installable here.
What pins this oracle instead:
  * closed-form known answers (tests/test_oracle.py),
  * two independent evaluators here that must agree (dense expm vs tight ZVODE),
  * operator construction (H, c_ops, rho0, e_ops) compared bit-for-bit against the
    REAL reference code imported under `oracle/qutip_shim` (tests/golden/make_golden.py;
    fixtures committed in tests/golden/).

What is restated, and from where (all paths relative to /root/reference):
  OPS                      definitions.py:42-61  (qutip conventions for the 2x2 table)
  spec_from_model          two_level_system.py:38-58 (the record being read)
  build_c_ops              basic_model.py:152-189
  build_H                  basic_model.py:193-255   (coupling strength on BOTH factors -> s^2)
  build_rho0               basic_model.py:259-281
  build_observables        basic_model.py:351-375   (first qubit, 'exc' default)
  liouvillian              qutip 4.7.5 `liouvillian()` semantics (column stacking):
                           L = -i(I (x) H - H^T (x) I) + sum_c [conj(c) (x) c
                               - 1/2 I (x) c^dag c - 1/2 (c^dag c)^T (x) I]
                           NOT basic_model.py:285-310, which is wrong.
  expect_exact             the exact solution mesolve approximates:
                           rho(t_k) = exp(L (t_k - t_0)) rho0, expect = Tr(O rho)
  expect_zvode             scipy ZVODE, the integrator qutip 4.7.5 drives
  expect_qutip_emulation   same at qutip's default Options (adams, order 12,
                           atol 1e-8, rtol 1e-6, nsteps=1e9 as basic_model.py:386 sets)
  total_dev                learning_chain.py:714-742
"""

from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.integrate


# --------------------------------------------------------------------------------------
# a1: operator table (definitions.py:42-61).  qutip: basis(2,0)=|0>, sigmap=|0><1|,
# sigmam=|1><0|; ket([1]) = basis(2,1).
def _unit_dm(vec):
    # qutip: ket2dm(ket.unit()) -- normalise the ket first, then the outer product
    v = np.array(vec, dtype=complex).reshape(-1, 1)
    v = v / np.linalg.norm(v)
    return v @ v.conj().T


OPS = {
    'sigmaz': np.array([[1, 0], [0, -1]], dtype=complex),
    'sigmax': np.array([[0, 1], [1, 0]], dtype=complex),
    'sigmay': np.array([[0, -1j], [1j, 0]], dtype=complex),
    'sigmap': np.array([[0, 1], [0, 0]], dtype=complex),
    'sigmam': np.array([[0, 0], [1, 0]], dtype=complex),
    'id2': np.eye(2, dtype=complex),
    'gnd': np.array([[0, 0], [0, 1]], dtype=complex),      # sigmam*sigmap
    'exc': np.array([[1, 0], [0, 0]], dtype=complex),      # sigmap*sigmam
    'plus': _unit_dm([1, 1]),          # ket2dm((|1> + |0>).unit())
    '2e+1g': _unit_dm([1, 2]),         # ket2dm((2|1> + |0>).unit())
    'custom': np.array([[0.5, 0.5], [0.5, 0.5]], dtype=complex),
    'mm': np.eye(2, dtype=complex) / 2,
}


# --------------------------------------------------------------------------------------
def spec_from_model(model) -> dict:
    """
    Plain-data description of any object that looks like the reference's BasicModel
    (`.TLSs` of records with is_qubit/energy/couplings/Ls/initial_state,
    two_level_system.py:38-58).  Couplings are stored per holder exactly as found:
    {partner: [(strength, [(op_holder, op_partner), ...])]}.
    """
    tls = []
    for t in model.TLSs:
        init = t.initial_state
        if init is None or (isinstance(init, bool) and not init):
            init_arr = None
        else:
            init_arr = np.array(init.full() if hasattr(init, 'full') else init, dtype=complex)
        couplings = []
        for partner, lst in t.couplings.items():
            j = model.TLSs.index(partner)
            for strength, op_pairs in lst:
                couplings.append((j, float(strength), [tuple(p) for p in op_pairs]))
        tls.append({'is_qubit': bool(t.is_qubit), 'energy': t.energy,
                    'Ls': dict(t.Ls), 'initial_state': init_arr, 'couplings': couplings})
    return {'tls': tls}


def _site_op(n: int, k: int, op: np.ndarray) -> np.ndarray:
    """op on site k (site 0 most significant), identity elsewhere: qutip.tensor order."""
    out = np.array([[1.0 + 0j]])
    for s in range(n):
        out = np.kron(out, op if s == k else OPS['id2'])
    return out


def build_c_ops(spec: dict) -> list[np.ndarray]:
    """basic_model.py:152-189: sqrt(rate) * op_k, TLS order then dict order."""
    n = len(spec['tls'])
    out = []
    for k, t in enumerate(spec['tls']):
        for name, rate in t['Ls'].items():
            out.append(_site_op(n, k, OPS[name] * np.sqrt(rate)))
    return out


def build_H(spec: dict) -> np.ndarray:
    """basic_model.py:193-255: sum_k E_k sigmaz_k + sum s*A_i (x) s*B_j (s on both factors)."""
    n = len(spec['tls'])
    d = 2 ** n
    H = np.zeros((d, d), dtype=complex)
    for k, t in enumerate(spec['tls']):
        H = H + _site_op(n, k, OPS['sigmaz'] * t['energy'])
    for i, t in enumerate(spec['tls']):
        for (j, strength, op_pairs) in t['couplings']:
            for (op_i, op_j) in op_pairs:
                term = np.array([[1.0 + 0j]])
                for s in range(n):
                    if s == i:
                        term = np.kron(term, strength * OPS[op_i])
                    elif s == j:
                        term = np.kron(term, strength * OPS[op_j])
                    else:
                        term = np.kron(term, OPS['id2'])
                H = H + term
    return H


def build_rho0(spec: dict) -> np.ndarray:
    """basic_model.py:259-281: product state; default qubit 'plus', defect 'gnd'."""
    out = np.array([[1.0 + 0j]])
    for t in spec['tls']:
        if t['initial_state'] is not None:
            out = np.kron(out, t['initial_state'])
        elif t['is_qubit']:
            out = np.kron(out, OPS['plus'])
        else:
            out = np.kron(out, OPS['gnd'])
    return out


def normalise_observables(observable_ops):
    """basic_model.py:351-360."""
    if isinstance(observable_ops, bool) and not observable_ops:
        observable_ops = 'exc'
    if isinstance(observable_ops, str):
        observable_ops = [observable_ops]
    if not (isinstance(observable_ops, list) and all(isinstance(x, str) for x in observable_ops)):
        raise RuntimeError('do not understand oservable operators input for dynamics calculation')
    return observable_ops


def build_observables(spec: dict, observable_ops) -> list[np.ndarray]:
    """basic_model.py:365-375: op on the FIRST qubit in TLS order, id2 elsewhere."""
    names = normalise_observables(observable_ops)
    n = len(spec['tls'])
    first_q = next((k for k, t in enumerate(spec['tls']) if t['is_qubit']), None)
    out = []
    for name in names:
        if first_q is None:
            out.append(np.eye(2 ** n, dtype=complex))
        else:
            out.append(_site_op(n, first_q, OPS[name]))
    return out


# --------------------------------------------------------------------------------------
def liouvillian(H: np.ndarray, c_ops: list[np.ndarray]) -> np.ndarray:
    """Column-stacking superoperator of the Lindblad generator (qutip `liouvillian`)."""
    d = H.shape[0]
    I = np.eye(d, dtype=complex)
    L = -1j * (np.kron(I, H) - np.kron(H.T, I))
    for c in c_ops:
        cdc = c.conj().T @ c
        L = L + np.kron(c.conj(), c) - 0.5 * np.kron(I, cdc) - 0.5 * np.kron(cdc.T, I)
    return L


def liouvillian_csr(H: np.ndarray, c_ops: list[np.ndarray]):
    """Same superoperator assembled sparse (CSR), the way qutip's `liouvillian` does it."""
    sp = scipy.sparse
    d = H.shape[0]
    I = sp.identity(d, dtype=complex, format='csr')
    Hs = sp.csr_matrix(H)
    L = -1j * (sp.kron(I, Hs, format='csr') - sp.kron(Hs.T, I, format='csr'))
    for c in c_ops:
        cs = sp.csr_matrix(c)
        cdc = cs.conj().T @ cs
        L = L + sp.kron(cs.conj(), cs, format='csr') - 0.5 * sp.kron(I, cdc, format='csr') - 0.5 * sp.kron(cdc.T, I, format='csr')
    return sp.csr_matrix(L)


def _is_herm(a: np.ndarray) -> bool:
    return bool(np.allclose(a, a.conj().T, rtol=0, atol=1e-12))


def _finish(expect: np.ndarray, obs: list[np.ndarray], rho0: np.ndarray) -> list[np.ndarray]:
    """float64 when O and rho0 are Hermitian, complex128 otherwise (qutip mesolve)."""
    herm0 = _is_herm(rho0)
    out = []
    for m, O in enumerate(obs):
        if herm0 and _is_herm(O):
            out.append(np.ascontiguousarray(expect[m].real))
        else:
            out.append(np.ascontiguousarray(expect[m]))
    return out


def _trace_rows(obs: list[np.ndarray]) -> np.ndarray:
    """Row vectors w_m with Tr(O_m rho) = w_m . vec_F(rho):  w = vec_F(O^T)."""
    return np.stack([O.T.ravel(order='F') for O in obs])


def expect_exact(spec: dict, times, observable_ops=False, *, dt_rtol: float = 1e-12):
    """
    Exact evaluator: rho(t_k) = exp(L (t_k - t_0)) rho0 by stepping with one dense
    scipy.linalg.expm per distinct interval length.  Returns list of K arrays (T,).
    """
    times = np.asarray(times, dtype=float)
    H = build_H(spec)
    c_ops = build_c_ops(spec)
    rho0 = build_rho0(spec)
    obs = build_observables(spec, observable_ops)
    L = liouvillian(H, c_ops)
    W = _trace_rows(obs)
    v = rho0.ravel(order='F').astype(complex)
    out = np.zeros((len(obs), len(times)), dtype=complex)
    props: list[tuple[float, np.ndarray]] = []
    for k in range(len(times)):
        if k > 0:
            dt = times[k] - times[k - 1]
            P = None
            for (dt0, P0) in props:
                if abs(dt - dt0) <= dt_rtol * max(abs(dt), abs(dt0)):
                    P = P0
                    break
            if P is None:
                P = scipy.linalg.expm(L * dt)
                props.append((dt, P))
            v = P @ v
        out[:, k] = W @ v
    return _finish(out, obs, rho0)


def expect_zvode(spec: dict, times, observable_ops=False, *, atol=1e-13, rtol=1e-12,
                 method='adams', order=12, nsteps=10 ** 9, sparse=True):
    """
    Independent evaluator: ZVODE on dv/dt = L v from t_0 = times[0], one integrate
    call per output time -- the loop qutip 4.7.5 mesolve runs.
    """
    times = np.asarray(times, dtype=float)
    H = build_H(spec)
    c_ops = build_c_ops(spec)
    rho0 = build_rho0(spec)
    obs = build_observables(spec, observable_ops)
    L = liouvillian_csr(H, c_ops) if sparse else liouvillian(H, c_ops)
    W = _trace_rows(obs)

    def rhs(t, y):
        return L @ y

    r = scipy.integrate.ode(rhs)
    r.set_integrator('zvode', method=method, order=order, atol=atol, rtol=rtol, nsteps=nsteps)
    r.set_initial_value(rho0.ravel(order='F').astype(complex), times[0])
    out = np.zeros((len(obs), len(times)), dtype=complex)
    for k in range(len(times)):
        if k > 0:
            r.integrate(times[k])
            if not r.successful():
                raise Exception('ODE integration error: Try to increase the allowed number '
                                'of substeps by increasing the nsteps parameter in the Options class.')
        out[:, k] = W @ r.y
    return _finish(out, obs, rho0)


def expect_qutip_emulation(spec: dict, times, observable_ops=False):
    """What qutip 4.7.5 mesolve would print: ZVODE adams/12, atol 1e-8, rtol 1e-6, CSR RHS."""
    return expect_zvode(spec, times, observable_ops, atol=1e-8, rtol=1e-6,
                        method='adams', order=12, nsteps=10 ** 9, sparse=True)


def total_dev(model_datasets, target_datasets, n_times: int) -> float:
    """learning_chain.py:738-742."""
    total = 0.0
    for i in range(len(model_datasets)):
        total += np.sum(np.square(abs(model_datasets[i] - target_datasets[i])))
    return total / (n_times * len(model_datasets))


# --------------------------------------------------------------------------------------
# Synthetic workload of SURVEY.md section 8(d) -- shared by tests and bench so that the CPU
# baseline and the GPU arm see identical inputs.

PAULIS = ('sigmax', 'sigmay', 'sigmaz')
BOUNDS = {'couplings': (0.1, 4.0), 'energies': (0.01, 0.4), 'Ls': (0.01, 0.4)}


def random_library_spec(rng: np.random.Generator, n_qubits: int, n_defects: int,
                        p_present: float = 0.5) -> dict:
    """
    Random model inside the configs.py bounds over the full ("EVERYTHING", configs.py:467-496)
    library: defect energies U(0.01,0.4); each library process present w.p. p_present;
    couplings U(0.1,4); rates U(0.01,0.4).  Production initial states: qubits 'plus',
    defects 'mm' (execute_learning_full.py:151-164).  Qubit energy 1 (ibid :161).
    """
    tls = []
    for q in range(n_qubits):
        tls.append({'is_qubit': True, 'energy': 1.0, 'Ls': {}, 'couplings': [],
                    'initial_state': OPS['plus'].copy()})
    for v in range(n_defects):
        tls.append({'is_qubit': False, 'energy': float(rng.uniform(*BOUNDS['energies'])),
                    'Ls': {}, 'couplings': [], 'initial_state': OPS['mm'].copy()})
    Q = n_qubits
    # qubit-defect couplings, pair-major then operator (vectorise_under_library order)
    for q in range(Q):
        for v in range(n_defects):
            for op in PAULIS:
                if rng.uniform() < p_present:
                    tls[q]['couplings'].append((Q + v, float(rng.uniform(*BOUNDS['couplings'])), [(op, op)]))
    for v1 in range(n_defects):
        for v2 in range(v1 + 1, n_defects):
            for op in PAULIS:
                if rng.uniform() < p_present:
                    tls[Q + v1]['couplings'].append((Q + v2, float(rng.uniform(*BOUNDS['couplings'])), [(op, op)]))
    for k in range(Q + n_defects):
        for op in PAULIS:
            if rng.uniform() < p_present:
                tls[k]['Ls'][op] = float(rng.uniform(*BOUNDS['Ls']))
    return {'tls': tls}
