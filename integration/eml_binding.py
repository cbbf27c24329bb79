"""
ctypes binding a maintainer of the reference (henry104413/effective_model_learning) drops next to basic_model.py
to replace the one third-party call of the hot path,

    qutip.mesolve(self.H, self.initial_DM, evaluation_times, c_ops=self.Ls, e_ops=observable_ops_full, options=...)
                                                                                        (reference basic_model.py:381-388)
by libeml_b200.so (C ABI: include/eml_b200.h):

    # basic_model.py, replacing lines 381-393
    import eml_binding
    returnable = eml_binding.expect(self, evaluation_times, observable_ops)      # list of K arrays of shape (T,)
    return returnable if not custom_function_on_return else custom_function_on_return(returnable)

It works on the REFERENCE'S OWN objects (BasicModel holding TwoLevelSystem records and qutip Qobj operators) and needs
nothing of the effective_model_learning_b200 Python package -- only the shared library.  The same functions accept this
repository's numpy-based classes (duck typing: .TLSs, .initial_DM, TLS.energy / .couplings / .Ls / .is_qubit), which is how
the GPU box, where the reference and qutip do not exist, exercises exactly this code path (tests/test_gpu_binding.py).

    marshal(model, times, observable_ops) -> dict     what crosses the C ABI, as plain numpy arrays (no GPU needed)
    evaluate(desc)                        -> complex array [K][T] through eml_plan_create / eml_eval_batch_host
    expect(model, times, observable_ops)  -> what mesolve(...).expect returned
"""
import ctypes as C
import os

import numpy as np

# the operator names of the reference's table (definitions.py:42-61), in a fixed order = rows of op_table
OPS = ['sigmaz', 'sigmax', 'sigmay', 'sigmap', 'sigmam', 'id2', 'gnd', 'exc', 'plus', '2e+1g', 'custom', 'mm']
TERM_ENERGY, TERM_COUPLING, TERM_LINDBLAD = 0, 1, 2


class EmlTerm(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('param', 'kind', 'site_a', 'site_b', 'op_a', 'op_b')]


class EmlProblemDesc(C.Structure):
    _fields_ = [('n_tls', C.c_int32), ('n_params', C.c_int32), ('n_terms', C.c_int32), ('terms', C.POINTER(EmlTerm)),
                ('n_ops', C.c_int32), ('op_table', C.POINTER(C.c_double)), ('rho0', C.POINTER(C.c_double)),
                ('obs_site', C.c_int32), ('n_obs', C.c_int32), ('obs', C.POINTER(C.c_double)),
                ('n_times', C.c_int32), ('times', C.POINTER(C.c_double)),
                ('n_target_sets', C.c_int32), ('targets', C.POINTER(C.c_double))]


_dp = C.POINTER(C.c_double)
_lib = None


def library():
    """libeml_b200.so: $EML_B200_LIB, else next to this file, else the in-tree build of this repository."""
    global _lib
    if _lib is None:
        here = os.path.dirname(os.path.abspath(__file__))
        candidates = [os.environ.get('EML_B200_LIB'), os.path.join(here, 'libeml_b200.so'),
                      os.path.join(here, '..', 'effective_model_learning_b200', 'libeml_b200.so')]
        path = next((p for p in candidates if p and os.path.exists(p)), None)
        if path is None:
            raise RuntimeError('libeml_b200.so not found (set EML_B200_LIB)')
        lib = C.CDLL(path)
        lib.eml_plan_create.argtypes = [C.POINTER(EmlProblemDesc), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.eml_plan_destroy.argtypes = [C.c_void_p]
        lib.eml_eval_batch_host.argtypes = [C.c_void_p, C.c_int64, _dp, C.POINTER(C.c_int32), _dp, _dp]
        lib.eml_last_error_string.restype = C.c_char_p
        _lib = lib
    return _lib


def _dense(x) -> np.ndarray:
    """qutip.Qobj -> ndarray (Qobj.full()); ndarrays pass through."""
    return np.asarray(x.full() if hasattr(x, 'full') else x, dtype=np.complex128)


def _ops_table(ops=None):
    if ops is None:
        import definitions          # the reference's own module (definitions.py:42-61)
        ops = definitions.ops
    return ops


def marshal(model, times, observable_ops, ops=None) -> dict:
    """
    One model -> the problem descriptor of include/eml_b200.h.  One parameter column per process, in the order the
    reference builds its operators: energies (basic_model.py:206-221), couplings as stored on their holder with the
    strength entering SQUARED (:226-251; the kernel squares EML_TERM_COUPLING columns), Lindblad rates (:152-189).
    """
    ops = _ops_table(ops)
    tls = list(model.TLSs)
    terms, params = [], []

    def col(v):
        params.append(float(v))
        return len(params) - 1

    for k, t in enumerate(tls):
        terms.append((col(t.energy), TERM_ENERGY, k, -1, OPS.index('sigmaz'), -1))
    for k, t in enumerate(tls):
        for partner, entries in t.couplings.items():
            for strength, pairs in entries:
                c = col(strength)
                terms += [(c, TERM_COUPLING, k, tls.index(partner), OPS.index(a), OPS.index(b)) for a, b in pairs]
    for k, t in enumerate(tls):
        terms += [(col(rate), TERM_LINDBLAD, k, -1, OPS.index(op), -1) for op, rate in t.Ls.items()]
    # the observable acts on the FIRST qubit (basic_model.py:365-375); without a qubit it degenerates to the identity
    qubits = [k for k, t in enumerate(tls) if t.is_qubit]
    names = list(observable_ops)
    return {'n_tls': len(tls),
            'terms': np.array(terms, dtype=np.int32).reshape(-1, 6),
            'params': np.array(params, dtype=np.float64),
            'op_table': np.stack([_dense(ops[n]) for n in OPS]),
            'rho0': np.ascontiguousarray(_dense(model.initial_DM)),
            'obs_site': qubits[0] if qubits else 0,
            'obs': np.stack([_dense(ops[n if qubits else 'id2']) for n in names]),
            'times': np.ascontiguousarray(times, dtype=np.float64)}


def evaluate(desc: dict) -> np.ndarray:
    """expect[K][T] (complex) of one marshalled model: eml_plan_create -> eml_eval_batch_host -> eml_plan_destroy."""
    lib = library()
    terms = np.ascontiguousarray(desc['terms'], dtype=np.int32)
    tarr = (EmlTerm * len(terms))(*[EmlTerm(*map(int, t)) for t in terms])
    op_table = np.ascontiguousarray(desc['op_table'], dtype=np.complex128)
    rho0 = np.ascontiguousarray(desc['rho0'], dtype=np.complex128)
    obs = np.ascontiguousarray(desc['obs'], dtype=np.complex128)
    times = np.ascontiguousarray(desc['times'], dtype=np.float64)
    params = np.ascontiguousarray(desc['params'], dtype=np.float64)[None]
    d = EmlProblemDesc(int(desc['n_tls']), params.shape[1], len(terms), tarr, len(op_table),
                       op_table.ctypes.data_as(_dp), rho0.ctypes.data_as(_dp), int(desc['obs_site']), len(obs),
                       obs.ctypes.data_as(_dp), len(times), times.ctypes.data_as(_dp), 0, None)
    plan = C.c_void_p()
    if lib.eml_plan_create(C.byref(d), 0, None, C.byref(plan)) != 0:
        raise RuntimeError(lib.eml_last_error_string().decode())
    out = np.empty((1, len(obs), len(times)), dtype=np.complex128)
    rc = lib.eml_eval_batch_host(plan, 1, params.ctypes.data_as(_dp), None, out.ctypes.data_as(_dp), None)
    lib.eml_plan_destroy(plan)
    if rc != 0:
        raise RuntimeError(lib.eml_last_error_string().decode())
    return out[0]


def _is_hermitian(x) -> bool:
    a = _dense(x)
    return bool(np.allclose(a, a.conj().T, rtol=0.0, atol=1e-12))


def expect(model, times, observable_ops, ops=None):
    """What qutip.mesolve(...).expect returned at basic_model.py:389: K arrays (T,), float64 for Hermitian e_ops and rho0."""
    ops = _ops_table(ops)
    out = evaluate(marshal(model, times, observable_ops, ops))
    herm = _is_hermitian(model.initial_DM)
    return [out[m].real.copy() if herm and _is_hermitian(ops[n]) else out[m].copy()
            for m, n in enumerate(observable_ops)]
