"""
Minimal stand-in for the parts of `qutip` 4.7.5 that the reference imports
(definitions.py:36,43-60; basic_model.py:381-388).  TEST INFRASTRUCTURE ONLY.

Purpose: qutip cannot be installed in the authoring container, so the reference cannot be
imported.  Putting this directory on sys.path lets `/root/reference`'s own modules
(definitions, basic_model, learning_model, learning_chain, process/params handlers) run
UNMODIFIED here, so that tests/golden/make_golden.py can record what the reference's real
code builds (H, c_ops, rho0, e_ops, RNG consumption, accept/reject logic).  The one thing
that is NOT the reference is `mesolve` itself: it is delegated to oracle.lindblad_oracle
(exact expm by default, or the ZVODE emulation of qutip's default tolerances).

Nothing in the product package or in GPU tests / bench imports this.
"""
import numpy as np

MESOLVE_MODE = 'exact'      # 'exact' | 'qutip_emulation'
MESOLVE_CALLS = 0


class Qobj:
    __array_priority__ = 100

    def __init__(self, data=None):
        if data is None:
            data = [[0.0]]
        if isinstance(data, Qobj):
            data = data._d
        self._d = np.array(data, dtype=complex)
        if self._d.ndim == 1:
            self._d = self._d.reshape(-1, 1)

    # --- the small algebra the reference uses
    def full(self):
        return self._d.copy()

    def __array__(self, dtype=None, copy=None):
        return self._d.astype(dtype) if dtype is not None else self._d.copy()

    @property
    def shape(self):
        return self._d.shape

    @property
    def isherm(self):
        return self._d.shape[0] == self._d.shape[1] and np.allclose(self._d, self._d.conj().T, atol=1e-12, rtol=0)

    def __mul__(self, other):
        if isinstance(other, Qobj):
            return Qobj(self._d @ other._d)
        return Qobj(self._d * other)

    def __rmul__(self, other):
        return Qobj(other * self._d)

    def __add__(self, other):
        if isinstance(other, Qobj):
            return Qobj(self._d + other._d)
        return Qobj(self._d + other)

    __radd__ = __add__

    def __sub__(self, other):
        return Qobj(self._d - (other._d if isinstance(other, Qobj) else other))

    def __truediv__(self, other):
        return Qobj(self._d / other)

    def __eq__(self, other):
        if not isinstance(other, Qobj):
            return False
        return self._d.shape == other._d.shape and bool(np.all(self._d == other._d))

    __hash__ = object.__hash__

    def dag(self):
        return Qobj(self._d.conj().T)

    def unit(self):
        if self._d.shape[1] == 1:
            return Qobj(self._d / np.linalg.norm(self._d))
        return Qobj(self._d / np.trace(self._d))

    def __repr__(self):
        return 'Qobj(shim)\n' + repr(self._d)


def tensor(*args):
    out = np.array([[1.0 + 0j]])
    for a in args:
        out = np.kron(out, a._d)
    return Qobj(out)


def sigmaz():
    return Qobj([[1, 0], [0, -1]])


def sigmax():
    return Qobj([[0, 1], [1, 0]])


def sigmay():
    return Qobj([[0, -1j], [1j, 0]])


def sigmap():
    return Qobj([[0, 1], [0, 0]])


def sigmam():
    return Qobj([[0, 0], [1, 0]])


def identity(n):
    return Qobj(np.eye(n))


qeye = identity


def basis(n, k=0):
    v = np.zeros((n, 1), dtype=complex)
    v[k, 0] = 1
    return Qobj(v)


def ket(seq, dim=2):
    out = np.array([[1.0 + 0j]])
    for s in seq:
        out = np.kron(out, basis(dim, int(s))._d)
    return Qobj(out)


def ket2dm(q):
    return Qobj(q._d @ q._d.conj().T)


def maximally_mixed_dm(n):
    return Qobj(np.eye(n) / n)


class Options:
    def __init__(self, **kw):
        self.atol = 1e-8
        self.rtol = 1e-6
        self.method = 'adams'
        self.order = 12
        self.nsteps = 1000
        for k, v in kw.items():
            setattr(self, k, v)


class Result:
    def __init__(self):
        self.expect = []
        self.times = None


def mesolve(H, rho0, tlist, c_ops=None, e_ops=None, args=None, options=None, **kw):
    """Same master equation and output convention as qutip 4.7.5 mesolve with e_ops given."""
    global MESOLVE_CALLS
    MESOLVE_CALLS += 1
    import scipy.linalg
    import scipy.sparse
    import scipy.integrate
    from oracle import lindblad_oracle as lo
    Hm = H.full()
    cs = [c.full() for c in (c_ops or [])]
    r0 = rho0.full()
    es = [e.full() for e in (e_ops or [])]
    tl = np.asarray(tlist, dtype=float)
    L = lo.liouvillian(Hm, cs)
    W = lo._trace_rows(es)
    out = np.zeros((len(es), len(tl)), dtype=complex)
    v = r0.ravel(order='F').astype(complex)
    if MESOLVE_MODE == 'exact':
        props = []
        for k in range(len(tl)):
            if k > 0:
                dt = tl[k] - tl[k - 1]
                P = next((P0 for (d0, P0) in props if abs(dt - d0) <= 1e-12 * max(abs(dt), abs(d0))), None)
                if P is None:
                    P = scipy.linalg.expm(L * dt)
                    props.append((dt, P))
                v = P @ v
            out[:, k] = W @ v
    else:
        opt = options or Options()
        Ls = scipy.sparse.csr_matrix(L)
        r = scipy.integrate.ode(lambda t, y: Ls @ y)
        r.set_integrator('zvode', method=opt.method, order=opt.order, atol=opt.atol, rtol=opt.rtol,
                         nsteps=int(opt.nsteps))
        r.set_initial_value(v, tl[0])
        for k in range(len(tl)):
            if k > 0:
                r.integrate(tl[k])
                if not r.successful():
                    raise Exception('ODE integration error: Try to increase the allowed number of substeps '
                                    'by increasing the nsteps parameter in the Options class.')
            out[:, k] = W @ r.y
    res = Result()
    res.times = tl
    res.expect = lo._finish(out, es, r0)
    return res
