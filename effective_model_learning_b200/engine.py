"""
Marshalling between the host-side model objects and the C ABI (include/eml_b200.h).

A *context* fixes everything a batch of models shares -- number of sites, per-site initial
states, observable site/operators, evaluation times, targets -- plus a growing *process
library*: the union of every process type (energy / coupling / Lindblad on given sites with
given operators) seen so far.  A model is sent to the device as one row of parameter values
under that library (0.0 = process absent), exactly the fixed-length representation the
reference defines in `BasicModel.vectorise_under_library` (basic_model.py:575-765).  The
context owns the `_native.Plan`; it is rebuilt only when the library grows.

No CPU fallback: without the CUDA library/device the evaluation raises RuntimeError.
"""
from __future__ import annotations

import collections

import numpy as np

from . import _native
from . import definitions

_MAX_CONTEXTS = 32
_contexts: "collections.OrderedDict[tuple, EvalContext]" = collections.OrderedDict()
default_device = 0
default_kernel = _native.KERNEL_AUTO
default_flags = 0                       # EmlOptions.reserved bits (e.g. _native.FLAG_NO_SECTOR_REDUCTION), for tests / A-B runs


def _initial_state_array(tls) -> np.ndarray:
    """Per-site initial density matrix with the reference's defaults (basic_model.py:267-279)."""
    if tls.has_initial_state():
        return np.asarray(tls.initial_state, dtype=np.complex128)
    return definitions.ops['plus'] if tls.is_qubit else definitions.ops['gnd']


def product_state(site_states) -> np.ndarray:
    rho = np.ones((1, 1), dtype=np.complex128)
    for s in site_states:
        rho = np.kron(rho, s)
    return rho


def first_qubit_site(model) -> int | None:
    for k, tls in enumerate(model.TLSs):
        if tls.is_qubit:
            return k
    return None


def model_processes(model) -> list[tuple[tuple, float]]:
    """
    [(process key, value)] for one model.  Keys:
      ('E', site)                                   energy * sigmaz          (basic_model.py:206-221)
      ('C', site_a, site_b, ((opA, opB), ...), n)   strength (enters squared; basic_model.py:226-251)
      ('L', site, op)                               Lindblad rate            (basic_model.py:152-189)
    Couplings are canonicalised to site_a < site_b (operator pairs swapped accordingly);
    n counts repeats of the same coupling type on a pair (each keeps its own column because
    strengths enter squared and therefore do not add).
    """
    index = {id(t): k for k, t in enumerate(model.TLSs)}
    out = []
    seen: dict[tuple, int] = {}
    for k, tls in enumerate(model.TLSs):
        out.append((('E', k), float(tls.energy)))
    for k, tls in enumerate(model.TLSs):
        for partner, lst in tls.couplings.items():
            j = index[id(partner)]
            for strength, op_pairs in lst:
                pairs = tuple((str(a), str(b)) for a, b in op_pairs)
                if k < j:
                    base = ('C', k, j, pairs)
                else:
                    base = ('C', j, k, tuple((b, a) for a, b in pairs))
                n = seen.get(base, 0)
                seen[base] = n + 1
                out.append((base + (n,), float(strength)))
    for k, tls in enumerate(model.TLSs):
        for op, rate in tls.Ls.items():
            out.append((('L', k, str(op)), float(rate)))
    return out


class EvalContext:
    """Everything shared by a batch + the process library + the native plan."""

    def __init__(self, n_tls, site_states, obs_site, obs_names, times, device=None, kernel=None, flags=None):
        self.n_tls = n_tls
        self.site_states = [np.array(s, dtype=np.complex128) for s in site_states]
        self.rho0 = product_state(self.site_states)
        self.obs_site = obs_site
        self.obs_names = list(obs_names)
        self.times = np.ascontiguousarray(times, dtype=np.float64)
        self.device = default_device if device is None else device
        self.kernel = default_kernel if kernel is None else kernel
        self.flags = default_flags if flags is None else flags
        self.columns: dict[tuple, int] = {}     # process key -> parameter column
        self.op_index: dict[str, int] = {}      # operator name -> row of op table
        self.op_rows: list[np.ndarray] = []
        self.targets = None
        self._plan = None
        self._pins = 0      # live native objects (BatchedChains engines) that hold this context's plan

    # -- library management
    def _op(self, name: str) -> int:
        i = self.op_index.get(name)
        if i is None:
            i = len(self.op_rows)
            self.op_index[name] = i
            self.op_rows.append(np.array(definitions.ops[name], dtype=np.complex128))
        return i

    def ensure_columns(self, keys) -> bool:
        grew = False
        for key in keys:
            if key not in self.columns:
                self.columns[key] = len(self.columns)
                grew = True
        if grew:
            self._drop_plan()
        return grew

    def _drop_plan(self):
        if self._plan is not None:
            if self._pins:
                raise RuntimeError('the native plan of this context is in use by a live BatchedChains engine: close the '
                                   'engine before growing the process library or clearing the targets')
            self._plan.close()
            self._plan = None

    def pin_plan(self) -> '_native.Plan':
        """The plan, kept alive until unpin_plan(): anything that would destroy it raises instead."""
        plan = self.plan
        self._pins += 1
        return plan

    def unpin_plan(self):
        self._pins = max(0, self._pins - 1)

    def terms(self) -> list[tuple]:
        """Flat EmlTerm list (param, kind, site_a, site_b, op_a, op_b) of the current library."""
        out = []
        for key, col in self.columns.items():
            if key[0] == 'E':
                out.append((col, _native.TERM_ENERGY, key[1], -1, self._op('sigmaz'), -1))
            elif key[0] == 'C':
                for (a, b) in key[3]:
                    out.append((col, _native.TERM_COUPLING, key[1], key[2], self._op(a), self._op(b)))
            else:
                out.append((col, _native.TERM_LINDBLAD, key[1], -1, self._op(key[2]), -1))
        return out

    def set_targets(self, targets):
        """targets: [K][T] or [S][K][T]; None clears."""
        self.targets = None if targets is None else np.ascontiguousarray(targets, dtype=np.complex128)
        if self._plan is not None and self.targets is not None:
            self._plan.set_targets(self.targets)
        elif self._plan is not None:
            self._drop_plan()

    @property
    def plan(self) -> _native.Plan:
        if self._plan is None:
            terms = self.terms()
            obs = np.stack([definitions.ops[n] for n in self.obs_names]).astype(np.complex128)
            self._plan = _native.Plan(self.n_tls, len(self.columns), terms, np.stack(self.op_rows) if self.op_rows
                                      else np.zeros((1, 2, 2), dtype=np.complex128),
                                      self.rho0, self.obs_site, obs, self.times, self.targets,
                                      device=self.device, kernel=self.kernel, flags=self.flags)
        return self._plan

    def params_of(self, processes_per_model) -> np.ndarray:
        self.ensure_columns(key for procs in processes_per_model for key, _ in procs)
        out = np.zeros((len(processes_per_model), len(self.columns)), dtype=np.float64)
        for b, procs in enumerate(processes_per_model):
            for key, val in procs:
                out[b, self.columns[key]] = val
        return out

    def evaluate(self, params, target_set=None, want_expect=True, want_loss=False, loss_out=None, expect_out=None):
        return self.plan.eval_host(params, target_set, want_expect=want_expect, want_loss=want_loss,
                                   loss_out=loss_out, expect_out=expect_out)


def context_for(model, obs_names, times, device=None) -> EvalContext:
    """Find or create the context matching this model's shape, initial states, observables and times."""
    states = [_initial_state_array(t) for t in model.TLSs]
    times = np.ascontiguousarray(times, dtype=np.float64)
    q = first_qubit_site(model)
    key = (len(model.TLSs), q, tuple(obs_names), times.tobytes(), b''.join(s.tobytes() for s in states),
           default_device if device is None else device, default_kernel, default_flags)
    ctx = _contexts.get(key)
    if ctx is None:
        if q is None:
            # no qubit: the reference's observable degenerates to the identity (basic_model.py:365-375)
            ctx = EvalContext(len(model.TLSs), states, 0, ['id2'] * len(obs_names), times, device)
        else:
            ctx = EvalContext(len(model.TLSs), states, q, obs_names, times, device)
        _contexts[key] = ctx
        for old_key in [k for k, c in _contexts.items() if not c._pins][:max(0, len(_contexts) - _MAX_CONTEXTS)]:
            _contexts.pop(old_key)._drop_plan()
    else:
        _contexts.move_to_end(key)
    return ctx


def clear_contexts():
    for key in [k for k, c in _contexts.items() if not c._pins]:
        _contexts.pop(key)._drop_plan()


def is_hermitian(a: np.ndarray) -> bool:
    return bool(np.allclose(a, a.conj().T, rtol=0.0, atol=1e-12))


def evaluate_models(models, evaluation_times, obs_names, targets=None, want_expect=True, device=None):
    """
    Batched evaluation of models that share shape / initial states.  Returns
    (expect [B][K][T] complex128 or None, loss [B] or None).  `targets` [K][T] enables the loss.
    """
    if len(models) == 0:
        raise RuntimeError('no models to evaluate')
    ctx = context_for(models[0], obs_names, evaluation_times, device)
    for m in models[1:]:
        if (len(m.TLSs) != ctx.n_tls or first_qubit_site(m) != first_qubit_site(models[0])
                or any(not np.array_equal(_initial_state_array(t), s) for t, s in zip(m.TLSs, ctx.site_states))):
            raise RuntimeError('models of one batch must share the number of TLSs, the first qubit and the initial states')
    params = ctx.params_of([model_processes(m) for m in models])
    if targets is not None:
        ctx.set_targets(np.asarray(targets))
    elif ctx.targets is not None:
        ctx.set_targets(None)
    return ctx.evaluate(params, want_expect=want_expect, want_loss=targets is not None)


def flops_per_application(kernel_name: str, n_tls: int) -> float:
    """
    FP64 floating-point operations one application of the Lindblad generator to the d x d density
    matrix costs in each kernel (used for the roofline in bench.py).
      generic : per element 2 complex dot products of length d (G rho and rho Gt, 8 flops per complex
                multiply-add), n 4-term site stencils, scale and accumulate.
      warp_mma: one complex d^3 product on the DMMA pipe (rho Hermitian => rho G^dag = (G rho)^dag),
                the X + X^dag combine, n REAL-weighted site stencils, scale and accumulate;
      warp_mma_realH: the same with a real Hamiltonian, where only G.im = -H needs the DMMA pipe.
      warp_cheb*: the same application inside the Chebyshev recurrence (the scale and accumulate of the Taylor
                series become the fused + phi_{k-1}/2 and the halving of the stored term).
    """
    d = 2 ** n_tls
    if kernel_name == 'generic':
        return d * d * (16.0 * d + 32.0 * n_tls + 4.0)
    if kernel_name in ('warp_mma', 'warp_mma_realH', 'warp_cheb', 'warp_cheb_realH', 'cta_mma', 'cta_mma_realH',
                       'cta_cheb', 'cta_cheb_realH', 'warp_cheb_parity', 'warp_cheb_parity_realH',
                       'cta_cheb_parity', 'cta_cheb_parity_realH', 'warp_cheb_parity_sector', 'warp_cheb_parity_sector_realH'):
        # DMMA: 4 (complex H) or 2 (real H: G.re diagonal, folded into the stencil) real d^3 products;
        # per complex element: diagonal weight 2, n site stencils 4n, X + X^dag 2, scale 2, accumulate 2
        mma = (2.0 if kernel_name.endswith('realH') else 4.0) * 2.0 * d ** 3
        if 'parity' in kernel_name:      # H block diagonal by parity: only the diagonal tiles of G are multiplied
            mma *= 0.5
        if kernel_name.startswith('warp_cheb_parity'):
            # d = 16 parity variants (round 2): the state is tile (0,1) (its conjugate transpose is only an operand) plus,
            # without the sector reduction, the two Hermitian even tiles.  Odd tile, per complex element (64 of them):
            # diagonal weight and 4 site flips; G00 m and m G11^dag together are the same DMMA count as before.
            quarter = d * d / 4.0
            if 'sector' in kernel_name:      # weights unhalved, everything fused: 2 FMA + 4 x 2 FMA = 20 flop per element
                return 0.5 * mma + quarter * (4.0 + 4.0 * n_tls)
            # full state: odd tile 2 MUL + 8 FMA + the doubling FMA pair = 22 flop; even tiles (2 x 64 elements) 10 FMA,
            # X' + X'^dag 2 ADD, halving of the stored term 2 MUL = 24 flop
            return mma + quarter * (2.0 + 4.0 * n_tls + 4.0) + 2.0 * quarter * (4.0 + 4.0 * n_tls + 4.0)
        if kernel_name.startswith('cta_cheb_parity'):
            # d = 32 orbit kernel (round 2, eml_cta_orbit.cu): only the 10 upper-triangle tiles (640 complex elements) are
            # state: diagonal weight + 5 site flips = 2 + 10 FMA per element; the 4 Hermitian tiles (256 elements) also pay
            # X' + X'^dag (2 ADD) and the halving of the stored term (2 MUL).  Same DMMA count as the tile-row kernel.
            return mma + 640.0 * (4.0 + 4.0 * n_tls) + 256.0 * 4.0
        return mma + d * d * (2.0 + 4.0 * n_tls + 6.0)
    raise RuntimeError('unknown kernel ' + kernel_name)
