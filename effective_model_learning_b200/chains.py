"""
`BatchedChains`: thousands of independent reversible-jump MCMC chains advanced together on one GPU
(scope row N1).  Each chain follows the per-step semantics of `LearningChain.run`
(reference learning_chain.py:346-596) on the fixed-length parameter vector of
`vectorise_under_library`; proposal, evaluation and accept/reject all run on the device
(csrc/eml_chains.cu + the evaluator kernels), the host only launches and reads statistics.

`LearningChain` remains the single-chain drop-in whose np.random stream and trajectories match the
reference; this engine uses its own counter-based per-chain random streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native, definitions, dist, synthetic

MOVES = ('tweak all parameters', 'add qubit L', 'remove qubit L', 'add defect L', 'remove defect L',
         'add qubit-defect coupling', 'remove qubit-defect coupling',
         'add defect-defect coupling', 'remove defect-defect coupling')
_PCLASS = ('couplings', 'energies', 'Ls')


def column_classes(layout: synthetic.LibraryLayout) -> np.ndarray:
    out = np.empty(layout.n_params, dtype=np.int32)
    for c, (key, cls) in enumerate(zip(layout.keys, layout.classes)):
        if cls == 'energies':
            out[c] = _native.COL_DEFECT_ENERGY
        elif cls == 'qubit_energy':
            out[c] = _native.COL_FIXED
        elif cls == 'couplings':
            out[c] = _native.COL_Q2D_COUPLING if key[1] < layout.Q else _native.COL_D2D_COUPLING
        else:
            out[c] = _native.COL_QUBIT_L if key[1] < layout.Q else _native.COL_DEFECT_L
    return out


class BatchedChains:
    """
    n_chains chains for a model family of `n_qubits` qubits + `n_defects` defects.

    target_datasets: [K][T] (shared by all chains) or [S][K][T] with `target_set[n_chains]` choosing one per chain.
    Hyperparameters take the names, meaning AND DEFAULTS of LearningChain's (configs.py dictionaries can be passed as
    **config): qubits start in `plus`, defects in `gnd` (basic_model.py:259-281) -- the reference's production runs pass
    defect_initial_state=ops['mm'] explicitly (execute_learning_full.py:151-164), and so must callers of this class;
    jump lengths and step options fall back to LearningChain.Defaults.  The one exception is params_bounds, which is
    required: the device-side proposal kernel samples new processes uniformly inside the bounds
    (process_handling.py:98-103) and has no gamma-prior branch.
    Chain c uses the random stream of global chain number first_chain + c.
    """

    def __init__(self, target_times, target_datasets, target_observables, *, n_chains: int, n_defects: int,
                 n_qubits: int = 1, qubit_energy: float = 1.0, qubit_initial_state=None, defect_initial_state=None,
                 target_set=None, seed: int = 0, first_chain: int = 0, device: int = 0, initial_params=None,
                 chain_step_options=None, temperature_proposal=0.0005, shock_anneal_at=False, complexity_factor=1,
                 jump_length_rescaling_factor=1.0, acceptance_window=10, acceptance_target=0.4, acceptance_band=0.2,
                 params_handler_hyperparams=None, qubit_Ls_library=None, defect_Ls_library=None,
                 qubit2defect_couplings_library=None, defect2defect_couplings_library=None,
                 params_bounds=None, **unused):
        if params_bounds is None or not params_bounds:
            raise RuntimeError('BatchedChains needs params_bounds (rejection-sampling bounds for every parameter class)')
        hyper = {'qubit_Ls_library': qubit_Ls_library or {}, 'defect_Ls_library': defect_Ls_library or {},
                 'qubit2defect_couplings_library': qubit2defect_couplings_library or {},
                 'defect2defect_couplings_library': defect2defect_couplings_library or {}}
        self.layout = synthetic.LibraryLayout(n_qubits, n_defects, hyper)
        self.hyperparameters = hyper
        self.n_chains = int(n_chains)
        obs = [target_observables] if isinstance(target_observables, str) else list(target_observables)
        targets = np.asarray(target_datasets, dtype=np.complex128)
        if targets.ndim == 1:
            targets = targets[None]
        from .learning_chain import LearningChain
        defaults = LearningChain.Defaults
        if qubit_initial_state is None or qubit_initial_state is False:
            qubit_initial_state = definitions.ops['plus']
        if defect_initial_state is None or defect_initial_state is False:
            defect_initial_state = definitions.ops['gnd']
        self.ctx = self.layout.context(np.asarray(target_times, dtype=np.float64), obs,
                                       qubit_state=qubit_initial_state, defect_state=defect_initial_state, device=device)
        self.ctx.set_targets(targets)
        plan = self.ctx.pin_plan()      # the native engine keeps a pointer to the plan: it must outlive the engine
        self._pinned = True

        steps = dict(chain_step_options or defaults.chain_step_options)
        jumps = params_handler_hyperparams or defaults.params_handler_hyperparams
        ji = jumps.get('initial_jump_lengths', defaults.params_handler_hyperparams['initial_jump_lengths'])
        ja = jumps.get('annealed_jump_lengths', ji)
        cls = column_classes(self.layout)
        self._cls = cls
        h = _native.EmlChainHyper()
        h.n_params = self.layout.n_params
        h.acceptance_window = int(acceptance_window)
        h.col_class = cls.ctypes.data_as(C.POINTER(C.c_int32))
        for m, name in enumerate(MOVES):
            h.priority[m] = float(steps.get(name, 0.0))
        if isinstance(temperature_proposal, (tuple, list)):
            h.temperature, h.temperature_scale = float(temperature_proposal[0]), float(temperature_proposal[1])
        else:
            h.temperature, h.temperature_scale = float(temperature_proposal), 0.0
        h.complexity_factor = float(complexity_factor)
        for k, name in enumerate(_PCLASS):
            h.bounds_lo[k], h.bounds_hi[k] = float(params_bounds[name][0]), float(params_bounds[name][1])
            h.jump_initial[k], h.jump_annealed[k] = float(ji[name]), float(ja[name])
        h.jump_rescale = float(jump_length_rescaling_factor)
        h.acceptance_target, h.acceptance_band = float(acceptance_target), float(acceptance_band)
        h.shock_anneal_at = int(shock_anneal_at) if shock_anneal_at else 0
        self._hyper = h

        if initial_params is None:
            # LearningChain.make_initial_model: defects at energy 0.1, no processes (learning_chain.py:618-656)
            row = np.zeros(self.layout.n_params)
            for c, kind in enumerate(self.layout.classes):
                row[c] = 0.1 if kind == 'energies' else (qubit_energy if kind == 'qubit_energy' else 0.0)
            initial_params = np.tile(row, (self.n_chains, 1))
        initial_params = np.ascontiguousarray(initial_params, dtype=np.float64)
        if initial_params.shape != (self.n_chains, self.layout.n_params):
            raise RuntimeError('initial_params must be [n_chains][n_params]')
        ts = None
        if target_set is not None:
            ts = np.ascontiguousarray(target_set, dtype=np.int32)
        lib = _native.load()
        self._lib = lib
        handle = C.c_void_p()
        try:
            _native.check(lib.eml_chains_create(plan.handle, C.byref(h), self.n_chains,
                                                initial_params.ctypes.data_as(C.POINTER(C.c_double)),
                                                ts.ctypes.data_as(C.POINTER(C.c_int32)) if ts is not None else None,
                                                C.c_uint64(seed & (2 ** 64 - 1)), int(first_chain), C.byref(handle)), 'eml_chains_create')
        except Exception:
            self._unpin()
            raise
        self._h = handle
        self.steps_done = 0

    def run(self, steps: int, history=None, stream: int = 0):
        """
        `steps` proposals per chain (asynchronous).  history: dict of torch CUDA tensors, any of
        'loss', 'aprob' [steps][C] float64, 'code' [steps][C] int32 (move << 1 | accepted),
        'params' [steps][C][P] float64 (current model after each step; accepted models = rows with code & 1).
        """
        ptr = {k: None for k in ('loss', 'aprob', 'code', 'params')}
        if history is not None:
            for k in ptr:
                if k in history:
                    ptr[k] = history[k].data_ptr()
        _native.check(self._lib.eml_chains_run(self._h, int(steps), C.c_void_p(ptr['loss']), C.c_void_p(ptr['aprob']),
                                               C.c_void_p(ptr['code']), C.c_void_p(ptr['params']),
                                               C.c_void_p(stream or None)), 'eml_chains_run')
        self.steps_done += int(steps)

    def read(self):
        """Synchronise and return a dict of host arrays: current/best params and loss, stats table."""
        Cn, P = self.n_chains, self.layout.n_params
        out = {'current_params': np.empty((Cn, P)), 'current_loss': np.empty(Cn), 'best_params': np.empty((Cn, P)),
               'best_loss': np.empty(Cn), 'stats': np.empty((Cn, _native.CHAIN_STATS))}
        dp = C.POINTER(C.c_double)
        _native.check(self._lib.eml_chains_read(self._h, out['current_params'].ctypes.data_as(dp),
                                                out['current_loss'].ctypes.data_as(dp),
                                                out['best_params'].ctypes.data_as(dp), out['best_loss'].ctypes.data_as(dp),
                                                out['stats'].ctypes.data_as(dp)), 'eml_chains_read')
        return out

    def checkpoint(self, n_chains_global: int | None = None) -> np.ndarray:
        """Global [n_chains][8] statistics table on every rank: ONE all_gather (dist.gather_chain_stats)."""
        stats = self.read()['stats']
        return dist.gather_chain_stats(stats, n_chains_global if n_chains_global is not None else self.n_chains)

    def export_state(self) -> bytes:
        """Complete chain state as an opaque blob (checkpoint); see import_state."""
        n = self._lib.eml_chains_state_bytes(self._h)
        buf = C.create_string_buffer(n)
        _native.check(self._lib.eml_chains_export(self._h, buf), 'eml_chains_export')
        return buf.raw

    def import_state(self, blob: bytes):
        """Resume from a blob exported by an engine with the same layout, hyperparameters, n_chains and first_chain."""
        if len(blob) != self._lib.eml_chains_state_bytes(self._h):
            raise RuntimeError('checkpoint size does not match this engine')
        _native.check(self._lib.eml_chains_import(self._h, C.c_char_p(blob)), 'eml_chains_import')

    def best_models(self, which=None):
        """LearningModel objects of the best model of each chain (or of the chains listed in `which`)."""
        state = self.read()
        idx = range(self.n_chains) if which is None else which
        out = []
        for c in idx:
            m = self.layout.model_from_params(state['best_params'][c], self.ctx.site_states[0], self.ctx.site_states[-1])
            m.final_loss = float(state['best_loss'][c])
            out.append(m)
        return out

    def bad_evaluations(self):
        """
        (per-chain counts, total) of proposals that were rejected because their evaluation did not converge or gave a
        non-finite loss (the reference's mesolve would have evaluated them; DESIGN.md, deviations).
        """
        per = np.zeros(self.n_chains, dtype=np.int32)
        tot = C.c_int64(0)
        _native.check(self._lib.eml_chains_bad_evaluations(self._h, per.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(tot)),
                      'eml_chains_bad_evaluations')
        return per, int(tot.value)

    def _unpin(self):
        if getattr(self, '_pinned', False):
            self._pinned = False
            self.ctx.unpin_plan()

    def close(self):
        if getattr(self, '_h', None):
            self._lib.eml_chains_destroy(self._h)
            self._h = None
        self._unpin()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
