"""
ctypes binding of libeml_b200.so (C ABI declared in include/eml_b200.h).

There is deliberately no fallback: if the shared library or a CUDA device is missing,
every compute entry point raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('EML_B200_LIB', os.path.join(_HERE, 'libeml_b200.so'))   # override: kernel-variant experiments

EML_OK = 0
EML_ERR_NOT_CONVERGED = -4
TERM_ENERGY, TERM_COUPLING, TERM_LINDBLAD = 0, 1, 2
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_WARP_MMA, KERNEL_CTA_MMA, KERNEL_WARP_CHEB, KERNEL_CTA_CHEB = 0, 1, 2, 3, 4, 5
FLAG_NO_SECTOR_REDUCTION = 1     # EmlOptions.reserved bit 0: evolve the whole density matrix even when a parity sector is steady

# every symbol include/eml_b200.h declares (tests check the .so exports all of them)
SYMBOLS = ('eml_version', 'eml_last_error_string', 'eml_device_count', 'eml_plan_create', 'eml_plan_destroy',
           'eml_plan_set_targets', 'eml_eval_batch', 'eml_eval_batch_host', 'eml_launch_count', 'eml_debug_phase_clocks',
           'eml_plan_kernel_name', 'eml_measure_fp64_peak', 'eml_chains_create', 'eml_chains_destroy',
           'eml_chains_run', 'eml_chains_read', 'eml_chains_stats_device', 'eml_chains_bad_evaluations', 'eml_chains_state_bytes',
           'eml_chains_export', 'eml_chains_import')
CHAIN_STATS = 8
COL_DEFECT_ENERGY, COL_Q2D_COUPLING, COL_D2D_COUPLING, COL_QUBIT_L, COL_DEFECT_L, COL_FIXED = range(6)


class EmlTerm(C.Structure):
    _fields_ = [('param', C.c_int32), ('kind', C.c_int32), ('site_a', C.c_int32), ('site_b', C.c_int32),
                ('op_a', C.c_int32), ('op_b', C.c_int32)]


class EmlProblemDesc(C.Structure):
    _fields_ = [('n_tls', C.c_int32), ('n_params', C.c_int32), ('n_terms', C.c_int32),
                ('terms', C.POINTER(EmlTerm)),
                ('n_ops', C.c_int32), ('op_table', C.POINTER(C.c_double)),
                ('rho0', C.POINTER(C.c_double)),
                ('obs_site', C.c_int32), ('n_obs', C.c_int32), ('obs', C.POINTER(C.c_double)),
                ('n_times', C.c_int32), ('times', C.POINTER(C.c_double)),
                ('n_target_sets', C.c_int32), ('targets', C.POINTER(C.c_double))]


class EmlOptions(C.Structure):
    _fields_ = [('theta_target', C.c_double), ('tol', C.c_double), ('kernel', C.c_int32), ('reserved', C.c_int32)]


class EmlChainHyper(C.Structure):
    _fields_ = [('n_params', C.c_int32), ('acceptance_window', C.c_int32), ('col_class', C.POINTER(C.c_int32)),
                ('priority', C.c_double * 9), ('temperature', C.c_double), ('temperature_scale', C.c_double),
                ('complexity_factor', C.c_double), ('bounds_lo', C.c_double * 3), ('bounds_hi', C.c_double * 3),
                ('jump_initial', C.c_double * 3), ('jump_annealed', C.c_double * 3), ('jump_rescale', C.c_double),
                ('acceptance_target', C.c_double), ('acceptance_band', C.c_double), ('shock_anneal_at', C.c_int64)]


_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU fallback)')
        lib = C.CDLL(LIB_PATH)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
        lib.eml_version.restype = C.c_int
        lib.eml_last_error_string.restype = C.c_char_p
        lib.eml_device_count.argtypes = [C.POINTER(C.c_int)]
        lib.eml_plan_create.argtypes = [C.POINTER(EmlProblemDesc), C.c_int, C.POINTER(EmlOptions), C.POINTER(vp)]
        lib.eml_plan_destroy.argtypes = [vp]
        lib.eml_plan_set_targets.argtypes = [vp, C.c_int32, dp]
        lib.eml_eval_batch.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp]
        lib.eml_eval_batch_host.argtypes = [vp, C.c_int64, dp, ip, dp, dp]
        lib.eml_launch_count.restype = C.c_int64
        lib.eml_plan_kernel_name.argtypes = [vp]
        lib.eml_plan_kernel_name.restype = C.c_char_p
        lib.eml_measure_fp64_peak.argtypes = [C.c_int, C.c_int, C.c_double, dp]
        lib.eml_chains_create.argtypes = [vp, C.POINTER(EmlChainHyper), C.c_int64, dp, ip, C.c_uint64, C.c_int64, C.POINTER(vp)]
        lib.eml_chains_destroy.argtypes = [vp]
        lib.eml_chains_run.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp]
        lib.eml_chains_read.argtypes = [vp, dp, dp, dp, dp, dp]
        lib.eml_chains_stats_device.argtypes = [vp]
        lib.eml_chains_stats_device.restype = vp
        lib.eml_chains_state_bytes.argtypes = [vp]
        lib.eml_chains_bad_evaluations.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        lib.eml_chains_state_bytes.restype = C.c_size_t
        lib.eml_chains_export.argtypes = [vp, vp]
        lib.eml_chains_import.argtypes = [vp, vp]
        _lib = lib
    return _lib


def last_error() -> str:
    return load().eml_last_error_string().decode()


class NotConvergedError(RuntimeError):
    """EML_ERR_NOT_CONVERGED: at least one model of the batch hit the series cap or the step-count / overflow guard."""


def check(rc: int, what: str):
    if rc == EML_ERR_NOT_CONVERGED:
        raise NotConvergedError(f'{what} failed ({rc}): {last_error()}')
    if rc != EML_OK:
        raise RuntimeError(f'{what} failed ({rc}): {last_error()}')


def device_count() -> int:
    n = C.c_int(0)
    load().eml_device_count(C.byref(n))
    return n.value


def launch_count() -> int:
    return int(load().eml_launch_count())


def measure_fp64_peak(device: int = 0, use_mma=False, seconds: float = 0.5) -> float:
    """use_mma: False = DFMA chains, True = DMMA m8n8k4 chains, 2 = both interleaved in every warp."""
    out = C.c_double(0.0)
    check(load().eml_measure_fp64_peak(device, int(use_mma), seconds, C.byref(out)), 'eml_measure_fp64_peak')
    return out.value


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Plan:
    """Owns one EmlPlan (device copies of the data shared by a batch of models)."""

    def __init__(self, n_tls, n_params, terms, op_table, rho0, obs_site, obs, times, targets=None,
                 device=0, kernel=KERNEL_AUTO, theta_target=0.0, tol=0.0, flags=0):
        lib = load()
        self._lib = lib
        self.n_tls, self.n_params = int(n_tls), int(n_params)
        self.op_table = np.ascontiguousarray(op_table, dtype=np.complex128).reshape(-1, 2, 2)
        self.rho0 = np.ascontiguousarray(rho0, dtype=np.complex128)
        self.obs = np.ascontiguousarray(obs, dtype=np.complex128).reshape(-1, 2, 2)
        self.times = np.ascontiguousarray(times, dtype=np.float64)
        self.n_obs, self.n_times = self.obs.shape[0], self.times.shape[0]
        d = 1 << self.n_tls
        if self.rho0.shape != (d, d):
            raise RuntimeError('rho0 has the wrong shape for n_tls')
        tarr = (EmlTerm * max(1, len(terms)))()
        for i, t in enumerate(terms):
            tarr[i] = EmlTerm(*[int(x) for x in t])
        tg = None
        nsets = 0
        if targets is not None:
            tg = np.ascontiguousarray(targets, dtype=np.complex128)
            if tg.ndim == 2:
                tg = tg[None]
            if tg.shape[1:] != (self.n_obs, self.n_times):
                raise RuntimeError('targets must have shape [S][K][T]')
            nsets = tg.shape[0]
        desc = EmlProblemDesc(self.n_tls, self.n_params, len(terms), tarr,
                              self.op_table.shape[0], _dptr(self.op_table.view(np.float64)),
                              _dptr(self.rho0.view(np.float64)),
                              int(obs_site), self.n_obs, _dptr(self.obs.view(np.float64)),
                              self.n_times, _dptr(self.times),
                              nsets, _dptr(tg.view(np.float64)) if tg is not None else None)
        opts = EmlOptions(float(theta_target), float(tol), int(kernel), int(flags))
        handle = C.c_void_p()
        check(lib.eml_plan_create(C.byref(desc), int(device), C.byref(opts), C.byref(handle)), 'eml_plan_create')
        self._h = handle
        self.device = int(device)
        self.n_target_sets = nsets

    @property
    def handle(self):
        return self._h

    @property
    def kernel_name(self) -> str:
        return self._lib.eml_plan_kernel_name(self._h).decode()

    def set_targets(self, targets):
        tg = np.ascontiguousarray(targets, dtype=np.complex128)
        if tg.ndim == 2:
            tg = tg[None]
        if tg.shape[1:] != (self.n_obs, self.n_times):
            raise RuntimeError('targets must have shape [S][K][T]')
        check(self._lib.eml_plan_set_targets(self._h, tg.shape[0], _dptr(tg.view(np.float64))), 'eml_plan_set_targets')
        self.n_target_sets = tg.shape[0]

    def eval_host(self, params, target_set=None, want_expect=True, want_loss=True, loss_out=None, expect_out=None):
        """params: [B][P] float64 host array.  Returns (expect [B][K][T] complex | None, loss [B] | None).
        loss_out / expect_out: optional preallocated result arrays; page-locked arrays (and page-locked params, e.g.
        views of torch pin_memory() tensors) are used for the DMA directly instead of the plan's staging buffers."""
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, max(self.n_params, 0)) \
            if self.n_params > 0 else np.zeros((int(np.shape(params)[0]), 0))
        B = params.shape[0]
        expect = loss = None
        if want_expect:
            expect = expect_out if expect_out is not None else np.empty((B, self.n_obs, self.n_times), dtype=np.complex128)
            if expect.shape != (B, self.n_obs, self.n_times) or expect.dtype != np.complex128 or not expect.flags.c_contiguous:
                raise RuntimeError('expect_out must be a C-contiguous complex128 array of shape [B][K][T]')
        if want_loss:
            loss = loss_out if loss_out is not None else np.empty(B, dtype=np.float64)
            if loss.shape != (B,) or loss.dtype != np.float64 or not loss.flags.c_contiguous:
                raise RuntimeError('loss_out must be a C-contiguous float64 array of shape [B]')
        ts = None
        if target_set is not None:
            ts = np.ascontiguousarray(target_set, dtype=np.int32)
            if ts.shape != (B,):
                raise RuntimeError('target_set must have one entry per model')
        rc = self._lib.eml_eval_batch_host(
            self._h, B, _dptr(params) if params.size else None,
            ts.ctypes.data_as(C.POINTER(C.c_int32)) if ts is not None else None,
            _dptr(expect.view(np.float64)) if expect is not None else None,
            _dptr(loss) if loss is not None else None)
        check(rc, 'eml_eval_batch_host')
        return expect, loss

    def eval_device(self, n_models, params_ptr, target_set_ptr=0, expect_ptr=0, loss_ptr=0, status_ptr=0, stream=0):
        """Raw device-pointer entry point (integers from torch's .data_ptr()); no sync, no allocation."""
        rc = self._lib.eml_eval_batch(self._h, int(n_models), C.c_void_p(params_ptr), C.c_void_p(target_set_ptr or None),
                                      C.c_void_p(expect_ptr or None), C.c_void_p(loss_ptr or None),
                                      C.c_void_p(status_ptr or None), C.c_void_p(stream or None))
        check(rc, 'eml_eval_batch')

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self._lib.eml_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
