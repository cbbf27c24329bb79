"""
Operator table and small helpers shared by the host-side model classes.

Mirrors the public names of the reference's `definitions.py` (`ops`, `T`, `Constants`,
`ops_labels`, `observable_shorthand2pretty`, `MSE`; reference definitions.py:21-118) but
holds plain complex128 numpy 2x2 arrays instead of `qutip.Qobj`, in qutip's conventions
(basis(2,0) first; sigmap = |0><1|).  The table is what the CUDA kernels receive as their
operator library (`EmlProblemDesc.op_table`).
"""
from __future__ import annotations

import numpy as np


class Constants:
    """Unit conversions (h_bar = e = 1); reference definitions.py:21-25."""
    K_to_eV = 1.381 / 1.602 * 1e-4
    t_to_sec = 4.136e-15


def _ket(*amplitudes) -> np.ndarray:
    return np.array(amplitudes, dtype=np.complex128).reshape(-1, 1)


def _dm_of_unit_ket(k: np.ndarray) -> np.ndarray:
    k = k / np.linalg.norm(k)          # qutip: ket.unit() first, then ket2dm
    return k @ k.conj().T


_sm = np.array([[0, 0], [1, 0]], dtype=np.complex128)   # |1><0|
_sp = _sm.conj().T.copy()                               # |0><1|

# insertion order is part of the contract: op codes sent to the device are indices into it
ops: dict[str, np.ndarray] = {
    'sigmaz': np.diag([1.0 + 0j, -1.0]),
    'sigmax': np.array([[0, 1], [1, 0]], dtype=np.complex128),
    'sigmay': np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    'sigmap': _sp,
    'sigmam': _sm,
    'id2': np.eye(2, dtype=np.complex128),
    'gnd': _sm @ _sp,
    'exc': _sp @ _sm,
    'plus': _dm_of_unit_ket(_ket(0, 1) + _ket(1, 0)),
    '2e+1g': _dm_of_unit_ket(2 * _ket(0, 1) + 1 * _ket(1, 0)),
    'custom': np.full((2, 2), 0.5, dtype=np.complex128),
    'mm': np.eye(2, dtype=np.complex128) / 2,
}


def T(arg1, arg2):
    """
    Kronecker product in qutip.tensor order; a plain number as first factor returns the
    second (loop seed), as reference definitions.py:28-38.
    """
    if isinstance(arg1, np.ndarray):
        return np.kron(arg1, arg2)
    if type(arg1) in (int, float):
        return arg2
    return None


ops_labels = {
    'defects energies': 'energy (eV)',
    'defects couplings': 'coupling (eV)',
    **{f'sigma {k}': rf'$\sigma_{s}$' + ' rate (eV)'
       for k, s in (('z', 'z'), ('x', 'x'), ('y', 'y'), ('plus', '+'), ('minus', '-'))},
}

observable_shorthand2pretty = {
    **{short: rf'$\langle\sigma_{s}\rangle$' for short, s in (('sx', 'x'), ('sy', 'y'), ('sz', 'z'), ('sp', '+'), ('sm', '-'))},
    **{long: rf'$\langle\sigma_{s}\rangle$' for long, s in (('sigmax', 'x'), ('sigmay', 'y'), ('sigmaz', 'z'),
                                                           ('sigmap', '+'), ('sigmam', '-'))},
}


def MSE(A, B):
    """Mean squared error of two equal-length numpy arrays (reference definitions.py:99-113)."""
    if not isinstance(A, np.ndarray) or not isinstance(B, np.ndarray):
        raise RuntimeError('error calculating MSE: arguments must be numpy arrays!\n')
    if len(A) != len(B):
        raise RuntimeError('error calculating MSE: arguments must be same length!\n')
    return np.sum(np.square(np.abs(A - B))) / len(A)
