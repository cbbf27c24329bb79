"""
Fixed-length parameter-vector layouts and the synthetic workload of SURVEY.md section 8(d).

`LibraryLayout` is the device-side view of `BasicModel.vectorise_under_library`
(reference basic_model.py:575-765): for a model family with Q qubits followed by D defects and
given process libraries, one column per possible process in the reference's order --
defect energies, qubit-defect couplings (pair-major, library order), defect-defect couplings,
qubit Ls, defect Ls -- followed by the (fixed) qubit energies.  A model is a row of values,
0.0 meaning "process absent".  The batched chain engine and the benchmark work on these rows.
"""
from __future__ import annotations

import numpy as np

from . import definitions, engine, learning_model

BOUNDS = {'couplings': (0.1, 4.0), 'energies': (0.01, 0.4), 'Ls': (0.01, 0.4)}   # configs.py:100-104


def _as_pairs(kind) -> tuple:
    if type(kind) == tuple and list(map(type, kind)) == [str, str]:
        return (kind,)
    return tuple(tuple(p) for p in kind)


class LibraryLayout:
    """Column layout of the parameter vectors of one model family (Q qubits then D defects)."""

    def __init__(self, n_qubits: int, n_defects: int, hyperparameters: dict):
        self.Q, self.D = n_qubits, n_defects
        self.n_tls = n_qubits + n_defects
        Q, D = n_qubits, n_defects
        keys, classes = [], []
        for v in range(D):
            keys.append(('E', Q + v)); classes.append('energies')
        for q in range(Q):
            for v in range(D):
                for kind in hyperparameters['qubit2defect_couplings_library']:
                    keys.append(('C', q, Q + v, _as_pairs(kind), 0)); classes.append('couplings')
        for v1 in range(D):
            for v2 in range(v1 + 1, D):
                for kind in hyperparameters['defect2defect_couplings_library']:
                    keys.append(('C', Q + v1, Q + v2, _as_pairs(kind), 0)); classes.append('couplings')
        for q in range(Q):
            for op in hyperparameters['qubit_Ls_library']:
                keys.append(('L', q, op)); classes.append('Ls')
        for v in range(D):
            for op in hyperparameters['defect_Ls_library']:
                keys.append(('L', Q + v, op)); classes.append('Ls')
        self.n_library = len(keys)                      # length of the reference's vector
        for q in range(Q):
            keys.append(('E', q)); classes.append('qubit_energy')
        self.keys = keys
        self.classes = classes
        self.n_params = len(keys)

    def context(self, times, obs_names, qubit_state=None, defect_state=None, device=None, kernel=None, flags=None) -> engine.EvalContext:
        """EvalContext whose parameter columns are exactly this layout."""
        qs = definitions.ops['plus'] if qubit_state is None else qubit_state
        ds = definitions.ops['mm'] if defect_state is None else defect_state
        ctx = engine.EvalContext(self.n_tls, [qs] * self.Q + [ds] * self.D, 0, list(obs_names), times,
                                 device=device, kernel=kernel, flags=flags)
        ctx.ensure_columns(self.keys)
        assert [ctx.columns[k] for k in self.keys] == list(range(self.n_params))
        return ctx

    def random_params(self, rng: np.random.Generator, p_present: float = 0.5, qubit_energy: float = 1.0) -> np.ndarray:
        """
        One random model inside the configs.py bounds: defect energies U(0.01,0.4); every library process
        present with probability p_present, couplings U(0.1,4), rates U(0.01,0.4); qubit energy 1
        (execute_learning_full.py:161).  Draw order = column order.
        """
        out = np.zeros(self.n_params)
        for c, cls in enumerate(self.classes):
            if cls == 'energies':
                out[c] = rng.uniform(*BOUNDS['energies'])
            elif cls == 'qubit_energy':
                out[c] = qubit_energy
            elif rng.uniform() < p_present:
                out[c] = rng.uniform(*BOUNDS[cls])
        return out

    def batch_params(self, seeds, p_present: float = 0.5) -> np.ndarray:
        return np.stack([self.random_params(np.random.default_rng(int(s)), p_present) for s in seeds])

    def model_from_params(self, row, qubit_state=None, defect_state=None) -> learning_model.LearningModel:
        """LearningModel (qubits first, couplings held by the first partner) for one parameter row."""
        qs = definitions.ops['plus'] if qubit_state is None else qubit_state
        ds = definitions.ops['mm'] if defect_state is None else defect_state
        m = learning_model.LearningModel()
        for k in range(self.n_tls):
            m.add_TLS(is_qubit=k < self.Q, energy=0.0, couplings={}, Ls={}, initial_state=qs if k < self.Q else ds)
        for key, v in zip(self.keys, row):
            v = float(v)
            if key[0] == 'E':
                m.TLSs[key[1]].energy = v
            elif v == 0.0:
                continue
            elif key[0] == 'C':
                m.TLSs[key[1]].couplings.setdefault(m.TLSs[key[2]], []).append((v, [tuple(p) for p in key[3]]))
            else:
                m.TLSs[key[1]].Ls[key[2]] = v
        m.build_operators()
        return m
