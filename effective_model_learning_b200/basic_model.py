"""
`BasicModel`: a model made of two-level systems with pairwise coherent couplings and
single-site Lindblad processes -- same public surface as the reference's BasicModel
(reference basic_model.py:23-765) with the dynamics evaluated by the CUDA library.

Differences that are deliberate:
  * operators are numpy arrays, not qutip.Qobj; `H`, `Ls`, `initial_DM` are built lazily
    from the TLS records (the device assembles its own copies from the parameter vector);
  * `calculate_dynamics` has ONE backend (libeml_b200.so); the legacy `dynamics_method`
    values are accepted and ignored (basic_model.py:335-338 did the same for unknown values);
  * `build_Liouvillian` returns the CORRECT column-stacked superoperator; the reference's
    (basic_model.py:285-310) mishandles complex jump operators;
  * no class-level `existing_models` registry (nothing reads it in the reference).
"""
from __future__ import annotations

import numpy as np

from . import definitions
from . import engine
from . import two_level_system

T = definitions.T
ops = definitions.ops

_SHORT = {'sigmax': ('sx', r'$\sigma_x$'), 'sigmay': ('sy', r'$\sigma_y$'), 'sigmaz': ('sz', r'$\sigma_z$'),
          'sigmap': ('sp', r'$\sigma_+$'), 'sigmam': ('sm', r'$\sigma_-$')}


class BasicModel:

    def __init__(self):
        self.TLSs: list[two_level_system.TwoLevelSystem] = []
        self.TLS_labels: dict[str, two_level_system.TwoLevelSystem] = {}
        self.Liouvillian = None
        self._built = False
        self._cache = {}

    # ------------------------------------------------------------------ construction
    def add_TLS(self, TLS_label: str = '', is_qubit: bool = False, energy=None, couplings: dict = None,
                Ls: dict = None, initial_state=False) -> two_level_system.TwoLevelSystem:
        """
        Appends a TwoLevelSystem.  Coupling partners may be given by label; a bare
        (strength, [...]) tuple or a bare (op, op) pair is wrapped into the list form
        {partner: [(strength, [(op_self, op_partner), ...])]} (reference basic_model.py:68-136).
        """
        couplings = {} if couplings is None else couplings
        Ls = {} if Ls is None else Ls
        for key in [k for k in couplings if isinstance(k, str)]:
            couplings[self.TLS_labels[key]] = couplings.pop(key)
        for partner, spec in couplings.items():
            if isinstance(spec, tuple):
                print('correction...')
                couplings[partner] = spec = [spec]
            for i, coupling in enumerate(spec):
                if len(coupling) != 2:
                    raise RuntimeError('Incorrect coupling specification:\n'
                                       + 'Use {partner: [(rate, [(op_self, op_partner), ...]]}')
                if isinstance(coupling[1], tuple):
                    spec[i] = (coupling[0], [coupling[1]])
        new_TLS = two_level_system.TwoLevelSystem(self, TLS_label, is_qubit, energy, couplings, Ls, initial_state)
        if TLS_label != '':
            self.TLS_labels[TLS_label] = new_TLS
        self.TLSs.append(new_TLS)
        return new_TLS

    def build_operators(self):
        """To be called after any change to the model; invalidates the lazily built operators."""
        self._built = True
        self._cache = {}

    # ------------------------------------------------------------------ full-space operators (lazy)
    def _site_embed(self, k: int, op: np.ndarray) -> np.ndarray:
        out = np.ones((1, 1), dtype=np.complex128)
        for s in range(len(self.TLSs)):
            out = np.kron(out, op if s == k else ops['id2'])
        return out

    def build_Ls(self) -> list[np.ndarray]:
        """sqrt(rate) * op on its site, TLS order then dict order (reference basic_model.py:152-189)."""
        out = [self._site_embed(k, ops[name] * np.sqrt(rate))
               for k, tls in enumerate(self.TLSs) for name, rate in tls.Ls.items()]
        self._cache['Ls'] = out
        return out

    def build_H(self):
        """sum_k E_k sigmaz_k + sum over couplings of (s*A)_i (s*B)_j (reference basic_model.py:193-255)."""
        if not self.TLSs:
            self._cache['H'] = None
            return None
        n = len(self.TLSs)
        H = np.zeros((2 ** n, 2 ** n), dtype=np.complex128)
        for k, tls in enumerate(self.TLSs):
            H += self._site_embed(k, ops['sigmaz'] * tls.energy)
        for i, tls in enumerate(self.TLSs):
            for partner, lst in tls.couplings.items():
                j = self.TLSs.index(partner)
                for strength, op_pairs in lst:
                    for op_i, op_j in op_pairs:
                        term = np.ones((1, 1), dtype=np.complex128)
                        for s in range(n):
                            term = np.kron(term, strength * ops[op_i] if s == i
                                           else strength * ops[op_j] if s == j else ops['id2'])
                        H += term
        self._cache['H'] = H
        return H

    def build_initial_DM(self) -> np.ndarray:
        """Product state; qubits default to 'plus', defects to 'gnd' (reference basic_model.py:259-281)."""
        rho = engine.product_state([engine._initial_state_array(t) for t in self.TLSs])
        self._cache['rho0'] = rho
        return rho

    @property
    def H(self):
        if not self._built:
            return None
        return self._cache['H'] if 'H' in self._cache else self.build_H()

    @property
    def Ls(self):
        if not self._built:
            return []
        return self._cache['Ls'] if 'Ls' in self._cache else self.build_Ls()

    @property
    def initial_DM(self):
        if not self._built:
            return None
        return self._cache['rho0'] if 'rho0' in self._cache else self.build_initial_DM()

    def build_Liouvillian(self) -> np.ndarray:
        """Column-stacked Lindblad superoperator (vec with order='F')."""
        H = self.H
        if H is None:
            raise RuntimeError('Liouvillian cannot be calculated due to missing TLSs or H')
        d = len(H)
        eye = np.eye(d)
        out = -1j * (np.kron(eye, H) - np.kron(H.T, eye))
        for c in self.Ls:
            cdc = c.conj().T @ c
            out += np.kron(c.conj(), c) - 0.5 * (np.kron(eye, cdc) + np.kron(cdc.T, eye))
        self.Liouvillian = out
        return out

    # ------------------------------------------------------------------ the hot path
    def calculate_dynamics(self, evaluation_times: np.ndarray, observable_ops=False, dynamics_method=False,
                           custom_function_on_return=False) -> list[np.ndarray]:
        """
        <O_m>(t_k) for each observable (applied to the first qubit, identities elsewhere) at
        each evaluation time; the initial state sits at evaluation_times[0].  Returns a list of
        K arrays of shape (T,): float64 for Hermitian observables, complex128 otherwise --
        the contract of reference basic_model.py:314-393 (qutip.mesolve(...).expect).
        """
        if dynamics_method is not False and dynamics_method not in ('qutip', 'liouvillian', 'cuda'):
            print('Model dynamics calculation method not recognised, hence using qutip')
        if not self._built or not self.TLSs:
            raise RuntimeError('Hamiltonian not defined -- cannot calculate dynamics')
        if isinstance(observable_ops, bool) and not observable_ops:
            observable_ops = 'exc'
        if isinstance(observable_ops, str):
            observable_ops = [observable_ops]
        if not (isinstance(observable_ops, list) and all(type(x) == str for x in observable_ops)):
            raise RuntimeError('do not understand oservable operators input for dynamics calculation')

        expect, _ = engine.evaluate_models([self], np.asarray(evaluation_times, dtype=np.float64), observable_ops)
        herm0 = all(engine.is_hermitian(engine._initial_state_array(t)) for t in self.TLSs)
        has_qubit = engine.first_qubit_site(self) is not None
        returnable = []
        for m, name in enumerate(observable_ops):
            if herm0 and (not has_qubit or engine.is_hermitian(ops[name])):
                returnable.append(np.ascontiguousarray(expect[0, m].real))
            else:
                returnable.append(np.ascontiguousarray(expect[0, m]))
        if not custom_function_on_return:
            return returnable
        return custom_function_on_return(returnable)

    # ------------------------------------------------------------------ descriptions
    def disp(self) -> None:
        print(self.model_description_str())

    def model_description_str(self) -> str:
        lines = ['______Model:______']
        for k, tls in enumerate(self.TLSs):
            lines.append('\n(' + str(k) + ') - ' + ('Qubit' if tls.is_qubit else 'Defect'))
            lines.append('Energy: ' + str(tls.energy))
            lines.append('Couplings:')
            for partner, lst in tls.couplings.items():
                lines.append('   Partner: TLS (' + str(self.TLSs.index(partner)) + ')')
                for strength, op_pairs in lst:
                    lines.append('      ' + str(strength) + ':')
                    lines.extend('         (' + str(a) + ', ' + str(b) + ')' for a, b in op_pairs)
            lines.append('Lindblad processes:')
            lines.extend('   ' + name + ': ' + str(rate) for name, rate in tls.Ls.items())
        lines.append('__________________')
        return '\n'.join(lines)

    def model_description_dict(self) -> dict:
        """JSON-compatible description, keys as reference basic_model.py:462-495."""
        out = {}
        for k, tls in enumerate(self.TLSs):
            out[str(k)] = {
                'type': 'Qubit' if tls.is_qubit else 'Defect',
                'energy': tls.energy,
                'couplings:': {str(self.TLSs.index(p)): lst for p, lst in tls.couplings.items()},
                'Ls:': dict(tls.Ls),
            }
        return out

    # ------------------------------------------------------------------ random population
    def random_setup(self, qubits_number: int = 1, defects_number_range: tuple = (2, 2),
                     qubits_energies: list = [5], defects_energies_range: tuple = (3, 7),
                     allowed_couplings: dict = {'qubit': [(1, (0.1, 1), [('sigmax', 'sigmax')])]},
                     allowed_Ls: dict = {'sigmaz': (1, (0.005, 0.1))}) -> None:
        """
        Replaces the configuration by one qubit plus a random number of defects
        (reference basic_model.py:499-571; same np.random call order).
        """
        self.TLSs, self.TLS_labels = [], {}

        def draw_Ls():
            return {name: np.random.uniform(*rng) for name, (prob, rng) in allowed_Ls.items()
                    if np.random.uniform() < prob}

        def draw_couplings():
            return {partner: [(np.random.uniform(*rng), kind) for (prob, rng, kind) in options
                              if np.random.uniform() < prob]
                    for partner, options in allowed_couplings.items()}

        self.add_TLS(TLS_label='qubit', is_qubit=True, energy=qubits_energies[0], couplings={}, Ls=draw_Ls())
        for _ in range(np.random.randint(defects_number_range[0], defects_number_range[1] + 1)):
            self.add_TLS(is_qubit=False, energy=np.random.uniform(*defects_energies_range),
                         couplings=draw_couplings(), Ls=draw_Ls())
        self.build_operators()

    # ------------------------------------------------------------------ fixed-length vectorisation
    def vectorise_under_library(self, hyperparameters: dict, model=False):
        """
        (values, labels, latex labels) with one entry per process the libraries allow:
        defect energies; qubit-defect couplings (pair-major, then library order); defect-defect
        couplings; qubit Ls; defect Ls.  Absent processes are 0.0.  Layout and label grammar of
        reference basic_model.py:575-765 (this is also the device parameter-vector layout).
        """
        model = model or self
        qubits = [t for t in model.TLSs if t.is_qubit]
        defects = [t for t in model.TLSs if not t.is_qubit]
        qn = {id(t): 'S' + str(i + 1) for i, t in enumerate(qubits)}
        dn = {id(t): 'V' + str(i + 1) for i, t in enumerate(defects)}
        values, labels, latex = [], [], []

        for t in defects:
            labels.append(dn[id(t)] + '-E-')
            latex.append(dn[id(t)] + r'$\ E$')
            values.append(t.energy)

        def coupling_block(pairs, names, library):
            for (t1, t2) in pairs:
                where = names[id(t1)] + ',' + names[id(t2)]
                for kind in list(library.keys()):
                    short = ''.join('-' + _SHORT[a][0] + ',' + _SHORT[b][0] for a, b in kind)
                    tex = ' ' + '+'.join(_SHORT[a][1] + r'$\otimes$' + _SHORT[b][1] for a, b in kind)
                    total = float(0)
                    for holder, partner in ((t1, t2), (t2, t1)):
                        for strength, op_pairs in holder.couplings.get(partner, []):
                            if set(op_pairs) == set(kind):
                                total += strength
                    labels.append(where + short)
                    latex.append(where + tex)
                    values.append(total)

        coupling_block([(q, d) for q in qubits for d in defects], {**qn, **dn},
                       hyperparameters['qubit2defect_couplings_library'])
        coupling_block([(d1, d2) for i, d1 in enumerate(defects) for d2 in defects[i + 1:]], dn,
                       hyperparameters['defect2defect_couplings_library'])

        for group, names, lib in ((qubits, qn, 'qubit_Ls_library'), (defects, dn, 'defect_Ls_library')):
            for t in group:
                for name in list(hyperparameters[lib].keys()):
                    labels.append(names[id(t)] + '-' + _SHORT[name][0] + '-')
                    latex.append(names[id(t)] + r'$\ $' + _SHORT[name][1])
                    values.append(t.Ls[name] if name in t.Ls else float(0))
        return values, labels, latex
