"""
`ProcessHandler`: reversible-jump moves that add or remove one process (a single-site Lindblad
operator or a pairwise coupling) -- the API of reference process_handling.py:26-694.

Every move returns (model, number of possible modifications of that type); with update=False
the model is left untouched (the count feeds the proposal probabilities).  np.random call
order per move is the reference's: one np.random.choice over the candidates, then (additions)
one np.random.uniform over the class bounds, or np.random.gamma(shape, scale) without bounds.

Deliberate difference: the reference decides which partner of a new coupling
"holds" it by iterating a Python set of objects, which depends on memory addresses.  Here the
holder is always the first of the pair in pair order (the qubit; the earlier defect).
"""
from __future__ import annotations

import copy

import numpy as np


def _as_pairs(kind) -> tuple:
    """('x','y') or (('x','y'),...) -> tuple of operator-pair tuples."""
    if type(kind) == tuple and list(map(type, kind)) == [str, str]:
        return (kind,)
    return tuple(kind)


class ProcessHandler:

    def __init__(self, chain=False, model=False, qubit_Ls_library=False, defect_Ls_library=False,
                 qubit2defect_couplings_library=False, defect2defect_couplings_library=False, bounds=False):
        self.bounds = bounds
        self.qubit_Ls_library = qubit_Ls_library
        self.initial_qubit_Ls_library = copy.deepcopy(qubit_Ls_library)
        self.defect_Ls_library = defect_Ls_library
        self.initial_defect_Ls_library = copy.deepcopy(defect_Ls_library)
        self.qubit2defect_couplings_library = qubit2defect_couplings_library
        self.defect2defect_couplings_library = defect2defect_couplings_library
        self.model = model
        self.chain = chain
        # optional callable(first, second) -> bool deciding whether the FIRST of the pair holds a new
        # coupling; None = always the first.  Only used to replay reference trajectories, whose holder
        # choice is address-dependent (see module docstring).
        self.coupling_holder_choice = None

    # ------------------------------------------------------------------ Lindblad processes
    def _library(self, passed, attribute, what):
        if passed:
            return passed
        held = getattr(self, attribute)
        if isinstance(held, dict):
            return held
        raise RuntimeError(what)

    def _new_value(self, parameter_class, shape_scale):
        if not self.bounds:
            return np.random.gamma(*shape_scale)
        return np.random.uniform(*self.bounds[parameter_class])

    def _add_L(self, model, library, qubits: bool, update: bool):
        candidates = [(tls, name) for tls in model.TLSs if tls.is_qubit == qubits
                      for name in library if name not in tls.Ls]
        if candidates and update:
            tls, name = candidates[np.random.choice(len(candidates))]
            tls.Ls[name] = self._new_value('Ls', library[name])
            model.build_operators()
        return model, len(candidates)

    def _remove_L(self, model, qubits: bool, update: bool):
        candidates = [(tls, name) for tls in model.TLSs if tls.is_qubit == qubits for name in tls.Ls]
        if candidates and update:
            tls, name = candidates[np.random.choice(len(candidates))]
            tls.Ls.pop(name)
            model.build_operators()
        return model, len(candidates)

    def add_random_qubit_L(self, model, *, qubit_Ls_library=False, update: bool = True):
        lib = self._library(qubit_Ls_library, 'qubit_Ls_library',
                            'Cannot add qubit Lindblad process as process library not specified')
        return self._add_L(model, lib, True, update)

    def add_random_defect_L(self, model, *, defect_Ls_library=False, update: bool = True):
        lib = self._library(defect_Ls_library, 'defect_Ls_library',
                            'Cannot add defect Lindblad process as process library not specified')
        return self._add_L(model, lib, False, update)

    def remove_random_qubit_L(self, model, update: bool = True):
        return self._remove_L(model, True, update)

    def remove_random_defect_L(self, model, update: bool = True):
        return self._remove_L(model, False, update)

    # ------------------------------------------------------------------ couplings
    @staticmethod
    def _pairs(model, qubit_defect: bool):
        if qubit_defect:
            return [(q, d) for q in model.TLSs if q.is_qubit for d in model.TLSs if not d.is_qubit]
        defects = [t for t in model.TLSs if not t.is_qubit]
        return [(d1, d2) for i, d1 in enumerate(defects) for d2 in defects[i + 1:]]

    def _add_coupling(self, model, library, qubit_defect: bool, update: bool):
        pairs = self._pairs(model, qubit_defect)
        # a coupling type is "present" on a pair if either partner holds the same SET of operator pairs
        present = set()
        for a, b in pairs:
            for holder, partner in ((a, b), (b, a)):
                for _, op_pairs in holder.couplings.get(partner, []):
                    present.add((id(a), id(b), frozenset(tuple(p) for p in op_pairs)))
        candidates = [(a, b, _as_pairs(kind), dist) for a, b in pairs for kind, dist in library.items()
                      if (id(a), id(b), frozenset(_as_pairs(kind))) not in present]
        if candidates and update:
            holder, partner, op_pairs, dist = candidates[np.random.choice(len(candidates))]
            strength = self._new_value('couplings', dist)
            if self.coupling_holder_choice is not None and not self.coupling_holder_choice(holder, partner):
                holder, partner = partner, holder
            holder.couplings.setdefault(partner, []).append((strength, list(op_pairs)))
            model.build_operators()
        return model, len(candidates)

    @staticmethod
    def _held_couplings(model, qubit_defect: bool):
        wanted = 1 if qubit_defect else 0
        return [(tls, partner, i) for tls in model.TLSs for partner in tls.couplings
                if int(tls.is_qubit) + int(partner.is_qubit) == wanted
                for i in range(len(tls.couplings[partner]))]

    def _remove_coupling(self, model, qubit_defect: bool, update: bool):
        candidates = self._held_couplings(model, qubit_defect)
        if candidates and update:
            tls, partner, i = candidates[np.random.choice(len(candidates))]
            tls.couplings[partner].pop(i)
            if not tls.couplings[partner]:
                tls.couplings.pop(partner)
            model.build_operators()
        return model, len(candidates)

    def add_random_qubit2defect_coupling(self, model, *, qubit2defect_couplings_library=False, update: bool = True):
        lib = self._library(qubit2defect_couplings_library, 'qubit2defect_couplings_library',
                            'Cannot add qubit-random defect coupling as library not specified')
        return self._add_coupling(model, lib, True, update)

    def add_random_defect2defect_coupling(self, model, *, defect2defect_couplings_library=False, update: bool = True):
        lib = self._library(defect2defect_couplings_library, 'defect2defect_couplings_library',
                            'Cannot add random defects coupling as library not specified')
        return self._add_coupling(model, lib, False, update)

    def remove_random_qubit2defect_coupling(self, model, update: bool = True):
        return self._remove_coupling(model, True, update)

    def remove_random_defect2defect_coupling(self, model, update: bool = True):
        return self._remove_coupling(model, False, update)

    def count_moves(self, model):
        """
        {move method name: number of possible modifications} for all twelve add / remove moves in one pass -- the counts
        the methods return with update=False.  None when a library is not a dict (the methods raise for that).
        """
        libs = (self.qubit_Ls_library, self.defect_Ls_library, self.qubit2defect_couplings_library,
                self.defect2defect_couplings_library)
        if not all(isinstance(lib, dict) for lib in libs):
            return None
        qL, dL, qd, dd = libs
        n = dict.fromkeys(('add_random_qubit_L', 'add_random_defect_L', 'remove_random_qubit_L', 'remove_random_defect_L',
                           'add_random_qubit2defect_coupling', 'remove_random_qubit2defect_coupling',
                           'add_random_defect2defect_coupling', 'remove_random_defect2defect_coupling'), 0)
        for tls in model.TLSs:
            if tls.is_qubit:
                n['add_random_qubit_L'] += sum(1 for name in qL if name not in tls.Ls)
                n['remove_random_qubit_L'] += len(tls.Ls)
            else:
                n['add_random_defect_L'] += sum(1 for name in dL if name not in tls.Ls)
                n['remove_random_defect_L'] += len(tls.Ls)
            for partner, held in tls.couplings.items():
                kind = int(tls.is_qubit) + int(partner.is_qubit)
                if kind == 1:
                    n['remove_random_qubit2defect_coupling'] += len(held)
                elif kind == 0:
                    n['remove_random_defect2defect_coupling'] += len(held)
        for qubit_defect, library, key in ((True, qd, 'add_random_qubit2defect_coupling'),
                                           (False, dd, 'add_random_defect2defect_coupling')):
            kinds = [frozenset(_as_pairs(kind)) for kind in library]
            for a, b in self._pairs(model, qubit_defect):
                held = a.couplings.get(b, []) + b.couplings.get(a, [])
                if not held:
                    n[key] += len(kinds)
                    continue
                present = {frozenset(tuple(p) for p in op_pairs) for _, op_pairs in held}
                n[key] += sum(1 for kind in kinds if kind not in present)
        return n

    # ------------------------------------------------------------------ libraries
    def define_qubit_Ls_library(self, qubit_Ls_library) -> None:
        self.qubit_Ls_library = qubit_Ls_library

    def reset_qubit_Ls_library(self) -> None:
        self.qubit_Ls_library = copy.deepcopy(self.initial_qubit_Ls_library)

    def define_defect_Ls_library(self, defect_Ls_library) -> None:
        self.defect_Ls_library = defect_Ls_library

    def reset_defect_Ls_library(self) -> None:
        self.defect_Ls_library = copy.deepcopy(self.initial_defect_Ls_library)

    def rescale_Ls_library_range(self, factor) -> None:
        raise RuntimeError('Unmaintained legacy process handler method rescale_Ls_library_range() called')

    # ------------------------------------------------------------------ counting / filtering
    def count_Ls(self, model) -> int:
        return sum(len(tls.Ls) for tls in model.TLSs)

    def count_defect2defect_couplings(self, model) -> int:
        return len(self._held_couplings(model, False))

    def count_qubit2defect_couplings(self, model) -> int:
        return len(self._held_couplings(model, True))

    def filter_params(self, model, thresholds: dict):
        """
        Drops every Lindblad process / coupling whose |value| is below its class threshold
        (reference process_handling.py:656-694; the reference pops list indices while
        iterating, which can skip entries -- here all sub-threshold couplings are removed).
        """
        for tls in model.TLSs:
            for name in [n for n, rate in tls.Ls.items() if abs(rate) < thresholds['Ls']]:
                tls.Ls.pop(name)
            for partner in list(tls.couplings):
                kept = [c for c in tls.couplings[partner] if not abs(c[0]) < thresholds['couplings']]
                if kept:
                    tls.couplings[partner][:] = kept
                else:
                    tls.couplings.pop(partner)
        model.build_operators()
        return model
