"""
`LearningModel`: a BasicModel that can perturb its own parameters and (re)configure itself
from a parameter vector -- the API of the reference's learning_model.py (change_params
:112-194, configure_to_params_vector :198-277), with the same np.random call order so that a
seeded chain draws the same numbers.
"""
from __future__ import annotations

import copy

import numpy as np

from . import basic_model

_MAX_ATTEMPTS = 10
_IMMUTABLE = (bool, int, float, str, type(None))
_SHORT2LONG = {'sx': 'sigmax', 'sy': 'sigmay', 'sz': 'sigmaz', 'sp': 'sigmap', 'sm': 'sigmam'}


def _jitter(value, width, low, high):
    """
    Up to ten Gaussian proposals around `value`; first one inside [low, high] wins, else the
    old value stays (rejection sampling of reference learning_model.py:144-191).
    """
    for _ in range(_MAX_ATTEMPTS):
        candidate = value + np.random.normal(0, width)
        if low <= candidate <= high:
            return candidate
    return value


class LearningModel(basic_model.BasicModel):

    def __init__(self, *args, jump_lengths: dict = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.jump_lengths = jump_lengths   # {'energies' | 'couplings' | 'Ls': std of the Gaussian jump}
        self.final_loss = False            # set by LearningChain.run on the best model

    def __deepcopy__(self, memo):
        """Structural copy (TLS records, couplings re-keyed to the new partners); numpy states shared-by-copy."""
        new = type(self).__new__(type(self))
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in ('TLSs', 'TLS_labels', '_cache'):
                if type(v) in _IMMUTABLE:
                    new.__dict__[k] = v
                elif type(v) is dict and all(type(a) in _IMMUTABLE and type(b) in _IMMUTABLE for a, b in v.items()):
                    new.__dict__[k] = dict(v)      # flat dictionaries of numbers (jump_lengths)
                else:
                    new.__dict__[k] = copy.deepcopy(v, memo)
        new._cache = {}
        new.TLSs = []
        twin = {}
        for t in self.TLSs:
            c = type(t)(new, t.TLS_id, t.is_qubit, t.energy, {}, dict(t.Ls),
                        t.initial_state if isinstance(t.initial_state, bool) or t.initial_state is None
                        else np.array(t.initial_state))
            twin[id(t)] = c
            memo[id(t)] = c
            new.TLSs.append(c)
        for t in self.TLSs:
            for partner, lst in t.couplings.items():
                twin[id(t)].couplings[twin[id(partner)]] = [(s, list(p)) for s, p in lst]
        new.TLS_labels = {label: twin[id(t)] for label, t in self.TLS_labels.items() if id(t) in twin}
        return new

    def set_jump_lengths(self, passed_jump_lengths: dict = False) -> None:
        self.jump_lengths = passed_jump_lengths

    def _resolve_jump_lengths(self, passed):
        if isinstance(passed, bool) and not passed:
            if isinstance(self.jump_lengths, bool) and not self.jump_lengths:
                raise RuntimeError('need to specify jump lengths for the operators present')
        else:
            self.jump_lengths = passed

    def old_change_params(self, passed_jump_lengths: dict = False, bounds: dict = False) -> None:
        """Legacy unbounded variant (reference learning_model.py:58-108): only Ls are kept non-negative."""
        self._resolve_jump_lengths(passed_jump_lengths)
        jl = self.jump_lengths
        for tls in self.TLSs:
            if not tls.is_qubit:
                tls.energy += np.random.normal(0, jl['energies'])
            for partner in tls.couplings:
                tls.couplings[partner] = [(s + np.random.normal(0, jl['couplings']), p)
                                          for s, p in tls.couplings[partner]]
            for name in tls.Ls:
                tls.Ls[name] = _jitter(tls.Ls[name], jl['Ls'], 0, np.inf)
        self.build_operators()

    def change_params(self, passed_jump_lengths: dict = False, bounds: dict = False) -> None:
        """
        Gaussian jump on every existing parameter except qubit energies, each redrawn up to ten
        times until inside its class bounds.  Draw order: per TLS -- defect energy, couplings it
        holds (partner order, then list order), Lindblad rates (dict order).
        """
        self._resolve_jump_lengths(passed_jump_lengths)
        jl = self.jump_lengths
        if not bounds:
            bounds = {'energies': (-np.inf, np.inf), 'couplings': (-np.inf, np.inf), 'Ls': (0, np.inf)}
        for tls in self.TLSs:
            if not tls.is_qubit:
                tls.energy = _jitter(tls.energy, jl['energies'], *bounds['energies'])
            for partner in tls.couplings:
                held = tls.couplings[partner]
                for i, (strength, op_pairs) in enumerate(held):
                    held[i] = (_jitter(strength, jl['couplings'], *bounds['couplings']), op_pairs)
            for name in tls.Ls:
                tls.Ls[name] = _jitter(tls.Ls[name], jl['Ls'], *bounds['Ls'])
        self.build_operators()

    def configure_to_params_vector(self, vectorised_model, D: int = 2, default_qubit_energy=1,
                                   qubit_initial_state=False, defect_initial_state=False):
        """
        Empties this model and repopulates it (one qubit, D defects) from
        (values, labels, latex labels) as produced by vectorise_under_library.
        Label grammar (reference learning_model.py:198-277): 'V2-E-' energy; 'S1,V2-sx,sx-sy,sy'
        coupling held by the first system; 'V1-sz-' Lindblad rate.  Zero values are skipped.
        """
        vals, labels, _ = vectorised_model
        self.TLSs = []
        self.add_TLS(is_qubit=True, energy=default_qubit_energy, initial_state=qubit_initial_state)
        for _ in range(D):
            self.add_TLS(is_qubit=False, initial_state=defect_initial_state)
        group = {'S': [t for t in self.TLSs if t.is_qubit], 'V': [t for t in self.TLSs if not t.is_qubit]}

        def system(tag):
            return group[tag[0]][int(tag[1]) - 1]

        for label, val in zip(labels, vals):
            if val == 0:
                continue
            n_systems = sum(ch in 'SV' for ch in label)
            if n_systems == 2:
                where, ops_all = label.split('-', 1)
                holder, partner = (system(tag) for tag in where.split(','))
                pairs = [tuple(_SHORT2LONG[o] for o in pair.split(',')) for pair in ops_all.split('-')]
                holder.couplings.setdefault(partner, []).append((val, pairs))
            elif n_systems == 1 and 'E' in label:
                system(label.split('-', 1)[0]).energy = val
            else:
                tag, op = label.split('-', 2)[0:2]
                system(tag).Ls[_SHORT2LONG[op]] = val
        self.build_operators()
        return self
