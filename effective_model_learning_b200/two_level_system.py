"""
Per-site record of a model (reference two_level_system.py:13-58).

couplings: {partner TwoLevelSystem: [(strength, [(op_on_self, op_on_partner), ...]), ...]}
           stored on ONE partner of each pair only.
Ls:        {operator name: rate}
initial_state: 2x2 density matrix (numpy) or False for the model default.
"""
from __future__ import annotations


class TwoLevelSystem:

    def __init__(self, model, TLS_id: str = "", is_qubit: bool = False, energy=None,
                 couplings: dict = None, Ls: dict = None, initial_state=False):
        self.model = model
        self.TLS_id = TLS_id
        self.is_qubit = is_qubit
        self.energy = energy
        self.couplings = {} if couplings is None else couplings
        self.Ls = {} if Ls is None else Ls
        self.initial_state = initial_state

    def has_initial_state(self) -> bool:
        s = self.initial_state
        return not (s is None or (isinstance(s, bool) and not s))
