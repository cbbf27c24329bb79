"""
`ParamsHandler`: the parameter-tweak move of the chain (reference params_handling.py:19-268).
Holds the current jump lengths and bounds and applies LearningModel.change_params.
"""
from __future__ import annotations

import copy

import numpy as np

from . import learning_model

_DEFAULT_JUMPS = {'couplings': 0.001, 'energies': 0.01, 'Ls': 0.00001}
_DEFAULTS = {'max_optimisation_steps': int(1e3), 'MH_acceptance': False, 'MH_temperature': 0.1,
             'initial_jump_lengths': _DEFAULT_JUMPS, 'jump_annealing_rate': 0,
             'acceptance_window': 200, 'acceptance_target': 0.4}


class ParamsHandler:

    def __init__(self, chain, hyperparams: dict = False, bounds: dict = False) -> None:
        self.config_done = False
        self.set_hyperparams(hyperparams if hyperparams else [])
        self.chain = chain
        self.bounds = bounds      # False: only Ls >= 0 is enforced

    def set_hyperparams(self, hyperparams: dict) -> None:
        """Every known key becomes an attribute (default when absent); jump lengths restart from the initial ones."""
        self.hyperparams_init_output = {}
        for key, default in _DEFAULTS.items():
            setattr(self, key, hyperparams[key] if key in hyperparams else default)
            self.hyperparams_init_output[key] = getattr(self, key)
        self.jump_lengths = copy.deepcopy(self.initial_jump_lengths)
        self.config_done = True

    def set_bounds(self, bounds: dict) -> None:
        self.bounds = bounds

    def set_jump_lengths(self, jump_lengths: dict) -> None:
        self.jump_lengths = copy.deepcopy(jump_lengths)

    def output_hyperparams_init(self) -> dict:
        return self.hyperparams_init_output

    def output_hyperparams_curr(self) -> dict:
        return {key: getattr(self, key) for key in self.hyperparams_init_output}

    def tweak_all_parameters(self, model):
        """Jitters every existing parameter of `model` in place and returns it."""
        if not isinstance(model, learning_model.LearningModel):
            raise RuntimeError('Model passed as argument needs to be instance of LearningModel\n'
                               + 'to automatically tweak all parameters!')
        if not self.config_done:
            raise RuntimeError('Parameter handler hyperparameters need to be specified!')
        model.change_params(self.jump_lengths, bounds=self.bounds)
        return model

    def rescale_jump_lengths(self, factor) -> None:
        if not self.config_done:
            raise RuntimeError('Parameter handler hyperparameters not specified!')
        for key in self.jump_lengths:
            self.jump_lengths[key] = self.jump_lengths[key] * factor

    def do_optimisation(self, initial_model):
        """
        Stand-alone parameter optimisation (unused by the chain, unmaintained in the reference,
        params_handling.py:169-253): greedy / Metropolis walk with change_params proposals.
        Returns (best model, best cost, all costs).
        """
        if not self.config_done:
            raise RuntimeError('Parameter handler hyperparameters not specified!')
        cost = self.chain.total_dev
        current = copy.deepcopy(initial_model)
        current_cost = best_cost = cost(current)
        best = copy.deepcopy(current)
        costs = [current_cost]
        for _ in range(self.max_optimisation_steps):
            proposed = copy.deepcopy(current)
            proposed.change_params(passed_jump_lengths=self.jump_lengths)
            proposed_cost = cost(proposed)
            costs.append(proposed_cost)
            if self.jump_annealing_rate:
                self.rescale_jump_lengths(np.exp(-self.jump_annealing_rate))
            take = proposed_cost < current_cost
            if not take and self.MH_acceptance:
                likelihood = 0 if self.MH_temperature == 0 else np.exp(-(proposed_cost - current_cost) / self.MH_temperature)
                take = np.random.uniform() < likelihood
            if take:
                current, current_cost = proposed, proposed_cost
                if proposed_cost < best_cost:
                    best_cost, best = proposed_cost, copy.deepcopy(proposed)
        return best, best_cost, costs
