"""
`LearningChain`: one reversible-jump MCMC chain over TLS models given target data -- the API
and step-by-step semantics of the reference's learning_chain.py (class :27-997, run loop
:314-614), including its np.random consumption order, so that a seeded
chain proposes the same moves and, with losses equal to 1e-8, takes the same accept/reject
decisions.

What differs on purpose:
  * the loss is evaluated ONCE per proposal by the CUDA library (the reference evaluates the
    proposal twice and the current model once more inside acceptance_probability,
    learning_chain.py:559,782-783 -- SURVEY.md F6); values are identical because evaluation
    is deterministic;
  * with no custom_function_on_dynamics_return the observables never leave the device: the
    kernel reduces them against the targets and returns only the loss;
  * class Defaults.params_priors / params_bounds are plain dicts (the reference's trailing
    commas make them 1-tuples, learning_chain.py:119,125);
  * a model the kernels refuse to integrate (norm bound x time span >= 1e5, or a capped series) gets loss +inf,
    i.e. the proposal is rejected with a RuntimeWarning, where the reference's mesolve (nsteps = 1e9) would grind
    through it; such models are counted in `unconverged_evaluations`.
"""
from __future__ import annotations

import copy
import time
import warnings

import numpy as np
import scipy as sp
import scipy.special
import scipy.stats

from . import _native
from . import engine
from . import learning_model
from . import params_handling
from . import process_handling

TYPE_MODEL = learning_model.LearningModel

_REVERSE = {
    'tweak all parameters': 'tweak all parameters',
    'add qubit L': 'remove qubit L', 'remove qubit L': 'add qubit L',
    'add defect L': 'remove defect L', 'remove defect L': 'add defect L',
    'add qubit-defect coupling': 'remove qubit-defect coupling',
    'remove qubit-defect coupling': 'add qubit-defect coupling',
    'add defect-defect coupling': 'remove defect-defect coupling',
    'remove defect-defect coupling': 'add defect-defect coupling',
}
_JUMP_METHODS = {
    'add qubit L': 'add_random_qubit_L', 'add defect L': 'add_random_defect_L',
    'remove qubit L': 'remove_random_qubit_L', 'remove defect L': 'remove_random_defect_L',
    'add qubit-defect coupling': 'add_random_qubit2defect_coupling',
    'remove qubit-defect coupling': 'remove_random_qubit2defect_coupling',
    'add defect-defect coupling': 'add_random_defect2defect_coupling',
    'remove defect-defect coupling': 'remove_random_defect2defect_coupling',
}
_NOT_HYPERPARAMS = ('self', 'initial', 'target_times', 'target_datasets', 'target_observables',
                    'iterations_till_progress_update')


_erf, _SQRT2 = sp.special.erf, np.sqrt(2)


def _weighted_choice(p) -> int:
    """
    Index drawn with probabilities p -- what np.random.choice(len(p), p=p) returns, from the same single draw of the
    global stream (numpy's legacy RandomState.choice: cdf = p.cumsum(); cdf /= cdf[-1]; searchsorted(random_sample(),
    side='right')), without its per-call validation of p and conversion of the label list (about a fifth of a chain
    step's host time).  tests/test_host.py compares the two draw for draw.
    """
    cdf = np.array(p, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(np.random.random_sample(), side='right'))


def _Phi(y):
    return 1 / 2 * (1 + _erf(y / _SQRT2))


class LearningChain:

    class Defaults:
        initial = False                  # LearningModel, or (qubit_energy, defects_number), or defects_number
        qubit_initial_state = False      # 2x2 density matrix
        defect_initial_state = False
        max_chain_steps = 10000
        chain_step_options = {'tweak all parameters': 10, **{move: 1 for move in _JUMP_METHODS}}
        temperature_proposal = 0.0005    # value, or (shape, scale) of a gamma
        jump_length_rescaling_factor = 1
        complexity_factor = 1
        acceptance_window = 10
        acceptance_target = 0.4
        acceptance_band = 0.2
        shock_anneal_at = False
        params_handler_hyperparams = {'initial_jump_lengths': {'couplings': 0.1, 'energies': 0.1, 'Ls': 0.01}}
        qubit_Ls_library = {op: (2, 0.03) for op in ('sigmax', 'sigmay', 'sigmaz')}
        defect_Ls_library = {op: (2, 0.03) for op in ('sigmax', 'sigmay', 'sigmaz')}
        qubit2defect_couplings_library = {((op, op),): (2, 0.3) for op in ('sigmax', 'sigmay', 'sigmaz')}
        defect2defect_couplings_library = {((op, 'sigmay'),): (2, 0.3) for op in ('sigmax', 'sigmay', 'sigmaz')}
        params_thresholds = {'Ls': 1e-7, 'couplings': 1e-6}
        params_priors = {'couplings': (1.04, 30), 'energies': (1.05, 35), 'Ls': (1.004, 23)}
        params_bounds = {'couplings': (0.1, 10), 'energies': (0.01, 1), 'Ls': (0.01, 1)}
        custom_function_on_dynamics_return = False
        iterations_till_progress_update = False
        store_all_proposals = False

    def complementary_step(self, step_type: str) -> str:
        """The move type that undoes `step_type`."""
        try:
            return _REVERSE[step_type]
        except KeyError:
            raise RuntimeError('Complementary step type could not be determined'
                               + 'due to step type not recognised')

    def __init__(self, target_times, target_datasets, target_observables, *,
                 initial=False, qubit_initial_state=False, defect_initial_state=False,
                 max_chain_steps: int = False, chain_step_options: dict = False,
                 temperature_proposal=False, shock_anneal_at: int = False, complexity_factor=1,
                 jump_length_rescaling_factor: float = False,
                 acceptance_window: float = False, acceptance_target: float = False, acceptance_band: float = False,
                 params_handler_hyperparams: dict = False,
                 qubit_Ls_library: dict = False, defect_Ls_library: dict = False,
                 qubit2defect_couplings_library: dict = False, defect2defect_couplings_library: dict = False,
                 params_thresholds: dict = False, params_priors: dict = False, params_bounds: dict = False,
                 custom_function_on_dynamics_return=False, iterations_till_progress_update: int = False,
                 store_all_proposals: bool = False):
        passed = copy.deepcopy({k: v for k, v in locals().items() if k != 'self'})
        self.init_hyperparams = {}
        # False (or not passed) -> class default; everything else kept as passed
        for key, val in passed.items():
            if isinstance(val, bool) and not val:
                val = getattr(self.Defaults, key, False)
            setattr(self, key, val)
            if key not in _NOT_HYPERPARAMS:
                self.init_hyperparams[key] = val

        kind = type(self.initial)
        if kind == learning_model.LearningModel:
            pass
        elif kind in (tuple, list) and len(self.initial) == 2:
            self.initial = self.make_initial_model(initial[0], initial[1])
        elif kind == int:
            self.initial = self.make_initial_model(1, initial)
        else:
            raise RuntimeWarning('initial model must be specified:\neither as instance of LearningModel,'
                                 + '\nor tuple or list of (qubit energy, number of defects),'
                                 + '\n or integer number of defects, assuming qubit energy = 1')

        total = 0
        for option, priority in self.chain_step_options.items():
            if type(priority) not in (int, float):
                raise RuntimeError('Chain step options need to have relative priorities\n'
                                   + 'specified by a single number')
            total += priority
        self.next_step_labels = list(self.chain_step_options)
        self.next_step_priorities_dict = {o: self.chain_step_options[o] / total for o in self.next_step_labels}
        self.next_step_priorities_list = [self.next_step_priorities_dict[o] for o in self.next_step_labels]

        self.params_handler = None
        self.process_handler = None
        self._contexts = {}

        self.explored_proposals = []
        self.explored_loss = []
        self.explored_acceptance_probability = []
        self.explored_log_posterior = []
        self.explored_log_likelihood_prior = []
        self.current = copy.deepcopy(self.initial)
        self.best = copy.deepcopy(self.current)
        self.chain_windows_acceptance_log = []

        self.MH_temperature = self.sample_T()
        self.initialise_process_handler()
        self.process_handler.filter_params(self.current, self.params_thresholds)
        self.current_loss = self.total_dev(self.current)
        self.best_loss = self.current_loss
        self._log_explored(self.current, self.current_loss)
        if self.store_all_proposals:
            self.explored_proposals.append(copy.deepcopy(self.initial))

        self.tot_RJ_steps = 0
        self.acc_RJ_steps = 0
        self.tot_tweak_steps = 0
        self.acc_tweak_steps = 0

    # ------------------------------------------------------------------ bookkeeping helpers
    def _log_explored(self, model, loss):
        minus_log_prior = self.prior(model, return_minus_log_of=True)
        self.explored_loss.append(loss)
        self.explored_log_posterior.append(-(loss / self.MH_temperature + minus_log_prior))
        self.explored_log_likelihood_prior.append((-loss / self.MH_temperature, -minus_log_prior))

    def _move_weights(self, model):
        """priority x number of ways, per move type, in label order."""
        if not self.process_handler:
            self.initialise_process_handler()
        if not self.params_handler:      # (as step() does)
            self.initialise_params_handler()
        # every jump move counted in one pass over the model (the same numbers step(model, x, update=False) returns one
        # move at a time); None: a library is missing -- the per-move path raises the reference's error for it
        ways = self.process_handler.count_moves(model)
        if ways is None:
            return [self.next_step_priorities_dict[x] * self.step(model, x, update=False)[1]
                    for x in self.next_step_labels]
        priorities = self.next_step_priorities_dict
        return [priorities[x] * (1 if x == 'tweak all parameters' else ways[_JUMP_METHODS[x]] if x in _JUMP_METHODS
                                 else self.step(model, x, update=False)[1])
                for x in self.next_step_labels]

    def _xi_product(self, model):
        """Product of truncated-Gaussian normalisers over every tweakable parameter of `model`."""
        if not self.params_handler:
            self.initialise_params_handler()
        jl, bounds = self.params_handler.jump_lengths, self.params_bounds
        # one vectorised evaluation of xi() over all parameters (the same IEEE operations element by element), then the
        # product in the reference's parameter order
        m, s, a, b = [], [], [], []
        for tls in model.TLSs:
            if not tls.is_qubit:
                m.append(tls.energy); s.append(jl['energies']); a.append(bounds['energies'][0]); b.append(bounds['energies'][1])
            for rate in tls.Ls.values():
                m.append(rate); s.append(jl['Ls']); a.append(bounds['Ls'][0]); b.append(bounds['Ls'][1])
            for partner in tls.couplings:
                for coupling in tls.couplings[partner]:
                    m.append(coupling[0]); s.append(jl['couplings'])
                    a.append(bounds['couplings'][0]); b.append(bounds['couplings'][1])
        out = 1
        if m:
            m, s = np.array(m, dtype=np.float64), np.array(s, dtype=np.float64)
            for v in self.xi(m, s, np.array(a, dtype=np.float64), np.array(b, dtype=np.float64)).tolist():
                out *= v
        return out

    def _gamma_prior_density(self, model):
        out = 1
        for tls in model.TLSs:
            out *= sp.stats.gamma.pdf(tls.energy, a=self.params_priors['energies'][0], scale=self.params_priors['energies'][1])
            for rate in tls.Ls.values():
                out *= sp.stats.gamma.pdf(rate, a=self.params_priors['Ls'][0], scale=self.params_priors['Ls'][1])
            for partner in tls.couplings:
                for coupling in tls.couplings[partner]:
                    out *= sp.stats.gamma.pdf(coupling[0], a=self.params_priors['couplings'][0],
                                              scale=self.params_priors['couplings'][1])
        return out

    # ------------------------------------------------------------------ the chain
    def run(self, steps: int = False):
        """
        Performs `steps`+1 proposals (the reference's `while i <= steps`) starting from the current
        model; appends to the instance trackers; returns the best model found.
        """
        if not steps:
            steps = self.max_chain_steps
        k = 0
        self.run_acceptance_tracker = []
        self.run_annealing_tracker = []
        self.run_step_type_tracker = []
        time_last = time.time()
        k2 = 0
        i = 0
        now_annealed = False
        # Two quantities of the CURRENT model are reused between steps instead of being recomputed from an unchanged model
        # (same values, so the same trajectory): its move weights (valid until a jump move is accepted) and its product of
        # truncated-Gaussian normalisers (valid until a proposal is accepted or the jump lengths change).  Local to this call:
        # whatever happens to self.current between calls of run() is seen.
        cur_weights = cur_xi = None
        while i <= steps:
            if bool(self.shock_anneal_at) and i == self.shock_anneal_at:
                # switch to the annealed jump lengths and restart from the best model so far
                self.params_handler.set_jump_lengths(self.params_handler_hyperparams['annealed_jump_lengths'])
                now_annealed = True
                i += 1
                self.current = copy.deepcopy(self.best)
                self.current_loss = self.best_loss
                cur_weights = cur_xi = None
                self.run_acceptance_tracker.append(True)
                if self.store_all_proposals:
                    self.explored_proposals.append(copy.deepcopy(self.current))
                self._log_explored(self.current, self.current_loss)
                self.run_annealing_tracker.append(now_annealed)
                self.run_step_type_tracker.append('jump to best')

            self.MH_temperature = self.sample_T()

            if k >= self.acceptance_window:
                k = 0
                n = len(self.run_acceptance_tracker)
                acceptance_ratio = sum(self.run_acceptance_tracker[n - self.acceptance_window:n]) / self.acceptance_window
                self.chain_windows_acceptance_log.append(acceptance_ratio)
                if acceptance_ratio - self.acceptance_target > self.acceptance_band:
                    self.cool_down()
                    cur_xi = None
                elif acceptance_ratio - self.acceptance_target < -self.acceptance_band:
                    self.heat_up()
                    cur_xi = None
            k += 1

            if bool(self.iterations_till_progress_update):
                if k2 >= self.iterations_till_progress_update:
                    k2 = 0
                    new_time = time.time()
                    print('\n\nIteration:\n' + str(i) + '\nElapsed (s):\n' + str(np.round(new_time - time_last, 2)),
                          flush=True)
                    time_last = new_time
                k2 += 1

            proposal = copy.deepcopy(self.current)

            if cur_weights is None:
                cur_weights = self._move_weights(proposal)      # (`proposal` is still a copy of the current model)
            weights = cur_weights
            norm = sum(weights)
            next_step = self.next_step_labels[_weighted_choice([w / norm for w in weights])]
            self.run_step_type_tracker.append(next_step)
            tweak = next_step == 'tweak all parameters'
            if tweak:
                self.tot_tweak_steps += 1
            else:
                self.tot_RJ_steps += 1

            proposal, _ = self.step(proposal, next_step, update=True)

            params_priors_ratio = 1
            if tweak:
                if not self.params_bounds:
                    # only L >= 0 is enforced: half-line truncated Gaussians
                    width = self.params_handler.jump_lengths['Ls']
                    p_there = p_back = 1
                    for tls in self.current.TLSs:
                        for rate in tls.Ls.values():
                            p_there *= 1 / (1 - _Phi(-rate / width))
                    for tls in proposal.TLSs:
                        for rate in tls.Ls.values():
                            p_back *= 1 / (1 - _Phi(-rate / width))
                    params_priors_ratio = self._gamma_prior_density(proposal) / self._gamma_prior_density(self.current)
                else:
                    p_there = self._xi_product(proposal)
                    if cur_xi is None:
                        cur_xi = self._xi_product(self.current)
                    p_back = cur_xi
            else:
                # (`weights` are the current model's: sum(weights) is what sum(self._move_weights(self.current)) gives)
                p_there = self.next_step_priorities_dict[next_step] / norm
                proposal_weights = self._move_weights(proposal)
                p_back = self.next_step_priorities_dict[self.complementary_step(next_step)] / sum(proposal_weights)

            proposal_loss = self.total_dev(proposal)
            self._log_explored(proposal, proposal_loss)

            acceptance_probability = self.acceptance_probability((self.current, self.current_loss),
                                                                 (proposal, proposal_loss),
                                                                 p_there, p_back, params_priors_ratio)
            self.explored_acceptance_probability.append(acceptance_probability)
            if np.random.uniform() < acceptance_probability:
                self.current = proposal
                self.current_loss = proposal_loss
                if proposal_loss < self.best_loss:
                    self.best_loss = proposal_loss
                    self.best = copy.deepcopy(proposal)
                self.run_acceptance_tracker.append(True)
                if self.store_all_proposals:
                    self.explored_proposals.append(copy.deepcopy(proposal))
                if tweak:
                    self.acc_tweak_steps += 1
                    cur_xi = p_there if self.params_bounds else None      # the proposal's product is the new current model's
                else:
                    self.acc_RJ_steps += 1
                    cur_weights, cur_xi = proposal_weights, None
            else:
                self.run_acceptance_tracker.append(False)
            self.run_annealing_tracker.append(now_annealed)
            i += 1

        if bool(self.iterations_till_progress_update):
            print('\n\nChain run completed.\n' + '_________________________\n\n', flush=True)
        self.best.final_loss = self.best_loss
        self.all_proposals = {'proposals': self.explored_proposals,
                              'loss': self.explored_loss,
                              'acceptance': self.run_acceptance_tracker,
                              'log_posterior': self.explored_log_posterior,
                              'log_likelihood_prior': self.explored_log_likelihood_prior,
                              'acceptance_probability': self.explored_acceptance_probability,
                              'annealed': self.run_annealing_tracker,
                              'step_types': self.run_step_type_tracker}
        return self.best

    def make_initial_model(self, qubit_energy, defects_number: int):
        """Qubit of the given energy plus `defects_number` defects at energy 0.1; no processes."""
        model = learning_model.LearningModel()
        model.add_TLS(TLS_label='qubit', is_qubit=True, initial_state=self.qubit_initial_state,
                      energy=qubit_energy, couplings={}, Ls={})
        for _ in range(defects_number):
            model.add_TLS(is_qubit=False, initial_state=self.defect_initial_state, energy=0.1, couplings={}, Ls={})
        model.build_operators()
        return model

    def step(self, model, step_type: str, update: bool = True):
        """(model, number of possible proposals of this type); modifies `model` only when update."""
        if not self.process_handler:
            self.initialise_process_handler()
        if not self.params_handler:
            self.initialise_params_handler()
        if step_type == 'tweak all parameters':
            return (self.params_handler.tweak_all_parameters(model), 1) if update else (model, 1)
        if step_type in _JUMP_METHODS:
            return getattr(self.process_handler, _JUMP_METHODS[step_type])(model, update=update)
        raise RuntimeError('Model proposal (chain step) option \'' + step_type + '\' not recognised')

    # ------------------------------------------------------------------ the hot path
    def _context(self, model) -> engine.EvalContext:
        """Device context of this chain for models shaped like `model` (targets resident on the GPU)."""
        states = [engine._initial_state_array(t) for t in model.TLSs]
        key = (len(model.TLSs), engine.first_qubit_site(model), b''.join(s.tobytes() for s in states))
        ctx = self._contexts.get(key)
        if ctx is None:
            q = engine.first_qubit_site(model)
            names = list(self.target_observables)
            ctx = engine.EvalContext(len(model.TLSs), states, q if q is not None else 0,
                                     names if q is not None else ['id2'] * len(names),
                                     np.asarray(self.target_times, dtype=np.float64))
            ctx.set_targets(np.stack([np.asarray(d, dtype=np.complex128) for d in self.target_datasets]))
            self._contexts[key] = ctx
        return ctx

    def total_dev(self, model) -> float:
        """
        sum over observables and times of |model - target|^2 / (T K)  (reference learning_chain.py:714-742),
        evaluated on the GPU.
        """
        if isinstance(self.target_datasets, np.ndarray) and isinstance(self.target_observables, str):
            self.target_datasets = [self.target_datasets]
            self.target_observables = [self.target_observables]
        if self.custom_function_on_dynamics_return:
            # a host callable post-processes the observables, so they have to come back first
            model_datasets = model.calculate_dynamics(evaluation_times=self.target_times,
                                                      observable_ops=self.target_observables,
                                                      custom_function_on_return=self.custom_function_on_dynamics_return)
            total_MSE = 0
            for i in range(len(model_datasets)):
                total_MSE += np.sum(np.square(abs(model_datasets[i] - self.target_datasets[i])))
            return total_MSE / (len(self.target_times) * len(model_datasets))
        if not getattr(model, '_built', False) or not model.TLSs:
            raise RuntimeError('Hamiltonian not defined -- cannot calculate dynamics')
        ctx = self._context(model)
        params = ctx.params_of([engine.model_processes(model)])
        try:
            _, loss = ctx.evaluate(params, want_expect=False, want_loss=True)
        except _native.NotConvergedError as err:
            # The kernels refuse a generator whose norm bound times the time span exceeds 1e5 (and report a capped
            # series); the reference's mesolve (nsteps = 1e9, basic_model.py:386) would grind through such a model.
            # Deviation: the proposal gets an infinite loss, i.e. it is rejected, and the chain carries on -- the same
            # outcome BatchedChains gives it (forced reject).  Counted in self.unconverged_evaluations.
            self.unconverged_evaluations = getattr(self, 'unconverged_evaluations', 0) + 1
            warnings.warn(f'model evaluation did not converge, loss set to +inf (proposal will be rejected): {err}',
                          RuntimeWarning, stacklevel=2)
            return float('inf')
        return float(loss[0])

    def acceptance_probability(self, current, proposal, there, back, params_priors_ratio: float = 1) -> float:
        """
        exp(-(loss_prop - loss_cur)/T) * prior(prop)/prior(cur) * back/there * params_priors_ratio.
        `current` / `proposal` are models or (model, loss) pairs (a bare model is evaluated here).
        """
        T = self.MH_temperature
        current, current_loss = self.process_argument_model(current)
        proposal, proposal_loss = self.process_argument_model(proposal)
        PP, PC = self.prior(proposal), self.prior(current)
        if abs(PP) < 1e-40 and abs(PC) < 1e-40:
            model_priors_ratio = 1
        elif abs(PC) < 1e-40:
            model_priors_ratio = np.inf
        else:
            model_priors_ratio = PP / PC
        return (np.exp(-1 / T * (proposal_loss - current_loss)) * model_priors_ratio * back / there
                * params_priors_ratio)

    def process_argument_model(self, arg):
        """Always (model, loss of model); a bare model has its loss computed."""
        if isinstance(arg, TYPE_MODEL):
            return (arg, self.total_dev(arg))
        if (type(arg) == tuple and len(arg) == 2 and isinstance(arg[0], TYPE_MODEL)
                and isinstance(arg[1], int | float)):
            return (arg[0], arg[1])
        raise RuntimeError('Learning chain\'s process_argument_model method failed: invalid input')

    def prior(self, model, return_minus_log_of: bool = False) -> float:
        """exp(-(number of processes)/complexity_factor), or its minus log."""
        # Lindblad processes + defect-defect + qubit-defect couplings (process_handler.count_*; qubit-qubit ones do not count)
        total_complexity = 0
        for tls in model.TLSs:
            total_complexity += len(tls.Ls)
            for partner, held in tls.couplings.items():
                if not (tls.is_qubit and partner.is_qubit):
                    total_complexity += len(held)
        if return_minus_log_of:
            return total_complexity / self.complexity_factor
        return np.exp(-1 * total_complexity / self.complexity_factor)

    def optimise_params(self, model_to_optimise) -> float:
        """Deprecated in the reference (learning_chain.py:857-883); kept for API completeness."""
        if not hasattr(self, 'costs_full') and not hasattr(self, 'costs_brief'):
            self.costs_full, self.costs_brief = [], []
        if not self.params_handler:
            self.initialise_params_handler()
        self.current, best_cost, costs = self.params_handler.do_optimisation(model_to_optimise)
        self.costs_brief.append(best_cost)
        self.costs_full = self.costs_full + costs
        return best_cost

    def cool_down(self):
        if not self.params_handler:
            self.initialise_params_handler()
        self.params_handler.rescale_jump_lengths(1 / self.jump_length_rescaling_factor)

    def heat_up(self):
        if not self.params_handler:
            self.initialise_params_handler()
        self.params_handler.rescale_jump_lengths(self.jump_length_rescaling_factor)

    def get_init_hyperparams(self):
        output = copy.deepcopy(self.init_hyperparams)
        output['initial guess'] = self.initial.model_description_dict()
        return output

    def initialise_params_handler(self):
        self.params_handler = params_handling.ParamsHandler(self)
        self.params_handler.set_hyperparams(self.params_handler_hyperparams)
        self.params_handler.set_bounds(self.params_bounds)

    def initialise_process_handler(self):
        self.process_handler = process_handling.ProcessHandler(
            self, bounds=self.params_bounds,
            qubit2defect_couplings_library=self.qubit2defect_couplings_library,
            defect2defect_couplings_library=self.defect2defect_couplings_library,
            qubit_Ls_library=self.qubit_Ls_library, defect_Ls_library=self.defect_Ls_library)

    def sample_T(self):
        """The MH temperature: the number itself, or a draw from gamma(shape, scale)."""
        tp = self.temperature_proposal
        if isinstance(tp, (int, float)) and not isinstance(tp, bool):
            return tp
        if (isinstance(tp, (tuple, list)) and len(tp) == 2
                and all(isinstance(x, (int, float)) for x in tp)):
            return np.random.gamma(*tp)
        raise RuntimeError('Metropolis-Hastings temperature proposal failed')

    def trunc_gaus(self, x, m, s, a, b):
        """Density at x of N(m, s) truncated to [a, b]."""
        return 1 / s * (1 / np.sqrt(2 * np.pi) * np.exp(-1 / 2 * ((x - m) / s) ** 2)) / (_Phi((b - m) / s) - _Phi((a - m) / s))

    def xi(self, m, s, a=-np.inf, b=np.inf):
        """Normaliser of N(m, s) truncated to [a, b]: Phi((b-m)/s) - Phi((a-m)/s)."""
        return _Phi((b - m) / s) - _Phi((a - m) / s)
