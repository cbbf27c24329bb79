"""
The straightforward form of `LearningChain.run` (test infrastructure): every step recomputes the move weights of the
current and of the proposed model with step(model, move, update=False) one move at a time, the truncated-Gaussian
normalisers one parameter at a time with the scalar xi(), and draws the move with np.random.choice(labels, p=...) -- the
structure of the reference's loop (learning_chain.py:346-596).  The product's `run` shortcuts all of these
(ProcessHandler.count_moves, a vectorised erf, numpy's choice restated on one random_sample(), quantities of the current
model carried between steps); tests/test_host.py runs both on the same seeds and compares every tracker bit for bit.
"""
import copy
import time

import numpy as np

from effective_model_learning_b200.learning_chain import LearningChain, _Phi


class PlainLoopChain(LearningChain):

    def _move_weights(self, model):
        return [self.next_step_priorities_dict[x] * self.step(model, x, update=False)[1]
                for x in self.next_step_labels]

    def _xi_product(self, model):
        jl, bounds = self.params_handler.jump_lengths, self.params_bounds
        out = 1
        for tls in model.TLSs:
            if not tls.is_qubit:
                out *= self.xi(tls.energy, jl['energies'], *bounds['energies'])
            for rate in tls.Ls.values():
                out *= self.xi(rate, jl['Ls'], *bounds['Ls'])
            for partner in tls.couplings:
                for coupling in tls.couplings[partner]:
                    out *= self.xi(coupling[0], jl['couplings'], *bounds['couplings'])
        return out

    def prior(self, model, return_minus_log_of: bool = False) -> float:
        ph = self.process_handler
        total_complexity = (ph.count_Ls(model) + ph.count_defect2defect_couplings(model)
                            + ph.count_qubit2defect_couplings(model))
        if return_minus_log_of:
            return total_complexity / self.complexity_factor
        return np.exp(-1 * total_complexity / self.complexity_factor)

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
        while i <= steps:
            if bool(self.shock_anneal_at) and i == self.shock_anneal_at:
                # switch to the annealed jump lengths and restart from the best model so far
                self.params_handler.set_jump_lengths(self.params_handler_hyperparams['annealed_jump_lengths'])
                now_annealed = True
                i += 1
                self.current = copy.deepcopy(self.best)
                self.current_loss = self.best_loss
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
                elif acceptance_ratio - self.acceptance_target < -self.acceptance_band:
                    self.heat_up()
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

            weights = self._move_weights(proposal)
            norm = sum(weights)
            next_step = str(np.random.choice(self.next_step_labels, p=[w / norm for w in weights]))
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
                    p_back = self._xi_product(self.current)
            else:
                p_there = self.next_step_priorities_dict[next_step] / sum(self._move_weights(self.current))
                p_back = (self.next_step_priorities_dict[self.complementary_step(next_step)]
                          / sum(self._move_weights(proposal)))

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
                else:
                    self.acc_RJ_steps += 1
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
