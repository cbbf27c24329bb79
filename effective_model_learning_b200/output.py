"""
On-disk results of a chain run (scope row N3): the data files the reference's `output.Output` writes
(reference output.py:79-128, 200-201, 250-326) under the same names and with the same content, so that the
downstream analysis scripts (clustering, collation) find what they expect:

  <filename>_overall_acceptance.json                   acceptance per step class
  <filename>_<model name>.pickle / .txt                selected models (pickled LearningModel / description text)
  <filename>_proposals.pickle                          LearningChain.all_proposals
  <filename>[_<chain>]_hyperparameters.txt / .pickle   initial chain hyperparameters
  <filename>_loss.pickle, [_unannealed|_annealed]_accepted_loss.pickle, ..._accepted_log_posterior.pickle,
  ..._accepted_log_likelihood_prior.pickle, _AP.pickle, _accepted_AP.pickle

Models hold numpy arrays instead of qutip.Qobj, so the pickles load without qutip (they need this package).
With `reference_module_names=True` the model classes are pickled under the reference's flat module names
(`learning_model.LearningModel`, ...), which is what the reference's analysis scripts look up (find_clusters.py:71-90);
`compat.install_reference_aliases()` makes those names resolve to this package.
The reference's svg figures are drawn only when matplotlib is importable; it is not part of this image and the
figures are not part of the data contract.
"""
from __future__ import annotations

import contextlib
import json
import pickle
import pprint

from . import compat


class Output:
    """Instantiation immediately writes the outputs selected by `toggles` (a class or object with boolean attributes)."""

    def __init__(self, *, toggles, filename: str, dynamics_ts=False, dynamics_datasets=None,
                 dynamics_datasets_labels=None, dynamics_formatting=False, observable_labels=None, loss=None,
                 acceptance_probability=None, best_loss=None, acceptance=None, overall_acceptance=False,
                 models_to_save=None, model_names=None, all_proposals=None, chain_hyperparams=False,
                 chain_name=False, fontsize=False, reference_module_names: bool = False):
        loss = [] if loss is None else loss
        acceptance_probability = [] if acceptance_probability is None else acceptance_probability
        models_to_save = [] if models_to_save is None else models_to_save
        model_names = [] if model_names is None else model_names
        all_proposals = {} if all_proposals is None else all_proposals
        on = lambda name: bool(getattr(toggles, name, False))   # noqa: E731
        self.written = []

        naming = compat.reference_module_names if reference_module_names else contextlib.nullcontext

        def dump(path, obj, binary=True):
            if binary:
                with open(path, 'wb') as f, naming():
                    pickle.dump(obj, f)
            else:
                with open(path, 'w') as f:
                    f.write(obj)
            self.written.append(path)

        if overall_acceptance:
            with open(filename + '_overall_acceptance.json', 'w') as f:
                json.dump(overall_acceptance, f)
            self.written.append(filename + '_overall_acceptance.json')

        for i, model in enumerate(models_to_save):
            name = '_' + (str(i) if (not model_names or len(model_names) < len(models_to_save)) else model_names[i])
            if on('pickle'):
                dump(filename + name + '.pickle', model)
            if on('text'):
                dump(filename + name + '.txt', model.model_description_str(), binary=False)

        if on('all_proposals'):
            dump(filename + '_proposals.pickle', all_proposals)

        if on('hyperparams'):
            chain = '_' + chain_name if chain_name else ''
            dump(filename + chain + '_hyperparameters.txt', pprint.pformat(chain_hyperparams), binary=False)
            dump(filename + chain + '_hyperparameters.pickle', chain_hyperparams)

        if on('loss'):
            dump(filename + '_loss.pickle', loss)

        annealed = list(all_proposals.get('annealed', []))
        accepted = list(all_proposals.get('acceptance', []))
        both = any(annealed) and not all(annealed)
        for stage, enabled in (('unannealed', both), ('annealed', both), ('', True)):
            if not enabled or not all_proposals:
                continue
            tag = '_' + stage if stage else ''
            keep = [bool(a) and (stage == '' or z == (stage == 'annealed')) for a, z in zip(accepted, annealed)]

            def select(key):
                return [x for x, k in zip(all_proposals[key][1:], keep) if k]
            if on('loss'):
                dump(filename + tag + '_accepted_loss.pickle', select('loss'))
            if on('log_posterior'):
                dump(filename + tag + '_accepted_log_posterior.pickle', select('log_posterior'))
                dump(filename + tag + '_accepted_log_likelihood_prior.pickle', select('log_likelihood_prior'))

        if on('acceptance_probability'):
            dump(filename + '_AP.pickle', acceptance_probability)
            dump(filename + '_accepted_AP.pickle',
                 [x for x, a in zip(all_proposals.get('acceptance_probability', []), accepted) if a])

        self._maybe_plot(toggles, filename, dynamics_ts, dynamics_datasets, dynamics_datasets_labels,
                         dynamics_formatting, observable_labels, loss, best_loss)

    def _maybe_plot(self, toggles, filename, ts, datasets, labels, formatting, observables, loss, best_loss):
        """Comparison and loss figures (reference output.py:132-201) when matplotlib exists; silently skipped otherwise."""
        try:
            import matplotlib
            matplotlib.use('Agg')
            import matplotlib.pyplot as plt
        except Exception:
            return
        if getattr(toggles, 'comparison', False) and datasets and observables:
            fmt = formatting if isinstance(formatting, list) and formatting else ['b+', 'b.', 'r-', 'g-']
            for m, obs in enumerate(observables):
                plt.figure()
                for j, data in enumerate(datasets):
                    plt.plot(ts[j] if isinstance(ts, list) else ts, data[m], fmt[j % len(fmt)],
                             label=labels[j] if labels and j < len(labels) else None, linewidth=1, markersize=1)
                plt.xlabel('time')
                plt.ylabel(str(obs))
                plt.legend()
                plt.savefig(filename + '_' + str(obs) + '_comparison.svg', bbox_inches='tight')
                plt.close()
        if getattr(toggles, 'loss', False) and len(loss):
            plt.figure()
            plt.plot(loss, 'm-', linewidth=0.3)
            plt.yscale('log')
            plt.xlabel('iteration')
            plt.ylabel('loss')
            plt.savefig(filename + '_loss.svg', bbox_inches='tight')
            plt.close()
