"""
Chain hyperparameter configurations: `default_chain_hyperparams`, the 18 named process-library
subsets `specific_experiment_chain_hyperparams`, and `get_hyperparams(name)` -- the same data
as the reference's configs.py (defaults :24-109, subsets :113-689, getter :693-700), generated
from a compact table instead of spelled-out dictionaries.  tests/golden/reference_configs.json
(dumped from the reference by tests/golden/make_golden.py) pins every value.

Keys of the dictionaries are LearningChain initialiser argument names.
"""
from __future__ import annotations

import copy

couplings_shape_scale = (1.1, 10)
Ls_shape_scale = (1.02, 4)

_PAULI = {'sx': 'sigmax', 'sy': 'sigmay', 'sz': 'sigmaz'}
_ALL = ('sx', 'sy', 'sz')


def _Ls_library(shorts) -> dict:
    return {_PAULI[s]: Ls_shape_scale for s in shorts}


def _couplings_library(shorts) -> dict:
    # same operator on both partners; one operator pair per coupling type
    return {((_PAULI[s], _PAULI[s]),): couplings_shape_scale for s in shorts}


default_chain_hyperparams = {
    'chain_step_options': {
        'tweak all parameters': 36,
        'add qubit L': 1,
        'remove qubit L': 1,
        'add defect L': 1,
        'remove defect L': 1,
        'add qubit-defect coupling': 1,
        'remove qubit-defect coupling': 1,
        'add defect-defect coupling': 1,
        'remove defect-defect coupling': 1,
    },
    'temperature_proposal': 0.0002,          # value, or (shape, scale) of a gamma to sample
    'complexity_factor': 10,                 # expected number of processes
    'jump_length_rescaling_factor': 1.0,
    'shock_anneal_at': int(500000),          # iteration at which the annealed jump lengths take over
    'acceptance_window': 10,
    'acceptance_target': 0.4,
    'acceptance_band': 0.2,
    'params_handler_hyperparams': {
        'initial_jump_lengths': {'couplings': 0.4, 'energies': 0.04, 'Ls': 0.04},
        'annealed_jump_lengths': {'couplings': 0.04, 'energies': 0.004, 'Ls': 0.004},
    },
    'qubit_Ls_library': _Ls_library(_ALL),
    'defect_Ls_library': _Ls_library(_ALL),
    'qubit2defect_couplings_library': _couplings_library(_ALL),
    'defect2defect_couplings_library': _couplings_library(_ALL),
    'params_thresholds': {'Ls': 1e-7, 'couplings': 1e-6},
    'params_priors': {'couplings': (1.1, 10), 'energies': (1.1, 10), 'Ls': (1.02, 4)},
    'params_bounds': {'couplings': (0.1, 4), 'energies': (0.01, 0.4), 'Ls': (0.01, 0.4)},
    'custom_function_on_dynamics_return': False,
    'iterations_till_progress_update': 1000,
}

# (experiment name, qubit Ls, defect Ls, qubit-defect couplings, defect-defect couplings)
# NB the reference's names 5 and 11 read "Lvirt-sz,sy,sz" although the library is sx,sy,sz.
_SUBSETS = (
    ('Lsyst-sz-Lvirt--Cs2v-sx-Cv2v--', ('sz',), (), ('sx',), ()),
    ('Lsyst-sz-Lvirt--Cs2v-sx-Cv2v-sx-', ('sz',), (), ('sx',), ('sx',)),
    ('Lsyst-sx,sy,sz-Lvirt--Cs2v-sx-Cv2v--', _ALL, (), ('sx',), ()),
    ('Lsyst-sx,sy,sz-Lvirt--Cs2v-sx-Cv2v-sx-', _ALL, (), ('sx',), ('sx',)),
    ('Lsyst-sx,sy,sz-Lvirt-sx,sy,sz-Cs2v-sx-Cv2v--', _ALL, _ALL, ('sx',), ()),
    ('Lsyst-sx,sy,sz-Lvirt-sz,sy,sz-Cs2v-sx-Cv2v-sx-', _ALL, _ALL, ('sx',), ('sx',)),
    ('Lsyst-sz-Lvirt--Cs2v-sx,sy,sz-Cv2v--', ('sz',), (), _ALL, ()),
    ('Lsyst-sz-Lvirt--Cs2v-sx,sy,sz-Cv2v-sx,sy,sz-', ('sz',), (), _ALL, _ALL),
    ('Lsyst-sx,sy,sz-Lvirt--Cs2v-sx,sy,sz-Cv2v--', _ALL, (), _ALL, ()),
    ('Lsyst-sx,sy,sz-Lvirt--Cs2v-sx,sy,sz-Cv2v-sx,sy,sz-', _ALL, (), _ALL, _ALL),
    ('Lsyst-sx,sy,sz-Lvirt-sx,sy,sz-Cs2v-sx,sy,sz-Cv2v--', _ALL, _ALL, _ALL, ()),
    ('Lsyst-sx,sy,sz-Lvirt-sz,sy,sz-Cs2v-sx,sy,sz-Cv2v-sx,sy,sz-', _ALL, _ALL, _ALL, _ALL),   # 11: everything
    ('Lsyst-sx-Lvirt--Cs2v-sx-Cv2v--', ('sx',), (), ('sx',), ()),
    ('Lsyst-sx-Lvirt--Cs2v-sz-Cv2v--', ('sx',), (), ('sz',), ()),
    ('Lsyst-sz-Lvirt--Cs2v-sz-Cv2v--', ('sz',), (), ('sz',), ()),
    ('Lsyst-sz-Lvirt--Cs2v-sx,sy-Cv2v--', ('sz',), (), ('sx', 'sy'), ()),
    ('Lsyst-sz-Lvirt--Cs2v-sx,sz-Cv2v--', ('sz',), (), ('sx', 'sz'), ()),
    ('Lsyst-sz-Lvirt--Cs2v-sx-Cv2v-sx,sy,sz-', ('sz',), (), ('sx',), _ALL),
)

specific_experiment_chain_hyperparams = {
    name: {
        'qubit_Ls_library': _Ls_library(qL),
        'defect_Ls_library': _Ls_library(dL),
        'qubit2defect_couplings_library': _couplings_library(q2d),
        'defect2defect_couplings_library': _couplings_library(d2d),
    }
    for name, qL, dL, q2d, d2d in _SUBSETS
}


def get_hyperparams(batch_name: str) -> dict:
    """Defaults with the named experiment's libraries substituted."""
    hyperparameters = copy.deepcopy(default_chain_hyperparams)
    hyperparameters.update(specific_experiment_chain_hyperparams[batch_name])
    return hyperparameters
