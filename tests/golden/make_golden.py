"""
Generates the committed fixtures in tests/golden/.  Run ONCE in the authoring container
(needs /root/reference, which does not exist on the GPU box; the tests only read the fixtures):

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

The REAL reference modules (definitions, basic_model, learning_model, learning_chain,
process_handling, params_handling, configs) are imported unmodified from /root/reference with
`oracle/qutip_shim` standing in for the uninstallable qutip 4.7.5; the only substituted
arithmetic is mesolve itself (exact expm stepping, see the shim).  Fixtures:

  reference_configs.json       configs.default_chain_hyperparams + the 18 library subsets
  reference_operators.npz      H, c_ops, rho0, e_ops the reference builds for seeded models
  reference_vectorise.json     vectorise_under_library output of the reference for those models
  oracle_observables.npz       oracle (exact) observables of seeded random library models, d = 4..32
  reference_chain_D{1,2}.json  step types / losses / acceptance probabilities / accept flags of the
                               reference's own LearningChain.run() under np.random.seed
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'qutip_shim'))
sys.path.insert(0, '/root/reference')

import qutip  # noqa: E402  (the shim)
import configs as ref_configs  # noqa: E402
import definitions as ref_definitions  # noqa: E402
import learning_model as ref_learning_model  # noqa: E402
import learning_chain as ref_learning_chain  # noqa: E402
from oracle import lindblad_oracle as lo  # noqa: E402

OBS3 = ['sigmax', 'sigmay', 'sigmaz']


def jsonable(x):
    if isinstance(x, dict):
        return {(k if isinstance(k, str) else repr(k)): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    return x


def spec_to_json(spec):
    out = []
    for t in spec['tls']:
        init = t['initial_state']
        out.append({'is_qubit': t['is_qubit'], 'energy': t['energy'], 'Ls': t['Ls'],
                    'initial_state': None if init is None else [[[z.real, z.imag] for z in row] for row in init],
                    'couplings': [[j, s, [list(p) for p in pairs]] for (j, s, pairs) in t['couplings']]})
    return {'tls': out}


def reference_model_from_spec(spec):
    """Build the REAL reference LearningModel from a plain spec."""
    m = ref_learning_model.LearningModel()
    made = []
    for t in spec['tls']:
        init = False if t['initial_state'] is None else qutip.Qobj(t['initial_state'])
        made.append(m.add_TLS(is_qubit=t['is_qubit'], energy=t['energy'], couplings={}, Ls=dict(t['Ls']),
                              initial_state=init))
    for i, t in enumerate(spec['tls']):
        for (j, strength, op_pairs) in t['couplings']:
            made[i].couplings.setdefault(made[j], []).append((strength, list(op_pairs)))
    m.build_operators()
    return m


def general_spec(rng):
    """A model using the non-Pauli part of the operator table as well."""
    names = list(lo.OPS)
    jump_ops = ['sigmax', 'sigmay', 'sigmaz', 'sigmap', 'sigmam', 'exc', 'gnd']
    n = int(rng.integers(1, 5))
    tls = []
    for k in range(n):
        init = [None, lo.OPS['plus'], lo.OPS['mm'], lo.OPS['2e+1g'], lo.OPS['gnd']][int(rng.integers(0, 5))]
        Ls = {op: float(rng.uniform(0.01, 0.4)) for op in rng.permutation(jump_ops)[:int(rng.integers(0, 3))]}
        tls.append({'is_qubit': bool(k == 1 or (n == 1)) if n > 1 else True, 'energy': float(rng.uniform(0.05, 1.2)),
                    'Ls': Ls, 'initial_state': None if init is None else init.copy(), 'couplings': []})
    for i in range(n):
        for j in range(n):
            if i != j and rng.uniform() < 0.4:
                pairs = [(str(rng.choice(names[:5])), str(rng.choice(names[:5]))) for _ in range(int(rng.integers(1, 3)))]
                tls[i]['couplings'].append((j, float(rng.uniform(0.1, 2.0)), pairs))
    return {'tls': tls}


def main():
    # ---------------------------------------------------------------- configs
    with open(os.path.join(HERE, 'reference_configs.json'), 'w') as f:
        json.dump({'default_chain_hyperparams': jsonable(ref_configs.default_chain_hyperparams),
                   'specific_experiment_chain_hyperparams': jsonable(ref_configs.specific_experiment_chain_hyperparams),
                   'order': list(ref_configs.specific_experiment_chain_hyperparams)}, f, indent=1)

    # ---------------------------------------------------------------- operators built by the reference
    arrays, specs, vec = {}, [], []
    hyper11 = ref_configs.get_hyperparams(list(ref_configs.specific_experiment_chain_hyperparams)[11])
    for i in range(24):
        rng = np.random.default_rng(500 + i)
        if i < 12:
            spec = lo.random_library_spec(rng, 1, i % 4)
        elif i < 16:
            spec = lo.random_library_spec(rng, 2, i % 3 + 1)
        else:
            spec = general_spec(rng)
        m = reference_model_from_spec(spec)
        specs.append(spec_to_json(spec))
        arrays[f'H_{i}'] = m.H.full()
        arrays[f'Ls_{i}'] = np.stack([c.full() for c in m.Ls]) if m.Ls else np.zeros((0,) + m.H.shape, dtype=complex)
        arrays[f'rho0_{i}'] = m.initial_DM.full()
        # e_ops exactly as calculate_dynamics builds them (basic_model.py:365-375)
        obs = []
        for op in OBS3:
            temp, got = 1, False
            for TLS in m.TLSs:
                if TLS.is_qubit and not got:
                    temp, got = ref_definitions.T(temp, ref_definitions.ops[op]), True
                else:
                    temp = ref_definitions.T(temp, ref_definitions.ops['id2'])
            obs.append(temp.full())
        arrays[f'obs_{i}'] = np.stack(obs)
        ts = np.linspace(0.0, 2.0, 41)
        arrays[f'expect_{i}'] = np.array([np.asarray(e, dtype=complex) for e in m.calculate_dynamics(ts, OBS3)])
        if i < 12:
            vals, labels, _ = m.vectorise_under_library(hyper11)
            vec.append({'values': [float(v) for v in vals], 'labels': labels})
        else:
            vec.append(None)
    np.savez_compressed(os.path.join(HERE, 'reference_operators.npz'), **arrays)
    with open(os.path.join(HERE, 'reference_operators_specs.json'), 'w') as f:
        json.dump({'specs': specs, 'times': list(np.linspace(0.0, 2.0, 41))}, f)
    with open(os.path.join(HERE, 'reference_vectorise.json'), 'w') as f:
        json.dump(vec, f)

    # ---------------------------------------------------------------- oracle observables (frozen)
    ts = np.linspace(0.0, 2.5, 200)
    frozen = {'times': ts}
    for (Q, D, count) in ((1, 1, 50), (1, 2, 50), (1, 3, 50), (2, 3, 4)):
        seeds = np.arange(1000, 1000 + count)
        out = np.zeros((count, 3, len(ts)))
        for n, seed in enumerate(seeds):
            spec = lo.random_library_spec(np.random.default_rng(int(seed)), Q, D)
            out[n] = np.array(lo.expect_exact(spec, ts, OBS3))
        frozen[f'expect_Q{Q}_D{D}'] = out
        frozen[f'seeds_Q{Q}_D{D}'] = seeds
    np.savez_compressed(os.path.join(HERE, 'oracle_observables.npz'), **frozen)

    # ---------------------------------------------------------------- reference chain trajectories
    for D, steps, seed in CHAIN_FIXTURES:
        chain_fixture(D, steps, seed)


CHAIN_FIXTURES = ((1, 400, 2024), (2, 250, 7), (3, 160, 31))      # D = 3: the headline size (d = 16)


def chain_fixture(D, steps, seed):
    """One seeded LearningChain.run() of the REAL reference (learning_chain.py:314-614) -> reference_chain_D{D}.json."""
    truth = lo.random_library_spec(np.random.default_rng(20260000 + 10 + D), 1, D)
    ts_c = np.linspace(0.0, 2.5, 100)
    noise = np.random.default_rng(99 + D)
    targets = [np.asarray(e) + noise.normal(0, 0.01, len(ts_c)) for e in lo.expect_exact(truth, ts_c, OBS3)]
    config = ref_configs.get_hyperparams(list(ref_configs.specific_experiment_chain_hyperparams)[11])
    config['iterations_till_progress_update'] = False
    np.random.seed(seed)
    chain = ref_learning_chain.LearningChain(
        target_times=ts_c, target_datasets=targets, target_observables=OBS3,
        initial=(1, D), qubit_initial_state=ref_definitions.ops['plus'],
        defect_initial_state=ref_definitions.ops['mm'], max_chain_steps=steps,
        store_all_proposals=True, **config)
    # Which partner "holds" a new coupling is decided in the reference by iterating a Python set
    # of objects (process_handling.py:247-257), i.e. by memory addresses.  Record
    # the outcome of every coupling addition so the port can be replayed with the same choices.
    holder_is_first = []
    ph = chain.process_handler
    def spying(orig):
        def wrapped(model, **kw):
            if not kw.get('update', True):
                return orig(model, **kw)
            before = [{id(p): len(v) for p, v in t.couplings.items()} for t in model.TLSs]
            out = orig(model, **kw)
            for a, t in enumerate(model.TLSs):
                for p, v in t.couplings.items():
                    if len(v) > before[a].get(id(p), 0):
                        holder_is_first.append(bool(a < model.TLSs.index(p)))
            return out
        return wrapped
    ph.add_random_qubit2defect_coupling = spying(ph.add_random_qubit2defect_coupling)
    ph.add_random_defect2defect_coupling = spying(ph.add_random_defect2defect_coupling)
    best = chain.run()
    vals, labels, _ = best.vectorise_under_library(config)
    with open(os.path.join(HERE, f'reference_chain_D{D}.json'), 'w') as f:
        json.dump({'D': D, 'steps': steps, 'seed': seed, 'config_index': 11,
                   'target_times': list(ts_c), 'targets': [list(map(float, t)) for t in targets],
                   'step_types': chain.run_step_type_tracker,
                   'loss': [float(x) for x in chain.explored_loss],
                   'acceptance_probability': [float(x) for x in chain.explored_acceptance_probability],
                   'acceptance': [bool(x) for x in chain.run_acceptance_tracker],
                   'best_loss': float(chain.best_loss), 'best_values': [float(v) for v in vals],
                   'best_labels': labels,
                   'counters': [chain.tot_RJ_steps, chain.acc_RJ_steps, chain.tot_tweak_steps, chain.acc_tweak_steps],
                   'coupling_holder_is_first': holder_is_first,
                   'mesolve_calls': qutip.MESOLVE_CALLS}, f)
    print(f'chain D={D}: {steps} steps, accepted {sum(chain.run_acceptance_tracker)}, best loss {chain.best_loss:.3e}, '
          f'mesolve calls so far {qutip.MESOLVE_CALLS}')


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == 'chains':      # only the listed chain fixtures: make_golden.py chains 3
        for D, steps, seed in CHAIN_FIXTURES:
            if str(D) in sys.argv[2:]:
                chain_fixture(D, steps, seed)
        sys.exit(0)
    main()
