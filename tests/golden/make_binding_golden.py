"""
Fixture of the reference-side binding (integration/eml_binding.py): run ONCE in the authoring container, where
/root/reference exists (the GPU box has neither the reference nor qutip; the tests only read the fixture):

    PYTHONHASHSEED=0 python tests/golden/make_binding_golden.py

The 24 seeded models of reference_operators_specs.json are built as the REAL reference's LearningModel / TwoLevelSystem
objects (imported unmodified from /root/reference, oracle/qutip_shim standing in for the uninstallable qutip) and
marshalled by eml_binding.marshal -- i.e. the exact arrays a reference-side caller hands to the C ABI are recorded
(reference_binding.npz).  tests/test_gpu_binding.py feeds them through eml_binding.evaluate on the GPU and compares with
what the reference's calculate_dynamics returned for the same models (reference_operators.npz: expect_i).
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
sys.path.insert(0, os.path.join(ROOT, 'integration'))

OBS3 = ['sigmax', 'sigmay', 'sigmaz']


def reference_descriptors():
    """[(descriptor of eml_binding.marshal)] for the 24 golden specs, built from the real reference classes."""
    import qutip  # noqa: F401  (the shim)
    import definitions as ref_definitions
    import learning_model as ref_learning_model
    import eml_binding
    with open(os.path.join(HERE, 'reference_operators_specs.json')) as f:
        golden = json.load(f)
    ts = np.array(golden['times'])
    out = []
    for spec in golden['specs']:
        m = ref_learning_model.LearningModel()
        made = []
        for t in spec['tls']:
            init = False
            if t['initial_state'] is not None:
                init = qutip.Qobj(np.array([[complex(*z) for z in row] for row in t['initial_state']]))
            made.append(m.add_TLS(is_qubit=t['is_qubit'], energy=t['energy'], couplings={}, Ls=dict(t['Ls']), initial_state=init))
        for i, t in enumerate(spec['tls']):
            for (j, strength, op_pairs) in t['couplings']:
                made[i].couplings.setdefault(made[j], []).append((strength, [tuple(p) for p in op_pairs]))
        m.build_operators()
        out.append(eml_binding.marshal(m, ts, OBS3, ref_definitions.ops))
    return out


def main():
    arrays = {}
    descs = reference_descriptors()
    for i, d in enumerate(descs):
        for key in ('terms', 'params', 'rho0', 'obs', 'times'):
            arrays[f'{key}_{i}'] = d[key]
        arrays[f'shape_{i}'] = np.array([d['n_tls'], d['obs_site']], dtype=np.int32)
    arrays['op_table'] = descs[0]['op_table']
    np.savez_compressed(os.path.join(HERE, 'reference_binding.npz'), **arrays)
    print(f'{len(descs)} descriptors, {sum(len(d["terms"]) for d in descs)} terms')


if __name__ == '__main__':
    main()
