"""CPU checks of bench.py: the reference arm runs here (no GPU) and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1', '--cpu-sample', '8'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'evals/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'd=16' in d['config']['workload'] and d['vs_baseline'] is None


def test_non_rank0_reference_arm_is_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_canonical_dense_flops_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    # s = 0, T = 200, K = 3
    for n, want in ((2, 0.726e6), (3, 22.2e6), (4, 1.090e9), (5, 64.67e9)):
        assert abs(bench.canonical_dense_flops(n, 200, 3) / want - 1) < 0.01


def test_synthetic_workload_equals_oracle_generator():
    from effective_model_learning_b200 import configs, synthetic
    from oracle import lindblad_oracle as lo
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[11])
    for Q, D in ((1, 1), (1, 3), (2, 3)):
        lay = synthetic.LibraryLayout(Q, D, hyper)
        for s in (1000, 1001, 5000):
            m = lay.model_from_params(lay.random_params(np.random.default_rng(s)))
            a = lo.build_H(lo.spec_from_model(m))
            b = lo.build_H(lo.random_library_spec(np.random.default_rng(s), Q, D))
            assert np.array_equal(a, b)
