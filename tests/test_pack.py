"""
Host test of the model-packing core (effective_model_learning_b200/csrc/eml_pack.cuh), which the one-CTA kernel
pack_models_kernel (csrc/eml_order.cu) runs on the device: tests/pack_harness.cpp compiles the same header with g++ and
replays the kernel's phases sequentially.  Checked: the list is a permutation for every batch shape and for degenerate
costs, the start keys are sorted, and on the application counts of the bench batch
(tests/golden/bench_application_counts.npy, written by tools/schedule_model.py --save) a greedy fetch of the packed
list keeps the warp slots busier than the longest-first list.
"""
import ctypes
import heapq
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('pack') / 'libpack.so')
    subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'effective_model_learning_b200', 'csrc'),
                    os.path.join(ROOT, 'tests', 'pack_harness.cpp'), '-o', out], check=True)
    return ctypes.CDLL(out)


def pack(lib, cost, slots, eps=0.05):
    n = len(cost)
    c = np.ascontiguousarray(cost, dtype=np.float32)
    order = np.full(n, -1, dtype=np.int32)
    keys = np.full(n, -1, dtype=np.int32)
    cap = lib.pack_order_host(n, c.ctypes.data_as(ctypes.c_void_p), slots, ctypes.c_float(eps),
                              order.ctypes.data_as(ctypes.c_void_p), keys.ctypes.data_as(ctypes.c_void_p))
    return order, keys, cap


def busy_fraction(cost, order, slots):
    """Greedy fetch: the next model of the list goes to the slot that is free first (all slots equally fast)."""
    h = [0.0] * slots
    for b in order:
        heapq.heapreplace(h, h[0] + cost[b])
    return cost.sum() / slots / max(h)


def bench_costs():
    napp = np.load(os.path.join(ROOT, 'tests', 'golden', 'bench_application_counts.npy')).astype(float)
    return 45.0 + 1.28 * napp      # the sector kernel's cost model (csrc/eml_warp_mma.cu launch_warp_mma)


def test_bench_application_counts_fixture():
    napp = np.load(os.path.join(ROOT, 'tests', 'golden', 'bench_application_counts.npy'))
    assert napp.shape == (4096,) and abs(napp.mean() - 246.73) < 0.01      # the device reports 246.730 for the same batch


@pytest.mark.parametrize('n,slots', [(4096, 1776), (4096, 1184), (1777, 1776), (2048, 1776), (6000, 1776), (14208, 1776),
                                     (16384, 2368), (300, 37), (5, 3), (1, 1)])
def test_packed_list_is_a_permutation_with_sorted_keys(lib, n, slots):
    rng = np.random.default_rng(n + slots)
    cost = bench_costs()[rng.integers(0, 4096, n)] * rng.uniform(0.9, 1.1, n)
    order, keys, cap = pack(lib, cost, slots)
    assert sorted(order.tolist()) == list(range(n))
    assert 64 <= cap < 1024
    assert np.all(np.diff(keys) >= 0) and keys.min() >= 0 and keys.max() <= cap + 1
    # the first models of the list are the first models of the bins: no more of them than bins
    assert (keys == 0).sum() <= slots


@pytest.mark.parametrize('cost', [np.zeros(100), np.full(100, np.nan), np.full(100, np.inf), np.arange(100.0),
                                  np.array([1e30, 1.0, 2.0]), np.full(5000, 7.0), -np.ones(10)])
def test_degenerate_costs_still_give_a_permutation(lib, cost):
    order, keys, cap = pack(lib, cost, 7)
    assert sorted(order.tolist()) == list(range(len(cost)))


def test_packing_beats_the_longest_first_list_on_the_bench_batch(lib):
    cost = bench_costs()
    slots = 148 * 12
    lpt = busy_fraction(cost, np.argsort(-cost, kind='stable'), slots)
    order, keys, cap = pack(lib, cost, slots)
    packed = busy_fraction(cost, order, slots)
    assert 0.80 < lpt < 0.84          # 2.3 models per slot: the quantisation DESIGN.md section 5 measures on the device
    assert packed > 0.92
    # robust against a wrong cost model: packed with other constants, executed with the real ones
    for fixed, per in ((20.0, 1.28), (80.0, 1.28), (45.0, 1.0), (45.0, 1.6)):
        napp = (cost - 45.0) / 1.28
        order, _, _ = pack(lib, fixed + per * napp, slots)
        assert busy_fraction(cost, order, slots) > 0.89
    # and against 3 % noise of the real run times
    noisy = cost * (1.0 + 0.03 * np.random.default_rng(0).standard_normal(len(cost)))
    order, _, _ = pack(lib, cost, slots)
    assert busy_fraction(noisy, order, slots) > busy_fraction(noisy, np.argsort(-cost, kind='stable'), slots) + 0.05


def test_register_and_memory_paths_of_the_packing_agree(lib):
    """pack_classes keeps the level bitmask in registers for capacities below 128 units: same tables as the general path."""
    rng = np.random.default_rng(11)
    took_registers = 0
    for trial in range(400):
        n = int(rng.integers(1, 6000))
        slots = int(rng.integers(1, 2500))
        shape = rng.choice(['bench', 'flat', 'two', 'one'])
        if shape == 'bench':
            q = np.clip((bench_costs()[rng.integers(0, 4096, n)] / 720.0 * 63).astype(int), 0, 63)
        elif shape == 'flat':
            q = rng.integers(0, 64, n)
        elif shape == 'two':
            q = rng.choice([3, 60], n)
        else:
            q = np.full(n, int(rng.integers(0, 64)))
        n_class = np.bincount(63 - q, minlength=64).astype(np.int32)      # class 0 = most expensive
        res = lib.pack_paths_agree(n_class.ctypes.data_as(ctypes.c_void_p), slots, ctypes.c_float(float(rng.uniform(0, 0.2))))
        assert res in (1, 2), (trial, n, slots, shape)
        took_registers += res == 2
    assert 50 < took_registers < 400
