"""
Multiset sweep on one node: the nested loop of the reference's bash launchers
(launch_execute_learning_multiset.sh:35-66 -- target datasets x configurations x defect numbers x repetitions,
one OS process per chain) as ONE job: every (dataset, configuration, D, repetition) is a chain.  Chains of one
(dataset, configuration, D) cost the same, chains of different D do not (d = 16 costs ~3.5x d = 4), so EVERY such group
is split over the ranks (`dist.balanced_owners`; one process per GPU): each rank owns 1/world of every model size and
runs one `BatchedChains` engine per group -- models of different sizes run in different launches of the same kernels,
concurrently on separate CUDA streams.  A chain's random stream is keyed by its GLOBAL index, so the results do not
depend on the partition.  There is no collective on the hot path; at checkpoints the per-chain statistics of all
ranks are gathered with one all_gather (`dist.gather_chain_stats`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import configs, dist
from .chains import BatchedChains


class ChainSpec:
    """One chain of the sweep (what one `python execute_learning*.py ...` process was in the reference)."""

    __slots__ = ('dataset', 'config', 'defects', 'repetition')

    def __init__(self, dataset: int, config: int, defects: int, repetition: int):
        self.dataset, self.config, self.defects, self.repetition = dataset, config, defects, repetition

    def key(self):
        return (self.dataset, self.config, self.defects)


def enumerate_chains(n_datasets: int, config_numbers, defects_numbers, repetitions_numbers) -> list[ChainSpec]:
    """Launcher order: dataset, then configuration, then defect number, then repetition (1-based)."""
    reps = list(repetitions_numbers)
    out = []
    for ds in range(n_datasets):
        for cfg in config_numbers:
            for i, D in enumerate(defects_numbers):
                n_rep = reps[i] if len(reps) == len(defects_numbers) else (reps[0] if len(reps) == 1 else 3)
                out.extend(ChainSpec(ds, cfg, D, r) for r in range(1, n_rep + 1))
    return out


class MultisetSweep:
    """
    datasets: list of (target_times, target_datasets [K][T], target_observables).
    Every rank constructs the same sweep; each only instantiates engines for its own chain range.
    """

    def __init__(self, datasets, config_numbers, defects_numbers, repetitions_numbers, *, seed: int = 0,
                 qubit_initial_state=None, defect_initial_state=None, device: int | None = None,
                 config_overrides: dict | None = None):
        self.rank, self.world = (torch.distributed.get_rank(), torch.distributed.get_world_size()) \
            if torch.distributed.is_initialized() else (0, 1)
        self.chains = enumerate_chains(len(datasets), config_numbers, defects_numbers, repetitions_numbers)
        self.n_chains = len(self.chains)
        if device is None:
            device = torch.cuda.current_device()
        names = list(configs.specific_experiment_chain_hyperparams)
        # groups of equal-cost chains (contiguous in the launcher order); every group is split over the ranks
        all_groups: dict[tuple, list[int]] = {}
        for gidx, spec in enumerate(self.chains):
            all_groups.setdefault(spec.key(), []).append(gidx)
        for members in all_groups.values():
            assert members == list(range(members[0], members[0] + len(members))), 'a group must be contiguous in the launcher order'
        self.owners = dist.balanced_owners([len(m) for m in all_groups.values()], self.world)
        self.owned = self.owners[self.rank]
        self._row_of = {int(g): i for i, g in enumerate(self.owned)}
        mine = set(int(g) for g in self.owned)
        groups = {key: [g for g in members if g in mine] for key, members in all_groups.items()}
        self.engines = []
        for key, members in groups.items():
            if not members:
                continue
            ds, cfg_no, D = key
            cfg = configs.get_hyperparams(names[cfg_no])
            cfg.update(config_overrides or {})
            ts, data, obs = datasets[ds]
            eng = BatchedChains(ts, data, obs, n_chains=len(members), n_defects=D,
                                qubit_initial_state=qubit_initial_state, defect_initial_state=defect_initial_state,
                                seed=seed, first_chain=members[0], device=device, **cfg)
            self.engines.append((key, members, eng, torch.cuda.Stream(device=device)))

    def run(self, steps: int):
        """`steps` proposals for every chain of this rank; engines advance concurrently on their own streams."""
        for _, _, eng, stream in self.engines:
            eng.run(steps, stream=stream.cuda_stream)

    def synchronize(self):
        for _, _, _, stream in self.engines:
            stream.synchronize()

    def local_stats(self) -> np.ndarray:
        """[n_owned][8] statistics of this rank's chains, rows in the order of `self.owned` (ascending global index)."""
        out = np.zeros((len(self.owned), len(dist.CHAIN_STATS_FIELDS)))
        for _, members, eng, stream in self.engines:
            stream.synchronize()
            out[[self._row_of[g] for g in members]] = eng.read()['stats']
        return out

    def checkpoint(self) -> np.ndarray:
        """Global [n_chains][8] table on every rank -- the one collective of the sweep."""
        return dist.gather_chain_stats(self.local_stats(), self.n_chains, owners=self.owners)

    def best_models(self, global_index: int):
        for _, members, eng, _ in self.engines:
            if global_index in members:
                return eng.best_models([members.index(global_index)])[0]
        raise RuntimeError('chain %d is not owned by rank %d' % (global_index, self.rank))

    def close(self):
        for _, _, eng, _ in self.engines:
            eng.close()
        self.engines = []
