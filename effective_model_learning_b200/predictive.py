"""
Batched re-evaluation callers of the evaluator kernels (scope row N2).

  posterior_predictive  the loop of reference correlations.py:283-310 -- thousands of sampled parameter vectors
                        (vectorise_under_library layout) -> dynamics of each -> mean and std per observable and
                        time -- as ONE batched launch instead of one mesolve per vector;
  dense_grid_dynamics   the post-run evaluation of the best model on a dense grid
                        (execute_learning_full.py:186-188: linspace(ts[0], ts[-1], max(10 len(ts), 1000))).
"""
from __future__ import annotations

import numpy as np

from . import synthetic


def posterior_predictive(vectors, hyperparameters: dict, evaluation_times, observable_ops, *, D: int,
                         default_qubit_energy=1, qubit_initial_state=None, defect_initial_state=None, device=None):
    """
    vectors: [S][n_library] parameter vectors in the order of BasicModel.vectorise_under_library(hyperparameters)
    for one qubit + D defects (what the reference's clustering pickles hold).  Returns
    (means, stds, evaluated) as dicts keyed by observable name: arrays (T,), (T,), (S, T).
    """
    obs = [observable_ops] if isinstance(observable_ops, str) else list(observable_ops)
    layout = synthetic.LibraryLayout(1, D, hyperparameters)
    vectors = np.asarray(vectors, dtype=np.float64)
    if vectors.ndim != 2 or vectors.shape[1] != layout.n_library:
        raise RuntimeError(f'vectors must be [samples][{layout.n_library}] for this library and D')
    params = np.concatenate([vectors, np.full((len(vectors), layout.Q), float(default_qubit_energy))], axis=1)
    ctx = layout.context(np.asarray(evaluation_times, dtype=np.float64), obs, qubit_state=qubit_initial_state,
                         defect_state=defect_initial_state, device=device)
    expect, _ = ctx.evaluate(params, want_expect=True, want_loss=False)
    evaluated = {name: np.ascontiguousarray(expect[:, m].real) for m, name in enumerate(obs)}
    means = {name: evaluated[name].mean(axis=0) for name in obs}
    stds = {name: evaluated[name].std(axis=0) for name in obs}
    return means, stds, evaluated


def dense_grid(ts, factor: int = 10, minimum: int = 1000) -> np.ndarray:
    ts = np.asarray(ts, dtype=np.float64)
    return np.linspace(ts[0], ts[-1], max(factor * len(ts), int(minimum)))


def dense_grid_dynamics(model, ts, observable_ops, factor: int = 10, minimum: int = 1000):
    """(evaluation_ts, datasets) of `model` on the reference's dense post-run grid."""
    evaluation_ts = dense_grid(ts, factor, minimum)
    return evaluation_ts, model.calculate_dynamics(evaluation_ts, observable_ops=observable_ops,
                                                   custom_function_on_return=False)
