"""
CPU model of how the persistent warps of the d = 16 evaluator share one batch (no GPU needed).

The kernel hands models to 148 SMs x 12 resident warps in longest-first order of a cost proxy
(sum |E| + sum s^2 + sum rate, eml_warp_mma.cu model_cost_kernel).  This script computes, for the models of the bench
batch (seeds 1000 + b), the generator applications the Chebyshev plan needs (oracle/chebyshev_action_model.py: same
constants as the kernel, mean 246.7 as measured on the device) and simulates greedy list scheduling under a cost
model  fixed + (1 + output share) * applications  with every warp running at the same speed.

    python tools/schedule_model.py [--chains 4096] [--warps 1776]

Result for the bench batch (4096 models, 2.3 per warp slot): index order 0.62, proxy order 0.81, exact-cost order 0.82 of
the warp-slot time busy -- the proxy is as good as the exact cost, the remaining loss is the QUANTISATION of 2.3 models
per slot (ncu on the device: 16 % of the warp-slot time idle).  A best-fit-decreasing PACKING, listed
by start time for the dynamic fetch, reaches 0.95 with exact costs and 0.84 with the (affine-fitted) proxy.
"""
import argparse
import heapq
import sys

import numpy as np

sys.path.insert(0, '.')
from oracle import chebyshev_action_model as cm  # noqa: E402
from oracle import lindblad_oracle as lo  # noqa: E402


def plan_applications(spec, times, kcap):
    H, c_ops = lo.build_H(spec), lo.build_c_ops(spec)
    dlo, dhi, dasym = cm.dissipator_ranges(c_ops, H.shape[0])
    R, S, ac, tau_max = cm.make_plan(cm.spread_bound(H), dlo, dhi, dasym, times[-1] - times[0], cm.zcap_for(kcap), kcap)
    T, kidx, n = len(times), 0, 0
    while kidx < T - 1:
        t0, q = times[kidx], 0
        while kidx + q + 1 <= T - 1 and times[kidx + q + 1] - t0 <= tau_max:
            q += 1
        nsub = 1
        if q == 0:
            nsub, q = int(np.ceil((times[kidx + 1] - t0) / tau_max)), 1
        n += nsub * min(cm.cheb_terms((times[kidx + q] - t0) / nsub * S), kcap)
        kidx += q
    return n


def proxy_cost(spec):
    c = 0.0
    for t in spec['tls']:
        c += abs(t['energy']) + sum(t['Ls'].values())
        for (_, s, pairs) in t['couplings']:
            c += s * s * len(pairs)
    return c


def best_fit_list(est, workers, eps=0.06, levels=256):
    """
    Best-fit-decreasing packing of the estimated costs, quantised to `levels` classes, into `workers` bins of capacity
    (1 + eps) * mean load; returns the models listed by their start time inside their bin (what a dynamic fetch needs to
    reproduce the packing), models that did not fit at the end.  One step moves min(#models of the class, #bins at the
    best-fitting residual level) models at once: ~300 steps for 4096 models.
    """
    q = np.maximum(1, np.floor(est * ((levels - 1) / est.max())).astype(int))
    cap = int(np.ceil(q.sum() / workers * (1.0 + eps)))
    bins_at = np.zeros(cap + 1, int)
    bins_at[cap] = workers
    start = np.full(len(est), 10 ** 9)
    for c in range(levels - 1, 0, -1):
        jobs = list(np.where(q == c)[0])
        while jobs:
            r = c
            while r <= cap and bins_at[r] == 0:
                r += 1
            if r > cap:
                break
            m = min(len(jobs), bins_at[r])
            for _ in range(m):
                start[jobs.pop()] = cap - r
            bins_at[r] -= m
            bins_at[r - c] += m
    return np.lexsort((-est, start))


def makespan(cost, order, workers):
    h = [0.0] * workers
    for b in order:
        heapq.heapreplace(h, h[0] + cost[b])
    return max(h)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--chains', type=int, default=4096)
    ap.add_argument('--warps', type=int, default=148 * 12)
    ap.add_argument('--fixed', type=float, default=47.0, help='set-up + epilogue in units of one application (ncu: ~12 %)')
    ap.add_argument('--output-share', type=float, default=0.27, help='output phase per application (ncu: ~17 %)')
    ap.add_argument('--save', default='', help='write the application counts (int16 .npy), e.g. tests/golden/bench_application_counts.npy')
    a = ap.parse_args()
    ts = np.linspace(0.0, 2.5, 200)
    specs = [lo.random_library_spec(np.random.default_rng(1000 + b), 1, 3) for b in range(a.chains)]
    napp = np.array([plan_applications(s, ts, 540) for s in specs], float)
    proxy = np.array([proxy_cost(s) for s in specs])
    if a.save:
        np.save(a.save, napp.astype(np.int16))
    cost = a.fixed + (1.0 + a.output_share) * napp
    ideal = cost.sum() / a.warps
    print(f'{a.chains} models on {a.warps} warp slots: applications mean {napp.mean():.1f} (min {napp.min():.0f}, max {napp.max():.0f}), '
          f'corr(proxy, applications) = {np.corrcoef(proxy, napp)[0, 1]:.3f}')
    for name, key in (('index order', None), ('proxy, longest first', proxy), ('exact cost, longest first', cost)):
        order = np.arange(a.chains) if key is None else np.argsort(-key, kind='stable')
        print(f'  {name:28s} busy fraction of the warp-slot time {ideal / makespan(cost, order, a.warps):.3f}')
    A = np.vstack([proxy, np.ones(a.chains)]).T
    fitted = A @ np.linalg.lstsq(A, cost, rcond=None)[0]
    print(f'  best-fit packing, exact cost   busy fraction {ideal / makespan(cost, best_fit_list(cost, a.warps), a.warps):.3f}')
    print(f'  best-fit packing, proxy (affine fit, rms error {np.sqrt(np.mean((fitted - cost) ** 2)) / cost.mean():.3f})   '
          f'busy fraction {ideal / makespan(cost, best_fit_list(fitted, a.warps), a.warps):.3f}')
