"""
Slow per-chain Python restatement of the batched chain engine (csrc/eml_chains.cu), TEST ONLY.
It uses the same counter-based random stream and, for the model-dependent pieces (number of ways of each move,
model complexity), the host classes that were validated against the real reference (ProcessHandler on
LearningModel objects), so it checks the device engine against LearningChain's semantics.
"""
import math

import numpy as np

from effective_model_learning_b200 import process_handling

MASK = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15
MOVES = ('tweak all parameters', 'add qubit L', 'remove qubit L', 'add defect L', 'remove defect L',
         'add qubit-defect coupling', 'remove qubit-defect coupling',
         'add defect-defect coupling', 'remove defect-defect coupling')
_METHOD = {'add qubit L': 'add_random_qubit_L', 'add defect L': 'add_random_defect_L',
           'remove qubit L': 'remove_random_qubit_L', 'remove defect L': 'remove_random_defect_L',
           'add qubit-defect coupling': 'add_random_qubit2defect_coupling',
           'remove qubit-defect coupling': 'remove_random_qubit2defect_coupling',
           'add defect-defect coupling': 'add_random_defect2defect_coupling',
           'remove defect-defect coupling': 'remove_random_defect2defect_coupling'}
PCLASS = ('couplings', 'energies', 'Ls')


def splitmix64(x):
    x = (x + GOLD) & MASK
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


class Rng:
    def __init__(self, seed, chain):
        self.key = splitmix64((seed ^ splitmix64(chain)) & MASK)
        self.ctr = 0

    def next(self):
        v = splitmix64(self.key ^ ((self.ctr * GOLD) & MASK))
        self.ctr += 1
        return v

    def uniform(self):
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)

    def normal(self):
        u1 = 1.0 - self.uniform()
        u2 = self.uniform()
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


def rng_gamma(rng, shape, scale):
    """Marsaglia-Tsang, the same draw order as Rng::gamma in csrc/eml_chains.cu."""
    boost = 1.0
    if shape < 1.0:
        boost = rng.uniform() ** (1.0 / shape)
        shape += 1.0
    d = shape - 1.0 / 3.0
    c = 1.0 / math.sqrt(9.0 * d)
    for _ in range(64):
        x = rng.normal()
        v = 1.0 + c * x
        if v <= 0.0:
            continue
        v = v * v * v
        u = rng.uniform()
        if math.log(u) < 0.5 * x * x + d - d * v + d * math.log(v):
            return boost * d * v * scale
    return boost * d * scale


def Phi(y):
    return 0.5 * (1.0 + math.erf(y / math.sqrt(2.0)))


def xi(m, s, a, b):
    return Phi((b - m) / s) - Phi((a - m) / s)


class ReferenceChain:
    def __init__(self, layout, hyper, config, row, loss, seed, chain):
        self.layout, self.hyper, self.cfg = layout, hyper, config
        self.cur = np.array(row, dtype=float)
        self.cur_loss = self.best_loss = float(loss)
        self.best = self.cur.copy()
        self.jump = {k: config['params_handler_hyperparams']['initial_jump_lengths'][k] for k in PCLASS}
        self.rng = Rng(seed, chain)
        self.tracker = []
        self.step_i = 0
        self.k = 0
        self.counters = [0, 0, 0, 0]      # tot_RJ, acc_RJ, tot_tweak, acc_tweak
        total = sum(config['chain_step_options'].get(m, 0.0) for m in MOVES)
        self.prio = [config['chain_step_options'].get(m, 0.0) / total for m in MOVES]
        self.ph = process_handling.ProcessHandler(False, bounds=config['params_bounds'],
                                                  qubit2defect_couplings_library=hyper['qubit2defect_couplings_library'],
                                                  defect2defect_couplings_library=hyper['defect2defect_couplings_library'],
                                                  qubit_Ls_library=hyper['qubit_Ls_library'],
                                                  defect_Ls_library=hyper['defect_Ls_library'])

    def ways(self, row):
        m = self.layout.model_from_params(row)
        w = [1] + [getattr(self.ph, _METHOD[name])(m, update=False)[1] for name in MOVES[1:]]
        n_proc = (self.ph.count_Ls(m) + self.ph.count_defect2defect_couplings(m) + self.ph.count_qubit2defect_couplings(m))
        return w, n_proc

    def tweakable(self, c, v):
        cls = self.layout.classes[c]
        return cls == 'energies' or (cls != 'qubit_energy' and v != 0.0)

    def pclass(self, c):
        cls = self.layout.classes[c]
        return 'energies' if cls == 'energies' else cls

    def xi_product(self, row):
        out = 1.0
        b = self.cfg['params_bounds']
        for c, v in enumerate(row):
            if self.tweakable(c, v):
                pc = self.pclass(c)
                out *= xi(v, self.jump[pc], b[pc][0], b[pc][1])
        return out

    def propose(self):
        cfg, rng = self.cfg, self.rng
        if cfg['shock_anneal_at'] and self.step_i == cfg['shock_anneal_at']:
            self.jump = dict(cfg['params_handler_hyperparams']['annealed_jump_lengths'])
            self.step_i += 1
            self.cur, self.cur_loss = self.best.copy(), self.best_loss
            self.tracker.append(True)
        tp = cfg['temperature_proposal']
        self.T = rng_gamma(rng, tp[0], tp[1]) if isinstance(tp, (tuple, list)) else tp
        w = cfg['acceptance_window']
        if self.k >= w:
            self.k = 0
            ratio = sum(self.tracker[len(self.tracker) - w:]) / w
            if ratio - cfg['acceptance_target'] > cfg['acceptance_band']:
                self.jump = {k: v * (1 / cfg['jump_length_rescaling_factor']) for k, v in self.jump.items()}
            elif ratio - cfg['acceptance_target'] < -cfg['acceptance_band']:
                self.jump = {k: v * cfg['jump_length_rescaling_factor'] for k, v in self.jump.items()}
        self.k += 1
        ways, n_cur = self.ways(self.cur)
        wsum = sum(p * x for p, x in zip(self.prio, ways))
        u = rng.uniform()
        cdf, mv = 0.0, len(MOVES) - 1
        for m in range(len(MOVES)):
            cdf += self.prio[m] * ways[m] / wsum
            if cdf > u:
                mv = m
                break
        prop = self.cur.copy()
        b = cfg['params_bounds']
        if mv == 0:
            for c, v in enumerate(self.cur):
                if not self.tweakable(c, v):
                    continue
                pc = self.pclass(c)
                for _ in range(10):
                    cand = v + self.jump[pc] * rng.normal()
                    if b[pc][0] <= cand <= b[pc][1]:
                        prop[c] = cand
                        break
            p_there, p_back = self.xi_product(prop), self.xi_product(self.cur)
            self.counters[2] += 1
        else:
            name = MOVES[mv]
            add = name.startswith('add')
            kind = {'qubit L': ('Ls', True), 'defect L': ('Ls', False), 'qubit-defect coupling': ('couplings', True),
                    'defect-defect coupling': ('couplings', False)}[name.split(' ', 1)[1]]
            cols = [c for c, (key, cls) in enumerate(zip(self.layout.keys, self.layout.classes))
                    if cls == kind[0] and ((key[1] < self.layout.Q) == kind[1]) and ((self.cur[c] == 0.0) == add)]
            n = ways[mv]
            assert n == len(cols)
            pick = min(int(rng.uniform() * n), n - 1)
            c = cols[pick]
            prop[c] = b[kind[0]][0] + (b[kind[0]][1] - b[kind[0]][0]) * rng.uniform() if add else 0.0
            ways_p, _ = self.ways(prop)
            wsum_p = sum(p * x for p, x in zip(self.prio, ways_p))
            comp = mv + 1 if add else mv - 1
            p_there, p_back = self.prio[mv] / wsum, self.prio[comp] / wsum_p
            self.counters[0] += 1
        _, n_prop = self.ways(prop)
        cf = cfg['complexity_factor']
        self.ratio = math.exp(-n_prop / cf) / math.exp(-n_cur / cf) * p_back / p_there
        self.mv, self.prop = mv, prop
        return prop

    def decide(self, prop_loss):
        with np.errstate(over='ignore'):
            a = float(np.exp(-1.0 / self.T * (prop_loss - self.cur_loss))) * self.ratio
        accept = self.rng.uniform() < a
        if accept:
            self.cur, self.cur_loss = self.prop.copy(), prop_loss
            if prop_loss < self.best_loss:
                self.best, self.best_loss = self.prop.copy(), prop_loss
            self.counters[3 if self.mv == 0 else 1] += 1
        self.tracker.append(accept)
        self.step_i += 1
        return a, accept
