"""
CPU model (numpy) of the ALGORITHM the warp_cheb / cta_cheb CUDA kernels run -- used to choose and to
regression-test its constants (term-count formula, enclosure margins, step caps) without a GPU.
TEST INFRASTRUCTURE ONLY: nothing in the product package imports it.

Algorithm ("Chebyshev action with Bessel-weighted dense output", rho kept as a d x d matrix):
    L = L_s - ac (ac centres the real range of the dissipator), At = L_s / R with [-iR, iR] enclosing the
    imaginary extent of the numerical range of L_s (R >= spread of H, bounded by Gershgorin and by
    2 ||(H - mean diag)^8||_1^(1/8)):
        exp(tau L) rho = e^{-ac tau} sum_k J_k(tau R) T_k,   T_0 = rho, T_1 = 2 At rho, T_{k+1} = 2 At T_k + T_{k-1}
    (Jacobi-Anger; T_k = 2 i^k T_k^cheb(At / i) rho).  The T_k do not depend on tau: ONE recurrence serves every
    output time of a macro step; only Tr(O T_k) is kept and each output is a Bessel-weighted sum of them, the
    weights generated per output by Miller's backward recurrence.  The series is truncated for the confocal ellipse
    (semi-axes A, B) through the corner of the rectangle enclosing the numerical range: cheb_terms(tau (A + B)).
Mirrors effective_model_learning_b200/csrc/eml_cheb.cuh and the CHEB branches of eml_warp_mma.cu / eml_cta_mma.cu.
"""
import numpy as np

from . import taylor_action_model as tm

C1, C0 = 9.25, 5.0          # |J_k(z)| < 1e-14 for k >= z + C1 z^(1/3) + C0   (0 <= z <= 2000)
HUMP = 14.0                 # cap on tau * B
DELTAS = (0.03, 0.06, 0.12, 0.25, 0.5, 1.0)


def cheb_terms(z: float) -> int:
    return int(np.ceil(z + C1 * np.cbrt(z) + C0))


def zcap_for(cap: int) -> float:
    z = float(cap)
    while z > 0 and cheb_terms(z) > cap:
        z -= 0.25
    return z


def miller(K: int, z: float) -> np.ndarray:
    """J_0..J_K(z) by backward recurrence from (J_{K+1}, J_K) = (0, 1), normalised with J_0 + 2 sum J_2k = 1."""
    j = np.zeros(K + 2)
    j[K] = 1.0
    for k in range(K, 0, -1):
        j[k - 1] = (2 * k / z) * j[k] - j[k + 1]
    return j[:K + 1] / (j[0] + 2 * j[2:K + 1:2].sum())


def spread_bound(H: np.ndarray) -> float:
    """Upper bound on lambda_max(H) - lambda_min(H): min(Gershgorin, 2 ||(H - mu)^8||_1^(1/8))."""
    d = H.shape[0]
    diag = H.diagonal().real
    rad = np.abs(H).sum(axis=1) - np.abs(H.diagonal())
    gersh = (diag + rad).max() - (diag - rad).min()
    M = H - diag.mean() * np.eye(d)
    for _ in range(3):
        M = M @ M
    return min(gersh, 2.0 * np.abs(M).sum(axis=0).max() ** 0.125 * (1 + 1e-12))


def dissipator_ranges(c_ops, d: int):
    """Real range [dlo, dhi] of the numerical range of the dissipator and a bound on its imaginary extent
    (Gershgorin on the symmetric / antisymmetric parts of the column-stacked superoperator; the kernels obtain the
    same numbers site by site, the ranges of a Kronecker sum add)."""
    N = d * d
    Dm = np.zeros((N, N), complex)
    I = np.eye(d)
    for c in c_ops:
        cdc = c.conj().T @ c
        Dm += np.kron(c.conj(), c) - 0.5 * np.kron(I, cdc) - 0.5 * np.kron(cdc.T, I)
    dg = Dm.diagonal().real
    sym = 0.5 * np.abs(Dm + Dm.conj().T); sym -= np.diag(sym.diagonal())
    asym = 0.5 * np.abs(Dm - Dm.conj().T); asym -= np.diag(asym.diagonal())
    rr = sym.sum(axis=1)
    return (dg - rr).min(), (dg + rr).max(), asym.sum(axis=1).max()


def make_plan(spread, dlo, dhi, dasym, span, zcap, kcap):
    ac = -0.5 * (dlo + dhi)
    ah, yext = 0.5 * (dhi - dlo), spread + dasym
    best, have_ok = None, False
    for delta in DELTAS:
        R = max(yext * (1 + delta), ah, 1e-3 / max(span, 1e-300))
        cc, xr2 = R * R - ah * ah - yext * yext, ah * ah * R * R
        disc = np.sqrt(cc * cc + 4 * xr2)
        u = 2 * xr2 / (cc + disc) if cc > 0 else 0.5 * (disc - cc)
        B, A = np.sqrt(u), np.sqrt(u + R * R)
        ok = span * B <= HUMP
        if (ok and not have_ok) or (ok == have_ok and (best is None or A + B < best[1])):
            best, have_ok = (R, A + B, B), ok
    R, S, B = best
    tau_max = zcap / S
    if B > 0:
        tau_max = min(tau_max, HUMP / B)
    lnr = np.log(S / R)
    if lnr * kcap > 560.0:
        kall = 560.0 / lnr
        tau_max = min(tau_max, max(kall - C1 * np.cbrt(kall) - C0, 1.0) / S)
    return R, S, ac, tau_max


def evaluate(H, c_ops, rho0, obs, times, kcap=512):
    """Returns expect (K, T complex), number of generator applications, info dict."""
    d = H.shape[0]
    times = np.asarray(times, float)
    T = len(times)
    G, Gt = tm.superop_parts(H, c_ops)
    dlo, dhi, dasym = dissipator_ranges(c_ops, d)
    R, S, ac, tau_max = make_plan(spread_bound(H), dlo, dhi, dasym, times[-1] - times[0], zcap_for(kcap), kcap)
    Gs = (G + 0.5 * ac * np.eye(d)) * (2 / R)
    Gts = (Gt + 0.5 * ac * np.eye(d)) * (2 / R)
    cs = [c * np.sqrt(2 / R) for c in c_ops]
    out = np.zeros((len(obs), T), dtype=complex)
    rho = rho0.astype(complex)
    for m, O in enumerate(obs):
        out[m, 0] = np.trace(O @ rho)
    kidx = napp = nsteps = 0
    growth = 0.0
    while kidx < T - 1:
        t0 = times[kidx]
        q = 0
        while kidx + q + 1 <= T - 1 and times[kidx + q + 1] - t0 <= tau_max:
            q += 1
        nsub = 1
        if q == 0:
            nsub, q = int(np.ceil((times[kidx + 1] - t0) / tau_max)), 1
        tau_end = (times[kidx + q] - t0) / nsub
        final = kidx + q == T - 1
        if not tau_end * R >= 1e-40:
            for m, O in enumerate(obs):
                out[m, kidx + 1:kidx + q + 1] = np.trace(O @ rho)
            kidx += q
            continue
        K = min(cheb_terms(tau_end * S), kcap)
        for sub in range(nsub):
            nsteps += 1
            last = sub == nsub - 1
            need_acc = not (final and last)
            if need_acc:
                cJ = miller(K, tau_end * R) * np.exp(-ac * tau_end)
                acc = np.zeros_like(rho)
            tr = np.zeros((K + 1, len(obs)), complex)
            v, ph, halve = rho, np.zeros_like(rho), 1.0
            for k in range(K + 1):
                for m, O in enumerate(obs):
                    tr[k, m] = np.trace(O @ v)
                if need_acc:
                    acc = acc + cJ[k] * v
                if k == K:
                    break
                w = tm.apply_L(Gs, Gts, cs, v) + 2 * ph        # (2 At) v + previous term
                napp += 1
                ph, halve, v = halve * v, 0.5, w
            growth = max(growth, np.abs(v).max())
            if last:
                for p in range(q):
                    tau = times[kidx + 1 + p] - t0 if nsub == 1 else tau_end
                    if not tau * R >= 1e-40:
                        out[:, kidx + 1 + p] = tr[0]
                        continue
                    Kp = min(K, cheb_terms(tau * S))
                    out[:, kidx + 1 + p] = np.exp(-ac * tau) * (miller(Kp, tau * R)[:, None] * tr[:Kp + 1]).sum(axis=0)
            if need_acc:
                rho = acc
        kidx += q
    return out, napp, {'R': R, 'S': S, 'ac': ac, 'tau_max': tau_max, 'steps': nsteps, 'last_term_max': growth}
