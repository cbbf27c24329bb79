"""
CPU model (numpy) of the ALGORITHM the CUDA kernels run -- used to choose and to
regression-test its constants (step-size target, Taylor cap, stopping tolerance) without a
GPU.  TEST INFRASTRUCTURE ONLY: nothing in the product package imports it.

Algorithm ("Taylor action with dense output", rho kept as a d x d matrix, never the
d^2 x d^2 superoperator):
    Lrho = G rho + rho Gt + sum_c c rho c^dag,   G = -iH - 1/2 sum c^dag c,  Gt = +iH - 1/2 sum c^dag c
    macro step [t_s, t_e] covering q >= 1 output times:  v_0 = rho(t_s),
    v_{j+1} = (h/(j+1)) L v_j  (h = t_e - t_s);  rho(t_s + x h) = sum_j x^j v_j.
    Only Tr(O v_j) is needed at interior output times, the full sum only at t_e.
"""
import numpy as np


def superop_parts(H, c_ops):
    d = H.shape[0]
    K = np.zeros((d, d), dtype=complex)
    for c in c_ops:
        K += c.conj().T @ c
    G = -1j * H - 0.5 * K
    Gt = 1j * H - 0.5 * K
    return G, Gt


def apply_L(G, Gt, c_ops, rho):
    out = G @ rho + rho @ Gt
    for c in c_ops:
        out += c @ rho @ c.conj().T
    return out


def norm_bound(G, Gt, c_ops):
    """Upper bound on the induced 1-norm of the column-stacked superoperator."""
    nb = np.abs(G).sum(axis=0).max() + np.abs(Gt).sum(axis=1).max()
    for c in c_ops:
        nb += np.abs(c).sum(axis=0).max() ** 2
    return nb


def evaluate(H, c_ops, rho0, obs, times, theta_target=16.0, m_cap=80, tol=1e-11, q_max=16):
    """Returns expect (K,T complex), number of L applications."""
    times = np.asarray(times, float)
    T = len(times)
    G, Gt = superop_parts(H, c_ops)
    nu = norm_bound(G, Gt, c_ops)
    out = np.zeros((len(obs), T), dtype=complex)
    rho = rho0.astype(complex).copy()
    for m, O in enumerate(obs):
        out[m, 0] = np.trace(O @ rho)
    napp = 0
    k = 0
    while k < T - 1:
        # choose q: as many output intervals as fit in theta_target
        q = 1
        while q < q_max and k + q < T - 1 and nu * (times[k + q + 1] - times[k]) <= theta_target:
            q += 1
        h_tot = times[k + q] - times[k]
        nsub = max(1, int(np.ceil(nu * h_tot / theta_target))) if q == 1 else 1
        # (q == 1 and the interval alone exceeds theta_target -> split into nsub substeps)
        for sub in range(nsub):
            h = h_tot / nsub
            last = (sub == nsub - 1)
            xs = np.array([(times[k + p] - times[k]) / h_tot for p in range(1, q + 1)]) if last and nsub == 1 else None
            v = rho.copy()
            acc = rho.copy()
            if xs is not None:
                y = np.zeros((len(obs), q), dtype=complex)
                for m, O in enumerate(obs):
                    y[m, :] = np.trace(O @ v)
                xp = np.ones(q)
            small = 0
            for j in range(m_cap):
                v = apply_L(G, Gt, c_ops, v) * (h / (j + 1))
                napp += 1
                acc += v
                if xs is not None:
                    xp = xp * xs
                    for m, O in enumerate(obs):
                        y[m, :] += xp * np.trace(O @ v)
                if np.abs(v).max() <= tol * max(1.0, np.abs(acc).max()):
                    small += 1
                    if small >= 2:
                        break
                else:
                    small = 0
            rho = acc
        if nsub == 1:
            out[:, k + 1:k + q + 1] = y
        else:
            for m, O in enumerate(obs):
                out[m, k + 1] = np.trace(O @ rho)
        k += q
    return out, napp
