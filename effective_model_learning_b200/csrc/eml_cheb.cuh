// Chebyshev (Bessel-coefficient) series of the Lindblad propagator -- the pieces shared by the kernels that use it.
//
//   L = L_s - ac,  At = L_s / R:   exp(tau L) rho = e^{-ac tau} sum_k J_k(tau R) T_k,   T_0 = rho, T_1 = 2 At rho,
//   T_{k+1} = 2 At T_k + T_{k-1}  (T_1 from (2 At) T_0 alone; T_k = 2 i^k T_k^cheb(At / i) rho for k >= 1).
// The terms do not depend on tau: one recurrence serves every output time of a macro step, only the 2x2 partial
// trace of each term is kept, and each output is a Bessel-weighted sum of them (weights by Miller's backward
// recurrence).  [-iR, iR] must enclose the imaginary extent of the numerical range of L_s; the series is truncated
// for the confocal ellipse (semi-axes A along the imaginary axis, B along the real axis) through the corner
// (ah, yext) of the enclosing rectangle: |J_k(tau R)| ((A + B) / R)^k behaves like J_k(tau (A + B)).
#pragma once
#include <cmath>
#include "eml_common.cuh"

namespace eml {

constexpr double CHEB_C1 = 9.25, CHEB_C0 = 5.0;   // |J_k(z)| < 1e-14 for k >= z + C1 z^(1/3) + C0   (0 <= z <= 2000)
constexpr double CHEB_HUMP = 14.0;                // cap on tau * B: partial sums stay below e^14 (what theta = 16 allowed the Taylor series)

__host__ __device__ __forceinline__ int cheb_terms(const double z) { return (int)ceil(z + CHEB_C1 * cbrt(z) + CHEB_C0); }

// largest z = tau (A + B) whose series fits `cap` terms (host)
inline double cheb_zcap_for(int cap) {
    double z = (double)cap;
    while (z > 0.0 && std::ceil(z + CHEB_C1 * std::cbrt(z) + CHEB_C0) > (double)cap) z -= 0.25;
    return z;
}

// Real range [dlo, dhi] of the numerical range of the dissipator and a bound dasym on its imaginary extent, from the
// per-site 4x4 real generators (Kronecker sum: the ranges add): row (ra, ca) has the diagonal entry
// jw[bp][0][ra][ca] - (kd[bp][ra] + kd[bp][ca]) / 2 and one off-diagonal entry jw[bp][1][ra][ca]; Gershgorin on the
// symmetric part, the antisymmetric part bounds the imaginary extent.
__device__ __forceinline__ void cheb_site_ranges(const double* jw, const double* kd, const int n_sites, double& dlo,
                                                 double& dhi, double& dasym) {
    dlo = 0.0; dhi = 0.0; dasym = 0.0;
    for (int bp = 0; bp < n_sites; ++bp) {
        double slo = 1e300, shi = -1e300, sas = 0.0;
#pragma unroll
        for (int rc = 0; rc < 4; ++rc) {
            const int ra = rc >> 1, ca = rc & 1;
            const double dg = jw[bp * 8 + rc] - 0.5 * (kd[bp * 2 + ra] + kd[bp * 2 + ca]);
            const double fa = jw[bp * 8 + 4 + rc], fb = jw[bp * 8 + 4 + (3 - rc)];
            const double rad = 0.5 * fabs(fa + fb);
            slo = fmin(slo, dg - rad); shi = fmax(shi, dg + rad); sas = fmax(sas, 0.5 * fabs(fa - fb));
        }
        dlo += slo; dhi += shi; dasym += sas;
    }
}

struct ChebPlan {
    double R;        // half-width of the Chebyshev interval on the imaginary axis
    double S;        // A + B: terms needed for a step tau are cheb_terms(tau * S)
    double ac;       // shift: L_s = L + ac is centred on the imaginary axis
    double tau_max;  // longest macro step
};

// spread: bound on lambda_max(H) - lambda_min(H); span: whole time span (to prefer enclosures that need no splitting)
__device__ __forceinline__ ChebPlan cheb_make_plan(const double spread, const double dlo, const double dhi, const double dasym,
                                                   const double span, const double zcap, const int kcap) {
    ChebPlan pl;
    pl.ac = -0.5 * (dlo + dhi);
    const double ah = 0.5 * (dhi - dlo), yext = spread + dasym;
    double Rb = 1.0, Sb = 1e300, Bb = 0.0;
    bool have_ok = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double delta = (i == 0) ? 0.03 : (i == 1) ? 0.06 : (i == 2) ? 0.12 : (i == 3) ? 0.25 : (i == 4) ? 0.5 : 1.0;
        const double Rc = fmax(fmax(yext * (1.0 + delta), ah), 1e-3 / fmax(span, 1e-300));
        const double cc = Rc * Rc - ah * ah - yext * yext, xr2 = ah * ah * Rc * Rc;
        const double disc = sqrt(cc * cc + 4.0 * xr2);
        const double u = (cc > 0.0) ? 2.0 * xr2 / (cc + disc) : 0.5 * (disc - cc);
        const double Bc = sqrt(u), Ac = sqrt(u + Rc * Rc);
        const bool ok = span * Bc <= CHEB_HUMP;
        if ((ok && !have_ok) || (ok == have_ok && Ac + Bc < Sb)) { Rb = Rc; Sb = Ac + Bc; Bb = Bc; have_ok = ok; }
    }
    pl.R = Rb; pl.S = Sb;
    double tau_max = zcap / Sb;
    if (Bb > 0.0) tau_max = fmin(tau_max, CHEB_HUMP / Bb);
    // the terms can grow like (S / R)^k: keep k log(S / R) below ~560 so that nothing overflows
    const double lnr = log(Sb / Rb);
    if (lnr * (double)kcap > 560.0) {
        const double kall = 560.0 / lnr;
        tau_max = fmin(tau_max, fmax(kall - CHEB_C1 * cbrt(kall) - CHEB_C0, 1.0) / Sb);
    }
    pl.tau_max = tau_max;
    return pl;
}

// ---- warp-cooperative versions for the prepare kernel (one warp per model): the scalar versions above cost a warp the same
//      number of instructions whether 1 or 32 lanes run them, so the independent pieces go to different lanes.
//      Same results as the scalar versions (same operations on the same values; ties keep the earlier candidate).
__device__ __forceinline__ void cheb_site_ranges_warp(const double* jw, const double* kd, const int n_sites, const int lane,
                                                      double& dlo, double& dhi, double& dasym) {
    // lane = 4 bp + rc for lane < 16: one row of one site's 4 x 4 generator
    const int bp = (lane >> 2) & 3, rc = lane & 3, ra = rc >> 1, ca = rc & 1;
    const bool on = lane < 16 && bp < n_sites;
    double lo = 1e300, hi = -1e300, as = 0.0;
    if (on) {
        const double dg = jw[bp * 8 + rc] - 0.5 * (kd[bp * 2 + ra] + kd[bp * 2 + ca]);
        const double fa = jw[bp * 8 + 4 + rc], fb = jw[bp * 8 + 4 + (3 - rc)];
        const double rad = 0.5 * fabs(fa + fb);
        lo = dg - rad; hi = dg + rad; as = 0.5 * fabs(fa - fb);
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {      // over the four rows of a site
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        as = fmax(as, __shfl_xor_sync(0xffffffffu, as, o));
    }
    if (!on) { lo = 0.0; hi = 0.0; as = 0.0; }
#pragma unroll
    for (int o = 4; o <= 8; o <<= 1) {      // the ranges of a Kronecker sum add
        lo += __shfl_xor_sync(0xffffffffu, lo, o);
        hi += __shfl_xor_sync(0xffffffffu, hi, o);
        as += __shfl_xor_sync(0xffffffffu, as, o);
    }
    dlo = __shfl_sync(0xffffffffu, lo, 0); dhi = __shfl_sync(0xffffffffu, hi, 0); dasym = __shfl_sync(0xffffffffu, as, 0);
}

__device__ __forceinline__ ChebPlan cheb_make_plan_warp(const double spread, const double dlo, const double dhi, const double dasym,
                                                        const double span, const double zcap, const int kcap, const int lane) {
    ChebPlan pl;
    pl.ac = -0.5 * (dlo + dhi);
    const double ah = 0.5 * (dhi - dlo), yext = spread + dasym;
    // lane i < 6 evaluates enclosure candidate i
    const int i = lane & 7;
    const double delta = (i == 0) ? 0.03 : (i == 1) ? 0.06 : (i == 2) ? 0.12 : (i == 3) ? 0.25 : (i == 4) ? 0.5 : 1.0;
    const double Rc = fmax(fmax(yext * (1.0 + delta), ah), 1e-3 / fmax(span, 1e-300));
    const double cc = Rc * Rc - ah * ah - yext * yext, xr2 = ah * ah * Rc * Rc;
    const double disc = sqrt(cc * cc + 4.0 * xr2);
    const double u = (cc > 0.0) ? 2.0 * xr2 / (cc + disc) : 0.5 * (disc - cc);
    double Bb = sqrt(u), Rb = Rc;
    double Sb = sqrt(u + Rc * Rc) + Bb;
    int ok = (i < 6 && span * Bb <= CHEB_HUMP) ? 1 : 0, who = i;
    if (i >= 6) Sb = 1e300;      // not a candidate: loses against every real one that is not ok, and ok is 0
    // first candidate that is ok with the smallest A + B, else the smallest A + B (ties: the earlier one) -- the scalar rule
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
        const int ok2 = __shfl_xor_sync(0xffffffffu, ok, o), who2 = __shfl_xor_sync(0xffffffffu, who, o);
        const double S2 = __shfl_xor_sync(0xffffffffu, Sb, o), R2 = __shfl_xor_sync(0xffffffffu, Rb, o);
        const double B2 = __shfl_xor_sync(0xffffffffu, Bb, o);
        const bool take = (ok2 > ok) || (ok2 == ok && (S2 < Sb || (S2 == Sb && who2 < who)));
        if (take) { ok = ok2; who = who2; Sb = S2; Rb = R2; Bb = B2; }
    }
    // NaN inputs: every comparison is false, lane 0 keeps candidate 0 (NaN) -- as the scalar version does
    Rb = __shfl_sync(0xffffffffu, Rb, 0); Sb = __shfl_sync(0xffffffffu, Sb, 0); Bb = __shfl_sync(0xffffffffu, Bb, 0);
    pl.R = Rb; pl.S = Sb;
    double tau_max = zcap / Sb;
    if (Bb > 0.0) tau_max = fmin(tau_max, CHEB_HUMP / Bb);
    const double lnr = log(Sb / Rb);
    if (lnr * (double)kcap > 560.0) {
        const double kall = 560.0 / lnr;
        tau_max = fmin(tau_max, fmax(kall - CHEB_C1 * cbrt(kall) - CHEB_C0, 1.0) / Sb);
    }
    pl.tau_max = tau_max;
    return pl;
}

// v * 2^-1 (hsub = 0x00100000) or v (hsub = 0) on the integer pipe; zeros and the very smallest numbers pass unchanged
__device__ __forceinline__ double scale_pow2(const double v, const int hsub) {
    const int hi = __double2hiint(v);
    const bool ok = (unsigned)(hi & 0x7fffffff) > 0x001fffffu;
    return __hiloint2double(ok ? hi - hsub : hi, __double2loint(v));
}

}  // namespace eml
