// d = 32 (n_tls = 5) evaluator for parity-conserving Hamiltonians, second generation: FOUR WARPS PER MODEL, EACH OWNING AN
// ORBIT OF TILES, only the upper triangle of rho is state.  Chebyshev / Bessel series as in eml_warp_mma.cu.
//
// Why: the first d = 32 kernel (eml_cta_mma.cu: warp w owns tile row w, every site stencil and every B fragment read from a
// published copy of the term, X'^dag through a second published copy, two __syncthreads per application) is bound by the
// shared-memory / shuffle data path, not by the FP64 pipe: ncu counts 1 509 shared-memory wavefronts per generator
// application at 0.88 wavefronts per cycle and SM, while its 128 DMMAs + ~400 DFMAs need only ~730 cycles
// (profiles/r01_ncu_full_cta_cheb_parity_realH_d32_summary.txt; tools/probes/mio_probe.cu: one LDS.128 = 4 cycles, one
// SHFL.32 = 1 cycle, one STS.64 = 2 cycles of ONE data path per SM, which DMMA does not share).
//
// Coordinates: x' = (parity of the physical index) << 4 | physical bits 3..0; tile index I = x' >> 3 = 2 p + b3, in-tile
// index = physical bits 2..0.  G = -iH - 1/2 sum c^dag c then has tiles only between tile rows of equal p.  A flip of one
// of the sites 0..2 or of the observed site maps tile (I, J) to (I ^ 2, J ^ 2), a flip of site 3 to (I ^ 3, J ^ 3): the 16
// tiles fall into four orbits under these maps, and rho = rho^dag leaves 10 unique tiles:
//     warp 0 ("D"):  the Hermitian tiles (0,0) (1,1) (2,2) (3,3)
//     warp 1:        m-tiles (0,1) (2,3)   and their conjugate transposes n = (1,0) (3,2)
//     warp 2:        m-tiles (0,2) (1,3)   and n = (2,0) (3,1)          (this orbit and D carry the partial trace)
//     warp 3:        m-tiles (0,3) (1,2)   and n = (3,0) (2,1)
// Every flip source of a warp's tiles is one of ITS OWN m / n tiles: the observed site and site 3 are pure register
// renaming, sites 0..2 are lane shuffles of the source tile -- no shared memory.  The m-tiles are the state; n = m^dag is
// kept as an operand (private 1 KB transposition tile per m-tile, conflict free, __syncwarp only).
//     off-diagonal (I, J):  w = J(v)_IJ + G[I,I] v[I,J] + G[I,I^1] v[I^1,J] + v[I,J] G[J,J]^dag + v[I,J^1] G[J,J^1]^dag
//                           B fragment of G[I,I] v[I,J] = conj(n) (own registers), A fragment of v[I,J] G^dag = m (own registers,
//                           the B side being the conjugates of G's A fragments); the two cross products read ONE LDS.128 per
//                           plane from the published term
//     diagonal (I, I):      X' = J(v)/2 + G[I,I] d[I] + G[I,I^1] v[I^1,I],  w = X' + X'^dag (private transposition tile)
// 32 DMMAs per warp and application (real H) as before, but per model and application: 10 x 24 DFMA of stencils instead
// of 16 x 24, 120 double shuffles + 64 LDS.128 + 64 STS.128 + 40 STS.64 (~670 data-path cycles) instead of ~1 500, and ONE
// CTA barrier (the published term is double buffered; the barrier sits between a warp's own work -- trace, stencils, the
// DMMAs on its own operands -- and the cross products, so waiting warps have work).
//
// Replaces qutip.mesolve + e_ops + loss for one model (reference basic_model.py:381-388, learning_chain.py:738-742).
#include <type_traits>
#include "eml_common.cuh"
#include "eml_cheb.cuh"

namespace eml {

constexpr int OMAXP = 192;
constexpr int OLD = 40;                 // row stride (doubles) of the published term: conflict-free 16-byte accesses over (g, t)
#ifndef EML_ORBIT_CTAS_PER_SM
#define EML_ORBIT_CTAS_PER_SM 2         // 3: <= 168 registers, single-buffered transposition tiles, 512-term trace table (73 KB of shared memory per CTA)
#endif
constexpr int ORB_CTAS = EML_ORBIT_CTAS_PER_SM;
constexpr int ORB_KCAP = ORB_CTAS >= 3 ? 512 : 1024;      // Chebyshev terms per macro step that fit the trace table
constexpr int ORB_XTBUF = ORB_CTAS >= 3 ? 1 : 2;          // buffers of the private transposition tiles (1: a second __syncwarp per application)
// kernel-variant switches for experiments (tools/build_variant.sh); the defaults are the measured best
#ifndef EML_ORBIT_PAIRBAR
#define EML_ORBIT_PAIRBAR 1             // per-application barrier among the two warps that exchange tiles (D <-> warp 1, warp 2 <-> warp 3) instead of the CTA
#endif
#ifndef EML_ORBIT_ROTATE
#define EML_ORBIT_ROTATE 1              // the second CTA of an SM runs its roles on the other schedulers (the D warp is the heaviest)
#endif

struct alignas(16) OrbitScratch {
    double Vs[2][2][32 * OLD];          // [buffer][re | im]: the whole current term, row-major
    double xt[ORB_XTBUF][10][128];              // private transposition tiles [buffer][slot][re 64 | im 64]; D: slots 0..3, warp r: 2 + 2 r, 3 + 2 r
    double jw[40];                      // [bitpos 0..4][type (0 diag, 1 flip)][ra][ca]      (from the prepare kernel's record)
    double gdiag[32];                   // real H: diagonal of G.re = -1/2 sum c^dag c by PHYSICAL index
    double lossw[4];
    unsigned int fetch;
    int flag;
    alignas(16) double tr[(ORB_KCAP + 4) * 4];      // partial traces of the terms: rho00, rho11 (warp 0), Re rho01, Im rho01 (warp 2)
    double cJ[ORB_KCAP + 2];
};

// (not volatile, and the products written tile by tile: pinning a round-robin issue order over the accumulators by hand, or
// fetching all cross operands before the first cross DMMA, was measured 7 % slower than the compiler's own schedule)
__device__ __forceinline__ void dmma_o(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ double neg_o(double v) { return __hiloint2double(__double2hiint(v) ^ 0x80000000, __double2loint(v)); }
__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ void cta_bar() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ int phys32(const int x) { return x ^ (((x ^ (x >> 1) ^ (x >> 2) ^ (x >> 3)) & 1) << 4); }
// transposition tile: element (R, C) of a plane at 2 pi + (R & 1), pi = 8 h + (a & 1) + 2 ((C >> 1) ^ h), a = R >> 1,
// h = 2 (C & 1) + (a >> 1): STS.64 of the accumulator fragments (R = g, C = 2t + e) and LDS.128 of the transposed pairs
// (R = 2t, 2t + 1; C = g) are both bank-conflict free (eml_warp_mma.cu)
__device__ __forceinline__ int xt_addr_o(const int R, const int C) {
    const int a = R >> 1, h = 2 * (C & 1) + (a >> 1);
    return 2 * (8 * h + (a & 1) + 2 * ((C >> 1) ^ h)) + (R & 1);
}

// Everything a warp needs across the series of one model, by role.
struct OrbitCtx {
    const DevProblem* P;
    OrbitScratch* S;
    long long b;
    const double* tgt;
    double* expect;
    int lane, g, t, w;
    double Rb, Sb, ac, tau_max, sc;
};

// stencil weights of one tile (I, J) for this lane: diagonal weight W[e], observed-site flip F4[e], site-3 flip F3 (scaled by
// `mul`); REAL_H folds the diagonal of G.re in
template <bool REAL_H>
__device__ __forceinline__ void tile_weights(const OrbitScratch& S, const int I, const int J, const int g, const int t, const double acs,
                                             const double mul, double (&W)[2], double (&F4)[2], double& F3) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int row = phys32(8 * I + g), col = phys32(8 * J + 2 * t + e);
        double x = S.jw[4 * 8 + (row >> 4) * 2 + (col >> 4)] + S.jw[3 * 8 + (I & 1) * 2 + (J & 1)] +
                   S.jw[2 * 8 + (g >> 2) * 2 + (t >> 1)] + S.jw[1 * 8 + ((g >> 1) & 1) * 2 + (t & 1)] + S.jw[0 * 8 + (g & 1) * 2 + e];
        if (REAL_H) x += S.gdiag[row] + S.gdiag[col];
        W[e] = (x + acs) * mul;
        F4[e] = S.jw[4 * 8 + 4 + (row >> 4) * 2 + (col >> 4)] * mul;
    }
    F3 = S.jw[3 * 8 + 4 + (I & 1) * 2 + (J & 1)] * mul;
}

// flip stencils of the sites 0, 1, 2 (lane shuffles of the source tile s2), of the observed site (s2, same position) and of
// site 3 (s3, same position) onto (xr, xi)
__device__ __forceinline__ void flips_o(double (&xr)[2], double (&xi)[2], const double (&s2r)[2], const double (&s2i)[2],
                                        const double (&s3r)[2], const double (&s3i)[2], const double (&F4)[2], const double F3,
                                        const double f2, const double f1, const double (&f0)[2]) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        xr[e] = fma(F4[e], s2r[e], xr[e]); xi[e] = fma(F4[e], s2i[e], xi[e]);
        xr[e] = fma(F3, s3r[e], xr[e]);    xi[e] = fma(F3, s3i[e], xi[e]);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) { xr[e] = fma(f2, shx(s2r[e], 18), xr[e]); xi[e] = fma(f2, shx(s2i[e], 18), xi[e]); }
#pragma unroll
    for (int e = 0; e < 2; ++e) { xr[e] = fma(f1, shx(s2r[e], 9), xr[e]); xi[e] = fma(f1, shx(s2i[e], 9), xi[e]); }
#pragma unroll
    for (int e = 0; e < 2; ++e) { xr[e] = fma(f0[e], shx(s2r[e ^ 1], 4), xr[e]); xi[e] = fma(f0[e], shx(s2i[e ^ 1], 4), xi[e]); }
}

// X += G_tile * B,  B = conj(b): b = this lane's C-layout pair of the tile whose conjugate transpose is the operand
template <bool REAL_H>
__device__ __forceinline__ void mma_G_conjb(double (&xr)[2], double (&xi)[2], const double (&a_re)[2], const double (&a_im)[2],
                                            const double (&br)[2], const double (&bi)[2]) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        if (!REAL_H) {
            dmma_o(xr[0], xr[1], a_re[ks], br[ks]);
            dmma_o(xi[0], xi[1], a_re[ks], neg_o(bi[ks]));
        }
        dmma_o(xr[0], xr[1], a_im[ks], bi[ks]);
        dmma_o(xi[0], xi[1], a_im[ks], br[ks]);
    }
}
// X += A * G_tile^dag,  A = this lane's C-layout pair (ar, ai) of a tile of the term; the B side is the conjugate of G's A fragment
template <bool REAL_H>
__device__ __forceinline__ void mma_v_Gdag(double (&xr)[2], double (&xi)[2], const double (&ar)[2], const double (&ai)[2],
                                           const double (&g_re)[2], const double (&g_im)[2]) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        if (!REAL_H) {
            dmma_o(xr[0], xr[1], ar[ks], g_re[ks]);
            dmma_o(xi[0], xi[1], ai[ks], g_re[ks]);
        }
        dmma_o(xr[0], xr[1], ai[ks], g_im[ks]);
        dmma_o(xi[0], xi[1], neg_o(ar[ks]), g_im[ks]);
    }
}

// observables and loss of one output time from the reduced density matrix r = (rho00, rho11, Re rho01, Im rho01)
__device__ __forceinline__ void finish_output_o(const OrbitCtx& C, const int k, const double (&r)[4], double& my_loss) {
    const DevProblem& P = *C.P;
    const int T = P.n_times, K = P.n_obs;
    for (int m = 0; m < K; ++m) {
        const double* O = P.obs + m * 8;
        const double ere = O[0] * r[0] + O[6] * r[1] + (O[2] + O[4]) * r[2] + (O[3] - O[5]) * r[3];
        const double eim = O[1] * r[0] + O[7] * r[1] + (O[3] + O[5]) * r[2] + (O[4] - O[2]) * r[3];
        if (C.expect != nullptr) {
            double* e = C.expect + (((size_t)C.b * K + m) * T + k) * 2;
            e[0] = ere; e[1] = eim;
        }
        if (C.tgt != nullptr) {
            const double dre = ere - C.tgt[((size_t)m * T + k) * 2], dim = eim - C.tgt[((size_t)m * T + k) * 2 + 1];
            my_loss += dre * dre + dim * dim;
        }
    }
}

// The whole series of one model as seen by one warp.  ROLE 0: the four diagonal tiles; ROLE 1..3: two m-tiles A, B and
// their conjugate transposes.  All four roles execute the same sequence of CTA barriers.
template <bool REAL_H, int ROLE>
__device__ __forceinline__ void orbit_series(const OrbitCtx& C, const double (&aG_re)[4][2][2], const double (&aG_im)[4][2][2],
                                             double& my_loss, int& n_apply, bool& capped) {
    constexpr bool DG = (ROLE == 0);
    constexpr int NT = DG ? 4 : 2;                                  // state tiles of this warp
    // tile coordinates: D: (i, i); off-diagonal roles: A, B
    constexpr int TI[4] = {DG ? 0 : 0, DG ? 1 : (ROLE == 1 ? 2 : 1), 2, 3};
    constexpr int TJ[4] = {DG ? 0 : ROLE, DG ? 1 : (ROLE == 1 ? 3 : (ROLE == 2 ? 3 : 2)), 2, 3};
    const DevProblem& P = *C.P;
    OrbitScratch& S = *C.S;
    const int lane = C.lane, g = C.g, t = C.t;
    const int T = P.n_times;
    const int gp = g & 1;
    const bool diag_lane = (t == (g >> 1));
    const int dl4 = 4 * (g ^ 4) + ((g ^ 4) >> 1), dl2 = 4 * (g ^ 2) + ((g ^ 2) >> 1), dl1 = 4 * (g ^ 1) + ((g ^ 1) >> 1);
    const bool pg = (g ^ (g >> 1) ^ (g >> 2)) & 1;
    const double Rb = C.Rb, Sb = C.Sb, ac = C.ac, tau_max = C.tau_max;

    // ---- weights (generator -> 2 (L + ac) / R; the Hermitian tiles of D take them halved: X' + X'^dag doubles)
    const double mul = DG ? 0.5 * C.sc : C.sc;
    double W[NT][2], F4[NT][2], F3[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) tile_weights<REAL_H>(S, TI[i], TJ[i], g, t, ac, mul, W[i], F4[i], F3[i]);
    const double f2 = S.jw[2 * 8 + 4 + (g >> 2) * 2 + (t >> 1)] * mul;
    const double f1 = S.jw[1 * 8 + 4 + ((g >> 1) & 1) * 2 + (t & 1)] * mul;
    const double f0[2] = {S.jw[0 * 8 + 4 + (g & 1) * 2 + 0] * mul, S.jw[0 * 8 + 4 + (g & 1) * 2 + 1] * mul};

    // ---- state at the step start: rho(t_k), this warp's unique tiles
    double acc_re[NT][2], acc_im[NT][2];
    {
        const int obs_site = P.obs_site;
        auto pos = [&](int s) { return (s == obs_site) ? 4 : (s < obs_site ? 3 - s : 4 - s); };
        auto orig_index = [&](int x) {
            int o = 0;
#pragma unroll
            for (int s = 0; s < 5; ++s) o |= ((x >> pos(s)) & 1) << (4 - s);
            return o;
        };
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = orig_index(phys32(8 * TI[i] + g)), c = orig_index(phys32(8 * TJ[i] + 2 * t + e));
                acc_re[i][e] = P.rho0[2 * (r * 32 + c)];
                acc_im[i][e] = P.rho0[2 * (r * 32 + c) + 1];
            }
    }

    // private transposition tiles and the published term
    const int slot0 = DG ? 0 : 2 + 2 * ROLE;
    const int xs0 = xt_addr_o(g, 2 * t), xs1 = xt_addr_o(g, 2 * t + 1), xl = xt_addr_o(2 * t, g);
    auto xt_put = [&](const int buf, const int i, const double (&xr)[2], const double (&xi)[2]) {
        EML_CHECK(buf >= 0 && buf < ORB_XTBUF && slot0 + i >= 0 && slot0 + i < 10 && xs0 < 64 && xs1 < 64 && xl + 1 < 64);
        double* const X = S.xt[buf][slot0 + i];
        X[xs0] = xr[0]; X[xs1] = xr[1]; X[64 + xs0] = xi[0]; X[64 + xs1] = xi[1];
    };
    // (re, im) of the conjugate transpose: (cr.x, cr.y), (-ci.x, -ci.y)
    auto xt_get = [&](const int buf, const int i, double2& cr, double2& ci) {
        const double* const X = S.xt[buf][slot0 + i];
        cr = *reinterpret_cast<const double2*>(X + xl);
        ci = *reinterpret_cast<const double2*>(X + 64 + xl);
    };
    auto publish = [&](const int vb, const int I, const int J, const double (&xr)[2], const double (&xi)[2]) {
        EML_CHECK((vb == 0 || vb == 1) && I >= 0 && I < 4 && J >= 0 && J < 4);
        *reinterpret_cast<double2*>(&S.Vs[vb][0][(8 * I + g) * OLD + 8 * J + 2 * t]) = make_double2(xr[0], xr[1]);
        *reinterpret_cast<double2*>(&S.Vs[vb][1][(8 * I + g) * OLD + 8 * J + 2 * t]) = make_double2(xi[0], xi[1]);
    };
    auto fetch = [&](const int vb, const int I, const int J, double (&xr)[2], double (&xi)[2]) {
        const double2 r = *reinterpret_cast<const double2*>(&S.Vs[vb][0][(8 * I + g) * OLD + 8 * J + 2 * t]);
        const double2 i = *reinterpret_cast<const double2*>(&S.Vs[vb][1][(8 * I + g) * OLD + 8 * J + 2 * t]);
        xr[0] = r.x; xr[1] = r.y; xi[0] = i.x; xi[1] = i.y;
    };

    // ---- partial trace of a term into the table (warp 0: rho00, rho11; warp 2: rho01).  Only the 8 lanes 4g + (g >> 1) hold
    //      partial-trace elements (at e = g & 1); the observed site's bit of a row in tile I is (I >> 1) ^ parity(g) ^ (I & 1).
    auto trace_put = [&](const double (&vr)[NT][2], const double (&vi)[NT][2], const int k) {
        if constexpr (ROLE == 0 || ROLE == 2) {
            double v0, v1;
            if constexpr (DG) {
                const double a0 = gp ? vr[0][1] : vr[0][0], a1 = gp ? vr[1][1] : vr[1][0];
                const double a2 = gp ? vr[2][1] : vr[2][0], a3 = gp ? vr[3][1] : vr[3][0];
                const double s03 = a0 + a3, s12 = a1 + a2;      // tiles 0, 3: observed bit = parity(g); tiles 1, 2: the opposite
                v0 = pg ? s12 : s03; v1 = pg ? s03 : s12;
            } else {
                // tile A = (0,2): an element of rho01 where parity(g) = 0, of rho10 = conj(rho01) otherwise; tile B = (1,3): the opposite
                const double ar = gp ? vr[0][1] : vr[0][0], ai = gp ? vi[0][1] : vi[0][0];
                const double br = gp ? vr[1][1] : vr[1][0], bi = gp ? vi[1][1] : vi[1][0];
                v0 = ar + br; v1 = pg ? bi - ai : ai - bi;
            }
            const bool hb = g & 4;
            double kk = (hb ? v1 : v0) + __shfl_sync(0xffffffffu, hb ? v0 : v1, dl4);
            kk += __shfl_sync(0xffffffffu, kk, dl2);
            kk += __shfl_sync(0xffffffffu, kk, dl1);
            EML_CHECK(k >= 0 && k <= ORB_KCAP);
            if (diag_lane && (g & 3) == 0) S.tr[k * 4 + (DG ? 0 : 2) + (g >> 2)] = kk;
        }
    };

    // first output time: the initial state
    trace_put(acc_re, acc_im, 0);
    cta_bar();
    if (C.w == 0 && lane == 0) {
        const double r[4] = {S.tr[0], S.tr[1], S.tr[2], S.tr[3]};
        finish_output_o(C, 0, r, my_loss);
    }

    int kidx = 0;
    while (kidx < T - 1) {           // every scalar below is computed identically by all threads
        const double t0 = P.times[kidx];
        int q = 0;      // output times within tau_max of t0: 32 candidates per ballot
        for (;;) {
            const int cand = kidx + q + 1 + lane;
            const unsigned in = __ballot_sync(0xffffffffu, cand <= T - 1 && P.times[min(cand, T - 1)] - t0 <= tau_max);
            const int run = __ffs(~in) - 1;
            if (run >= 0) { q += run; break; }
            q += 32;
        }
        int nsub = 1;
        if (q == 0) { nsub = (int)ceil((P.times[kidx + 1] - t0) / tau_max); q = 1; }
        const double tau_end = (P.times[kidx + q] - t0) / nsub;
        const bool final_step = (kidx + q == T - 1);
        if (!(tau_end * Rb >= 1e-40)) {     // repeated time points: the state does not move
            cta_bar();
            trace_put(acc_re, acc_im, 0);
            cta_bar();
            const double r[4] = {S.tr[0], S.tr[1], S.tr[2], S.tr[3]};
            for (int p = threadIdx.x; p < q; p += 128) finish_output_o(C, kidx + 1 + p, r, my_loss);
            kidx += q;
            continue;
        }
        int Kc = cheb_terms(tau_end * Sb);
        if (Kc > ORB_KCAP) Kc = ORB_KCAP;
        for (int sub = 0; sub < nsub; ++sub) {
            const bool last = (sub == nsub - 1);
            const bool need_acc = !(final_step && last);
            double cscale = 0.0;
            cta_bar();     // the previous step's readers of tr / cJ / the published term are done
            if (need_acc) {
                const double tz = 2.0 / (tau_end * Rb);
                double jn = 1.0, jn1 = 0.0, ssum = 0.0;
                for (int k = Kc; k >= 1; --k) {
                    EML_CHECK(k <= ORB_KCAP + 1);
                    if ((k & 127) == (int)threadIdx.x) S.cJ[k] = jn;
                    if (!(k & 1)) ssum += jn;
                    const double jm = fma((double)k * tz, jn, -jn1);
                    jn1 = jn; jn = jm;
                }
                if (threadIdx.x == 0) S.cJ[0] = jn;
                cscale = exp(-ac * tau_end) / (2.0 * ssum + jn);
            }
            // ---- term 0 = the state; previous term = 0
            double v_re[NT][2], v_im[NT][2], p_re[NT][2], p_im[NT][2];
            [[maybe_unused]] double n_re[2][2], n_im[2][2];      // off-diagonal roles: n = m^dag (true values)
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    v_re[i][e] = acc_re[i][e]; v_im[i][e] = acc_im[i][e];
                    p_re[i][e] = 0.0; p_im[i][e] = 0.0;
                    acc_re[i][e] = 0.0; acc_im[i][e] = 0.0;
                }
            if constexpr (!DG) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; ++i) xt_put(ORB_XTBUF - 1, i, v_re[i], v_im[i]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    double2 cr, ci;
                    xt_get(ORB_XTBUF - 1, i, cr, ci);
                    n_re[i][0] = cr.x; n_re[i][1] = cr.y; n_im[i][0] = -ci.x; n_im[i][1] = -ci.y;
                }
                if constexpr (ORB_XTBUF == 1) __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                publish(0, TI[i], TJ[i], v_re[i], v_im[i]);
                if constexpr (!DG) publish(0, TJ[i], TI[i], n_re[i], n_im[i]);
            }
            if (need_acc) cta_bar();      // cJ is complete (the loop's own barrier only orders the published term)

            auto application = [&](const int k, auto na_tag, auto first_tag) {
                constexpr bool NA = decltype(na_tag)::value, FIRST = decltype(first_tag)::value;
                const int vb = k & 1;
                trace_put(v_re, v_im, k);
                if constexpr (NA) {
                    const double c = S.cJ[k] * cscale;
#pragma unroll
                    for (int i = 0; i < NT; ++i)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            acc_re[i][e] = fma(c, v_re[i][e], acc_re[i][e]);
                            acc_im[i][e] = fma(c, v_im[i][e], acc_im[i][e]);
                        }
                }
                // ---- own work: diagonal weight, previous term, flips, the products on the warp's own operands
                double x_re[NT][2], x_im[NT][2];
#pragma unroll
                for (int i = 0; i < NT; ++i) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        x_re[i][e] = fma(W[i][e], v_re[i][e], p_re[i][e]);
                        x_im[i][e] = fma(W[i][e], v_im[i][e], p_im[i][e]);
                    }
                    if constexpr (DG) {
                        // sources: tiles (I ^ 2, I ^ 2) and (I ^ 3, I ^ 3)
                        flips_o(x_re[i], x_im[i], v_re[i ^ 2], v_im[i ^ 2], v_re[i ^ 3], v_im[i ^ 3], F4[i], F3[i], f2, f1, f0);
                        mma_G_conjb<REAL_H>(x_re[i], x_im[i], aG_re[i][i & 1], aG_im[i][i & 1], v_re[i], v_im[i]);
                    } else {
                        // (I ^ 2, J ^ 2) and (I ^ 3, J ^ 3):  role 1: m[other], n[other];  role 2: n[same], n[other];  role 3: n[other], n[same]
                        const int o = i ^ 1;
                        if constexpr (ROLE == 1) flips_o(x_re[i], x_im[i], v_re[o], v_im[o], n_re[o], n_im[o], F4[i], F3[i], f2, f1, f0);
                        if constexpr (ROLE == 2) flips_o(x_re[i], x_im[i], n_re[i], n_im[i], n_re[o], n_im[o], F4[i], F3[i], f2, f1, f0);
                        if constexpr (ROLE == 3) flips_o(x_re[i], x_im[i], n_re[o], n_im[o], n_re[i], n_im[i], F4[i], F3[i], f2, f1, f0);
                        const int I = TI[i], J = TJ[i];
                        mma_G_conjb<REAL_H>(x_re[i], x_im[i], aG_re[I][I & 1], aG_im[I][I & 1], n_re[i], n_im[i]);      // G[I,I] v[I,J]
                        mma_v_Gdag<REAL_H>(x_re[i], x_im[i], v_re[i], v_im[i], aG_re[J][J & 1], aG_im[J][J & 1]);      // v[I,J] G[J,J]^dag
                    }
                }
                // ---- the published term k is complete.  D and role 1 read only each other's tiles, roles 2 and 3 likewise: the two
                //      pairs synchronise separately (named barriers of 64 threads) and may drift apart within a macro step
                if constexpr (EML_ORBIT_PAIRBAR) asm volatile("bar.sync %0, 64;" :: "n"(1 + (ROLE >> 1)) : "memory");
                else cta_bar();
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const int I = TI[i], J = TJ[i];
                    double b_re[2], b_im[2];
                    fetch(vb, J, I ^ 1, b_re, b_im);                                                             // G[I,I^1] v[I^1,J]: conj of tile (J, I^1)
                    mma_G_conjb<REAL_H>(x_re[i], x_im[i], aG_re[I][(I & 1) ^ 1], aG_im[I][(I & 1) ^ 1], b_re, b_im);
                    if constexpr (!DG) {
                        double a_re[2], a_im[2];
                        fetch(vb, I, J ^ 1, a_re, a_im);                                                         // v[I,J^1] G[J,J^1]^dag
                        mma_v_Gdag<REAL_H>(x_re[i], x_im[i], a_re, a_im, aG_re[J][(J & 1) ^ 1], aG_im[J][(J & 1) ^ 1]);
                    }
                }
                // ---- next term.  T_2 = (2 At) T_1 + 2 T_0, afterwards T_{k+1} = (2 At) T_k + T_{k-1}; D keeps the previous term halved
                const int xb = (ORB_XTBUF == 2) ? (k & 1) : 0;
#pragma unroll
                for (int i = 0; i < NT; ++i) xt_put(xb, i, x_re[i], x_im[i]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    double2 cr, ci;
                    xt_get(xb, i, cr, ci);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double pm = DG ? (FIRST ? 1.0 : 0.5) : (FIRST ? 2.0 : 1.0);
                        p_re[i][e] = pm * v_re[i][e]; p_im[i][e] = pm * v_im[i][e];
                    }
                    if constexpr (DG) {
                        v_re[i][0] = x_re[i][0] + cr.x; v_re[i][1] = x_re[i][1] + cr.y;
                        v_im[i][0] = x_im[i][0] - ci.x; v_im[i][1] = x_im[i][1] - ci.y;
                    } else {
                        v_re[i][0] = x_re[i][0]; v_re[i][1] = x_re[i][1]; v_im[i][0] = x_im[i][0]; v_im[i][1] = x_im[i][1];
                        n_re[i][0] = cr.x; n_re[i][1] = cr.y; n_im[i][0] = -ci.x; n_im[i][1] = -ci.y;
                    }
                }
                if constexpr (ORB_XTBUF == 1) __syncwarp();      // the single buffer is rewritten by the next application
                if (k + 1 < Kc) {
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        publish(vb ^ 1, TI[i], TJ[i], v_re[i], v_im[i]);
                        if constexpr (!DG) publish(vb ^ 1, TJ[i], TI[i], n_re[i], n_im[i]);
                    }
                }
            };
            auto run_terms = [&](auto na_tag) {
                application(0, na_tag, std::true_type{});
                for (int k = 1; k < Kc; ++k) application(k, na_tag, std::false_type{});
                trace_put(v_re, v_im, Kc);          // the last term
                if constexpr (decltype(na_tag)::value) {
                    const double c = S.cJ[Kc] * cscale;
#pragma unroll
                    for (int i = 0; i < NT; ++i)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            acc_re[i][e] = fma(c, v_re[i][e], acc_re[i][e]);
                            acc_im[i][e] = fma(c, v_im[i][e], acc_im[i][e]);
                        }
                }
            };
            if (need_acc) run_terms(std::true_type{}); else run_terms(std::false_type{});
            {   // a non-finite or absurd last term means the enclosure failed
                const double chk = fabs(v_re[0][0]) + fabs(v_im[0][0]) + fabs(v_re[NT - 1][1]);
                if (!(chk < 1e300)) capped = true;
            }
            n_apply += Kc;
            cta_bar();     // the trace table is complete
            if (last) {
                // ---- outputs of the step: warp w takes the blocks of 32 outputs w, w + 4, ...  Bessel weights J_k(tau_p R) by
                //      Miller's backward recurrence, terms in groups of four (k a multiple of 4; the table is zero beyond Kc)
                if (threadIdx.x < 12) S.tr[(Kc + 1) * 4 + threadIdx.x] = 0.0;
                cta_bar();
                for (int pbase = 32 * C.w; pbase < q; pbase += 128) {
                    const int p = q - 1 - pbase - lane;      // blocks are cut from the end of the list: the partial block is the cheapest
                    const bool act = p >= 0;
                    const double tau = act ? ((nsub == 1) ? P.times[kidx + 1 + p] - t0 : tau_end) : 0.0;
                    const double z = tau * Rb;
                    const bool direct = !(z >= 1e-40);
                    int Kp = 0;
                    if (act && !direct) { Kp = cheb_terms(tau * Sb); if (Kp > Kc) Kp = Kc; }
                    const int kmax = __reduce_max_sync(0xffffffffu, Kp);
                    const double tz = direct ? 0.0 : 2.0 / z;
                    double a0 = 0.0, a2 = 0.0, a3 = 0.0, ssum = 0.0, jn = 0.0, jn1 = 0.0;
                    int k = (kmax + 3) & ~3;
                    double kd = (double)k;
                    while (k >= 4) {
                        double cm = kd * tz;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int kk = k - j;
                            EML_CHECK(kk >= 1 && kk <= Kc + 3 && kk <= ORB_KCAP + 3);
                            const double2 ta = *reinterpret_cast<const double2*>(&S.tr[kk * 4]);
                            const double2 tb = *reinterpret_cast<const double2*>(&S.tr[kk * 4 + 2]);
                            jn = __hiloint2double(kk == Kp ? 0x3ff00000 : __double2hiint(jn), __double2loint(jn));   // exactly 0.0 before the start
                            a0 = fma(jn, ta.x, a0); a2 = fma(jn, tb.x, a2); a3 = fma(jn, tb.y, a3);
                            if (!(j & 1)) ssum += jn;
                            const double jm = fma(cm, jn, -jn1);
                            jn1 = jn; jn = jm;
                            if (j < 3) cm -= tz;
                        }
                        k -= 4; kd -= 4.0;
                    }
                    const double2 ta = *reinterpret_cast<const double2*>(&S.tr[0]);
                    const double2 tb = *reinterpret_cast<const double2*>(&S.tr[2]);
                    double r[4];
                    if (direct) {
                        r[0] = ta.x; r[1] = ta.y; r[2] = tb.x; r[3] = tb.y;
                    } else {
                        const double nrm = exp(-ac * tau) / (2.0 * ssum + jn);
                        // the generator is trace preserving: rho11 = Tr rho(step start) - rho00 saves a quarter of the sums
                        r[0] = fma(jn, ta.x, a0) * nrm; r[1] = (ta.x + ta.y) - r[0];
                        r[2] = fma(jn, tb.x, a2) * nrm; r[3] = fma(jn, tb.y, a3) * nrm;
                    }
                    if (act) finish_output_o(C, kidx + 1 + p, r, my_loss);
                }
            }
        }
        kidx += q;
    }
}

// ---- prepare kernel (one CTA of four warps per model, persistent): K1 (reference basic_model.py:152-255) -- G = -iH - 1/2 sum c^dag c,
//      the real jump weights -- and the Chebyshev enclosure of eml_cheb.cuh (Gershgorin, 2 ||(H - mu)^8||_1^(1/8) on the DMMA pipe,
//      per-site dissipator ranges, six candidate ellipses), at five CTAs per SM instead of inside the evaluator's two.
//      ONE RECORD per model: [NF][32] A fragments of G per lane, [40] jw, [32] diagonal of G.re, [4] R, S, ac, tau_max (R = NaN: not
//      evaluable), and the number of generator applications the evaluation will need (the cost the evaluator's CTAs are ordered by).
template <bool REAL_H>
__host__ __device__ constexpr int orbit_rec_doubles() { return (REAL_H ? 16 : 32) * 32 + 40 + 32 + 4; }
constexpr int PREP32_MAXP = OMAXP;
template <bool REAL_H>
struct alignas(16) Prep32Scratch {
    double G[2][32 * OLD];                  // G.re, G.im in PHYSICAL coordinates
    double M[REAL_H ? 3 : 4][32 * 36];      // squaring scratch (real H: M[0] and M[2] only)
    double par[PREP32_MAXP];
    double jw[40];
    double kd[10];
    double gdiag[32];
    double colsum[32], rowlo[32], diagh[32];
    double misc[4];
};

template <bool REAL_H>
__global__ void __launch_bounds__(128, 4)
prepare32_kernel(const DevProblem P, const long long n_models, const double* __restrict__ params, double* __restrict__ rec,
                 float* __restrict__ cost, const double cheb_zcap) {
    constexpr int N = 5, D = 32, NF = REAL_H ? 16 : 32, REC = orbit_rec_doubles<REAL_H>();
    extern __shared__ __align__(16) unsigned char prep32_smem_raw[];
    Prep32Scratch<REAL_H>& S = *reinterpret_cast<Prep32Scratch<REAL_H>*>(prep32_smem_raw);
    const int tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int T = P.n_times;
    const int obs_site = P.obs_site;
    auto pos = [&](int s) { return (s == obs_site) ? N - 1 : (s < obs_site ? N - 2 - s : N - 1 - s); };
    for (long long b = blockIdx.x; b < n_models; b += gridDim.x) {
        __syncthreads();     // the previous model's readers of the scratch are done
        // ---- K1 (reference basic_model.py:152-255): parameters, G = -iH - 1/2 sum c^dag c in PHYSICAL coordinates (row stride OLD)
        double* G_re = S.G[0];
        double* G_im = S.G[1];
        for (int i = tid; i < P.n_params; i += 128) S.par[i] = params[b * P.n_params + i];
        for (int i = tid; i < 32 * OLD; i += 128) { G_re[i] = 0.0; G_im[i] = 0.0; }
        __syncthreads();
        if (tid < D) {      // thread r owns row r of G (fixed summation order: results are reproducible bit for bit); zero entries skipped
            const int r = tid;
            for (int ti = 0; ti < P.n_terms; ++ti) {
                const EmlTerm tm = P.terms[ti];
                const double v = S.par[tm.param];
                if (v == 0.0) continue;
                const int sa = pos(tm.site_a);
                const int ra = (r >> sa) & 1;
                if (tm.kind == EML_TERM_ENERGY) {
                    for (int a = 0; a < 2; ++a) {
                        const cplx A = ld_op(P.op_table, tm.op_a, ra, a);
                        const int c = (r & ~(1 << sa)) | (a << sa);
                        if (A.im != 0.0) G_re[r * OLD + c] += v * A.im;        // -i h
                        if (A.re != 0.0) G_im[r * OLD + c] -= v * A.re;
                    }
                } else if (tm.kind == EML_TERM_COUPLING) {
                    const int sb = pos(tm.site_b);
                    const int rb = (r >> sb) & 1;
                    const double v2 = v * v;                  // strength on both factors (reference basic_model.py:245,247)
                    for (int a = 0; a < 2; ++a)
                        for (int bb = 0; bb < 2; ++bb) {
                            const cplx AB = cmul(ld_op(P.op_table, tm.op_a, ra, a), ld_op(P.op_table, tm.op_b, rb, bb));
                            const int c = (r & ~((1 << sa) | (1 << sb))) | (a << sa) | (bb << sb);
                            if (AB.im != 0.0) G_re[r * OLD + c] += v2 * AB.im;
                            if (AB.re != 0.0) G_im[r * OLD + c] -= v2 * AB.re;
                        }
                } else {
                    for (int a = 0; a < 2; ++a) {
                        cplx M = {0.0, 0.0};
                        for (int z = 0; z < 2; ++z) {
                            const cplx m = cmul(cconj(ld_op(P.op_table, tm.op_a, z, ra)), ld_op(P.op_table, tm.op_a, z, a));
                            M.re += m.re; M.im += m.im;
                        }
                        const int c = (r & ~(1 << sa)) | (a << sa);
                        if (M.re != 0.0) G_re[r * OLD + c] -= 0.5 * v * M.re;
                        if (M.im != 0.0) G_im[r * OLD + c] -= 0.5 * v * M.im;
                    }
                }
            }
        }
        if (tid >= 64 && tid < 64 + 40) {
            const int e_ = tid - 64;
            const int bp = e_ >> 3, type = (e_ >> 2) & 1, ra = (e_ >> 1) & 1, ca = e_ & 1;
            const int a = type ? 1 - ra : ra, bb = type ? 1 - ca : ca;
            double wgt = 0.0;
            for (int ti = 0; ti < P.n_terms; ++ti) {
                const EmlTerm tm = P.terms[ti];
                if (tm.kind != EML_TERM_LINDBLAD || pos(tm.site_a) != bp) continue;
                const double v = S.par[tm.param];
                if (v == 0.0) continue;
                const cplx x = cmul(ld_op(P.op_table, tm.op_a, ra, a), cconj(ld_op(P.op_table, tm.op_a, ca, bb)));
                wgt += v * x.re;
            }
            S.jw[e_] = wgt;
        }
        if (tid >= 32 && tid < 42) {
            const int bp = (tid - 32) >> 1, ra = (tid - 32) & 1;
            double kdv = 0.0;
            for (int ti = 0; ti < P.n_terms; ++ti) {
                const EmlTerm tm = P.terms[ti];
                if (tm.kind != EML_TERM_LINDBLAD || pos(tm.site_a) != bp) continue;
                const double v = S.par[tm.param];
                if (v == 0.0) continue;
                for (int z = 0; z < 2; ++z) {
                    const cplx m = ld_op(P.op_table, tm.op_a, z, ra);
                    kdv += v * (m.re * m.re + m.im * m.im);
                }
            }
            S.kd[tid - 32] = kdv;
        }
        __syncthreads();
        if (tid < D) {
            double cs = 0.0;
            for (int i = 0; i < D; ++i) cs += hypot(G_re[i * OLD + tid], G_im[i * OLD + tid]);
            S.colsum[tid] = cs;
            S.gdiag[tid] = G_re[tid * OLD + tid];
        }
        __syncthreads();
        if (tid == 0) {
            double mc = 0.0;
            double poison = 0.0;     // fmax drops NaNs: carry them into nu explicitly
            for (int i = 0; i < D; ++i) { mc = fmax(mc, S.colsum[i]); poison += S.colsum[i]; }
            double nu = 2.0 * mc + 0.0 * poison;
            for (int bp = 0; bp < N; ++bp) {
                double best = 0.0;
                for (int a = 0; a < 2; ++a)
                    for (int bb = 0; bb < 2; ++bb)
                        best = fmax(best, fabs(S.jw[bp * 8 + a * 2 + bb]) + fabs(S.jw[bp * 8 + 4 + (1 - a) * 2 + (1 - bb)]));
                nu += best;
            }
            S.misc[0] = nu;
        }
        __syncthreads();
        const double nu = S.misc[0];
        double* const R = rec + (size_t)b * REC;
        if (!(nu * (P.times[T - 1] - P.times[0]) < 1e5)) {     // non-finite generator or absurd step count: reported by the evaluator
            if (tid < 4) R[NF * 32 + 72 + tid] = nan("");
            if (tid == 0 && cost != nullptr) cost[b] = 0.f;
            __syncthreads();
            continue;
        }

        // ---- record: A fragments of G for all four tile rows X, column tiles kt = 2 (X >> 1) + kk, k index permuted pi(ks, t) = 2t + ks
        //      (fragment f = (2 X + kk) 2 + ks for every lane, coalesced; im plane, then re plane for a complex H), weights, diagonal
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int f = 4 * w + q, X = f >> 2, kk = (f >> 1) & 1, ks = f & 1;
            const int row = phys32(8 * X + g), col = phys32(8 * (2 * (X >> 1) + kk) + 2 * t + ks);
            R[f * 32 + lane] = G_im[row * OLD + col];
            if (!REAL_H) R[(16 + f) * 32 + lane] = G_re[row * OLD + col];
        }
        if (tid < 40) R[NF * 32 + tid] = S.jw[tid];
        if (tid >= 64 && tid < 96) R[NF * 32 + 40 + (tid - 64)] = S.gdiag[tid - 64];

        // ---- Chebyshev enclosure (eml_cheb.cuh): G is still in the Vs[0] planes; Vs[1] and the trace table are scratch
        ChebPlan plan{};
        {
            if (tid < D) {
                double rad = 0.0;
                for (int c = 0; c < D; ++c)
                    if (c != tid) rad += REAL_H ? fabs(G_im[tid * OLD + c]) : hypot(G_re[tid * OLD + c], G_im[tid * OLD + c]);
                const double ctr = -G_im[tid * OLD + tid];
                S.colsum[tid] = ctr + rad; S.rowlo[tid] = ctr - rad; S.diagh[tid] = ctr;
            }
            __syncthreads();
            double hi = -1e300, lo = 1e300, mu = 0.0;
            for (int i = 0; i < D; ++i) { hi = fmax(hi, S.colsum[i]); lo = fmin(lo, S.rowlo[i]); mu += S.diagh[i]; }
            mu *= 1.0 / D;
            double spread = hi - lo;
            // 2 rho(H - mu) <= 2 ||(H - mu)^8||_1^(1/8): three squarings, planes [re | im] of 32 x 32
            // (real H: on the DMMA pipe, warp w computes tile row w; row stride 36 makes the fragment loads conflict free)
            constexpr int LDM = REAL_H ? 36 : D;
            double* Ma_re = S.M[0]; double* Ma_im = S.M[1];
            double* Mb_re = S.M[2];     double* Mb_im = S.M[REAL_H ? 2 : 3];
            
            for (int i = tid; i < D * D; i += 128) {
                const int r = i / D, c = i % D;
                Ma_re[r * LDM + c] = -G_im[r * OLD + c] - ((r == c) ? mu : 0.0);
                if (!REAL_H) Ma_im[i] = (r == c) ? 0.0 : G_re[r * OLD + c];
            }
            __syncthreads();
            for (int sq = 0; sq < 3; ++sq) {
                if constexpr (REAL_H) {
#pragma unroll
                    for (int J = 0; J < 4; ++J) {
                        double d0 = 0.0, d1 = 0.0;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            dmma_o(d0, d1, Ma_re[(8 * w + g) * LDM + 4 * ks + t], Ma_re[(4 * ks + t) * LDM + 8 * J + g]);
                        *reinterpret_cast<double2*>(&Mb_re[(8 * w + g) * LDM + 8 * J + 2 * t]) = make_double2(d0, d1);
                    }
                } else
                for (int i = tid; i < D * D; i += 128) {
                    const int r = i / D, c = i % D;
                    double sre = 0.0, sim = 0.0;
#pragma unroll 4
                    for (int k = 0; k < D; ++k) {
                        const double are = Ma_re[r * D + k], bre = Ma_re[k * D + c];
                        const double aim = Ma_im[r * D + k], bim = Ma_im[k * D + c];
                        sre = fma(are, bre, fma(-aim, bim, sre));
                        sim = fma(are, bim, fma(aim, bre, sim));
                    }
                    Mb_re[i] = sre;
                    Mb_im[i] = sim;
                }
                __syncthreads();
                double* tmp = Ma_re; Ma_re = Mb_re; Mb_re = tmp;
                tmp = Ma_im; Ma_im = Mb_im; Mb_im = tmp;
            }
            if (tid < D) {
                double cs = 0.0;
                for (int r = 0; r < D; ++r) cs += REAL_H ? fabs(Ma_re[r * LDM + tid]) : hypot(Ma_re[r * D + tid], Ma_im[r * D + tid]);
                S.colsum[tid] = cs;
            }
            __syncthreads();
            double cs = 0.0;
            for (int i = 0; i < D; ++i) cs = fmax(cs, S.colsum[i]);
            spread = fmin(spread, 2.0 * sqrt(sqrt(sqrt(cs))) * (1.0 + 1e-12));
            double dlo, dhi, dasym;
            cheb_site_ranges(S.jw, S.kd, N, dlo, dhi, dasym);
            plan = cheb_make_plan(spread, dlo, dhi, dasym, P.times[T - 1] - P.times[0], cheb_zcap, ORB_KCAP);
        }
        if (tid == 0) {
            R[NF * 32 + 72] = plan.R; R[NF * 32 + 73] = plan.S; R[NF * 32 + 74] = plan.ac; R[NF * 32 + 75] = plan.tau_max;
            if (cost != nullptr) {
                const double span = P.times[T - 1] - P.times[0];
                const double steps = ceil(span / plan.tau_max);
                const int k1 = min(cheb_terms(fmin(span, plan.tau_max) * plan.S), ORB_KCAP);
                cost[b] = (float)(fmax(steps, 1.0) * (double)k1);
            }
        }
    }
}
template __global__ void prepare32_kernel<true>(const DevProblem, const long long, const double*, double*, float*, const double);
template __global__ void prepare32_kernel<false>(const DevProblem, const long long, const double*, double*, float*, const double);


template <bool REAL_H>
__global__ void __launch_bounds__(128, ORB_CTAS)
eval_cta_orbit_kernel(const DevProblem P, const long long n_models, const double* __restrict__ prep,
                      const int* __restrict__ target_set, double* __restrict__ expect, double* __restrict__ loss,
                      int* __restrict__ status, unsigned int* __restrict__ counter, const int* __restrict__ order,
                      const int sms) {
    extern __shared__ __align__(16) unsigned char orbit_smem_raw[];
    OrbitScratch& S = *reinterpret_cast<OrbitScratch*>(orbit_smem_raw);
    const int tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int T = P.n_times, K = P.n_obs;

    for (;;) {
        __syncthreads();
        if (tid == 0) S.fetch = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int bidx = S.fetch;
        if ((long long)bidx >= n_models) break;
        const long long b = (order != nullptr) ? order[bidx] : bidx;

        // ---- the prepare kernel's record of this model: A fragments (coalesced), weights, diagonal, plan
        constexpr int NF = REAL_H ? 16 : 32, REC = orbit_rec_doubles<REAL_H>();
        const double* const R = prep + (size_t)b * REC;
        ChebPlan plan;
        plan.R = R[NF * 32 + 72]; plan.S = R[NF * 32 + 73]; plan.ac = R[NF * 32 + 74]; plan.tau_max = R[NF * 32 + 75];
        if (!(plan.R == plan.R)) {       // not evaluable (non-finite parameters or an absurd step count): reported, not attempted
            if (tid == 0) {
                if (loss != nullptr) loss[b] = nan("");
                if (status != nullptr) status[b] = -1;
            }
            continue;
        }
        double aG_re[4][2][2], aG_im[4][2][2];
#pragma unroll
        for (int X = 0; X < 4; ++X)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const int f = (2 * X + kk) * 2 + ks;
                    aG_im[X][kk][ks] = R[f * 32 + lane];
                    aG_re[X][kk][ks] = REAL_H ? 0.0 : R[(16 + f) * 32 + lane];
                }
        if (tid < 40) S.jw[tid] = R[NF * 32 + tid];
        if (tid >= 64 && tid < 96) S.gdiag[tid - 64] = R[NF * 32 + 40 + (tid - 64)];

        const double sc = 2.0 / plan.R;
#pragma unroll
        for (int X = 0; X < 4; ++X)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) { aG_re[X][kk][ks] *= sc; aG_im[X][kk][ks] *= sc; }
        __syncthreads();     // jw and the diagonal are in shared memory

        OrbitCtx C;
        C.P = &P; C.S = &S; C.b = b; C.expect = expect;
        const int tset = (target_set != nullptr) ? target_set[b] : 0;
        C.tgt = (P.targets != nullptr) ? P.targets + (size_t)tset * K * T * 2 : nullptr;
        C.lane = lane; C.g = g; C.t = t; C.w = w;
        C.Rb = plan.R; C.Sb = plan.S; C.ac = plan.ac; C.tau_max = plan.tau_max; C.sc = sc;
        double my_loss = 0.0;
        int n_apply = 0;
        bool capped = false;
        // CTAs b and b + (number of SMs) share an SM: rotating the roles by two puts their D warps on different schedulers
        const int role = EML_ORBIT_ROTATE ? ((w + 2 * ((int)(blockIdx.x / sms) & 1)) & 3) : w;
        if (role == 0) orbit_series<REAL_H, 0>(C, aG_re, aG_im, my_loss, n_apply, capped);
        else if (role == 1) orbit_series<REAL_H, 1>(C, aG_re, aG_im, my_loss, n_apply, capped);
        else if (role == 2) orbit_series<REAL_H, 2>(C, aG_re, aG_im, my_loss, n_apply, capped);
        else orbit_series<REAL_H, 3>(C, aG_re, aG_im, my_loss, n_apply, capped);

        // every warp produced outputs: combine the per-warp losses in a fixed order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
        const int any_capped = __syncthreads_or(capped ? 1 : 0);
        if (lane == 0) S.lossw[w] = my_loss;
        __syncthreads();
        if (tid == 0) {
            const double tot = (S.lossw[0] + S.lossw[1]) + (S.lossw[2] + S.lossw[3]);
            if (loss != nullptr) loss[b] = (C.tgt != nullptr) ? tot / ((double)T * (double)K) : 0.0;
            if (status != nullptr) status[b] = any_capped ? -1 : n_apply;
        }
    }
}

template __global__ void eval_cta_orbit_kernel<true>(const DevProblem, const long long, const double*, const int*, double*, double*, int*, unsigned int*, const int*, const int);
template __global__ void eval_cta_orbit_kernel<false>(const DevProblem, const long long, const double*, const int*, double*, double*, int*, unsigned int*, const int*, const int);

// eml_warp_mma.cu: one-CTA counting sort of per-model costs, largest first
void launch_order_from_cost(int n_models, const float* cost, int* order, cudaStream_t stream);

size_t cta_orbit_prep_bytes_per_model(bool real_h) {
    return sizeof(double) * (size_t)(real_h ? orbit_rec_doubles<true>() : orbit_rec_doubles<false>());
}

// order: ordering buffer of the plan (2 * 2^18 ints: order + float cost scratch) or nullptr; prep_ws: records for prep_cap models
cudaError_t launch_cta_orbit(const DevProblem& P, long long n_models, const double* params, const int* target_set, double* expect,
                             double* loss, int* status, unsigned int* counter, int* order, bool real_h, double* prep_ws,
                             long long prep_cap, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if (prep_ws == nullptr || prep_cap < 1) return cudaErrorInvalidValue;
    using Kern = void (*)(const DevProblem, const long long, const double*, const int*, double*, double*, int*, unsigned int*,
                          const int*, const int);
    using PrepKern = void (*)(const DevProblem, const long long, const double*, double*, float*, const double);
    const Kern kern = real_h ? (Kern)eval_cta_orbit_kernel<true> : (Kern)eval_cta_orbit_kernel<false>;
    const PrepKern prep = real_h ? (PrepKern)prepare32_kernel<true> : (PrepKern)prepare32_kernel<false>;
    const size_t smem = sizeof(OrbitScratch);
    const size_t psmem = real_h ? sizeof(Prep32Scratch<true>) : sizeof(Prep32Scratch<false>);
    const size_t rec = real_h ? orbit_rec_doubles<true>() : orbit_rec_doubles<false>();
    static int per_sm_cache[64][2] = {{0}}, prep_per_sm_cache[64][2] = {{0}};
    int per_sm = (dev < 64) ? per_sm_cache[dev][real_h ? 1 : 0] : 0;
    int prep_per_sm = (dev < 64) ? prep_per_sm_cache[dev][real_h ? 1 : 0] : 0;
    if (per_sm == 0) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&prep_per_sm, prep, 128, psmem)) != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        if (prep_per_sm < 1) prep_per_sm = 1;
        if (dev < 64) { per_sm_cache[dev][real_h ? 1 : 0] = per_sm; prep_per_sm_cache[dev][real_h ? 1 : 0] = prep_per_sm; }
    }
    const double zcap = cheb_zcap_for(ORB_KCAP);
    for (long long off = 0; off < n_models; off += prep_cap) {
        const long long n = (n_models - off < prep_cap) ? n_models - off : prep_cap;
        if ((e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream)) != cudaSuccess) return e;
        // longest first, by the generator applications the prepare kernel computes, when CTAs take more than one model
        const bool sort = order != nullptr && n > (long long)per_sm * sms && n <= (1 << 18);
        float* const cost = sort ? reinterpret_cast<float*>(order + (1 << 18)) : nullptr;
        long long pctas = n;
        if (pctas > (long long)prep_per_sm * sms) pctas = (long long)prep_per_sm * sms;
        prep<<<(unsigned)pctas, 128, psmem, stream>>>(P, n, params + off * P.n_params, prep_ws, cost, zcap);
        if (sort) launch_order_from_cost((int)n, cost, order, stream);
        long long ctas = n;
        if (ctas > (long long)per_sm * sms) ctas = (long long)per_sm * sms;
        kern<<<(unsigned)ctas, 128, smem, stream>>>(P, n, prep_ws, target_set ? target_set + off : nullptr,
                                                   expect ? expect + (size_t)off * P.n_obs * P.n_times * 2 : nullptr, loss ? loss + off : nullptr,
                                                   status ? status + off : nullptr, counter, sort ? order : nullptr, sms);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace eml
