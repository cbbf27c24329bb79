// d = 32 (n_tls = 5) evaluator: ONE CTA OF FOUR WARPS PER MODEL, FP64 DMMA.
//
// Same algorithms and register layout as eml_warp_mma.cu (Chebyshev / Bessel series by default, Taylor series with
// dense output as the cross-check, both acting on the d x d density matrix; Hermitian B-operand trick, real-weighted
// site stencils, accumulators started from J(v)/2; parity-blocked G when H conserves the index parity), scaled up:
// warp w owns the tile row I = w (rows 8w .. 8w+7) of rho, of the current series term v and of X = G v, each as
// 4 column tiles x 2 elements per lane -- the same 16 complex numbers per lane as the d = 16 kernel.  What a warp
// needs from the other rows goes through shared memory, twice per generator application:
//   Vs  (row-major, ld 40)  the whole current term: B fragments v[k][n] = conj(v[n][k]) and all five site stencils
//                           are read from it with conflict-free LDS.128;
//   Xs  (column-major, ld 34)  X' = J(v)/2 + G v, so that X'^dag is read back with LDS.128.
// Two __syncthreads per application.  REAL_H as in the warp kernel (half the DMMAs when H is a real matrix).
//
// Replaces qutip.mesolve + e_ops + loss for one model (reference basic_model.py:381-388, learning_chain.py:738-742).
#include <cstdlib>
#include <type_traits>
#include "eml_common.cuh"
#include "eml_cheb.cuh"

namespace eml {

constexpr int CQ = 16;        // outputs per expansion
constexpr int CMAXP = 192;
constexpr int LDV = 40;       // row stride (doubles) of Vs: conflict-free LDS.128 over (g, t)
constexpr int LDX = 34;       // column stride of Xs: conflict-free STS.64 over (g, t)

constexpr int CTA_CHEB_KCAP = 1024;   // Chebyshev terms per macro step that fit the trace table (32 KB)

template <bool CHEB>
struct alignas(16) CtaScratch {
    static constexpr int KCAP = CHEB ? CTA_CHEB_KCAP : 0;
    double Vs_re[32 * LDV], Vs_im[32 * LDV];     // also hosts G during assembly: G_re = Vs_re (ld 40), G_im = Vs_im
    double Xs_re[32 * LDX], Xs_im[32 * LDX];
    double par[CMAXP];
    double jw[40];            // [bitpos 0..4][type (0 diag, 1 flip)][ra][ca]
    double kd[10];            // [bitpos][ra]: diagonal of sum rate * A^dag A per site
    double colsum[32], rowlo[32], diagh[32];
    double trp[4][4];         // per-warp partial traces: b0.re, b0.im, b1.re, b1.im
    double misc[4];
    double lossw[4];
    unsigned int mxs[4];
    unsigned int fetch;
    unsigned int pad_;
    alignas(16) double tr[(KCAP + 1) * 4];       // partial traces of the Chebyshev terms (also scratch for the spread bound)
    double cJ[KCAP + 1];                          // Bessel coefficients of the step end
};

#define CUNI(c) (__all_sync(0xffffffffu, (c)) != 0)

__constant__ double CTA_INV_TABLE[MCAP + 2];     // 1/(j+1)

__device__ __forceinline__ void dmma_c(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ double neg_c(double v) { return __hiloint2double(__double2hiint(v) ^ 0x80000000, __double2loint(v)); }

template <bool REAL_H, bool CHEB, bool PARITY>
__global__ void __launch_bounds__(128, 2)
eval_cta_mma_kernel(const DevProblem P, const long long n_models, const double* __restrict__ params,
                    const int* __restrict__ target_set, double* __restrict__ expect, double* __restrict__ loss,
                    int* __restrict__ status, unsigned int* __restrict__ counter, const int* __restrict__ order,
                    const double cheb_zcap) {
    constexpr int N = 5, D = 32;
    static_assert(!PARITY || CHEB, "the parity-blocked variant exists for the Chebyshev kernel");
    // PARITY (every Hamiltonian term conserves the parity of the basis index, see eml_warp_mma.cu): the TOP bit of the
    // kernel's row / column index is the parity of the physical index, x' = x ^ (parity(x & 15) << 4).  G then has
    // non-zero tiles only between tile rows / columns with the same top bit: half the DMMAs and B-fragment loads; a
    // flip of any site below the observed one also flips the top tile bit of row and column.
    auto phys = [](const int x) { return PARITY ? x ^ (((x ^ (x >> 1) ^ (x >> 2) ^ (x >> 3)) & 1) << 4) : x; };
    extern __shared__ __align__(16) unsigned char cta_smem_raw[];
    using Scratch = CtaScratch<CHEB>;
    Scratch& S = *reinterpret_cast<Scratch*>(cta_smem_raw);
    const int tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int T = P.n_times, K = P.n_obs;
    const int obs_site = P.obs_site;
    auto pos = [&](int s) { return (s == obs_site) ? N - 1 : (s < obs_site ? N - 2 - s : N - 1 - s); };
    auto orig_index = [&](int x) {
        int o = 0;
#pragma unroll
        for (int s = 0; s < N; ++s) o |= ((x >> pos(s)) & 1) << (N - 1 - s);
        return o;
    };
    const int gp = g & 1;
    const bool diag_lane = (t == (g >> 1));
    const int dl4 = 4 * (g ^ 4) + ((g ^ 4) >> 1), dl2 = 4 * (g ^ 2) + ((g ^ 2) >> 1), dl1 = 4 * (g ^ 1) + ((g ^ 1) >> 1);

    for (;;) {
        __syncthreads();
        if (tid == 0) S.fetch = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int bidx = S.fetch;
        if ((long long)bidx >= n_models) break;
        const long long b = (order != nullptr) ? order[bidx] : bidx;

        // ---- parameters, clear G (G lives in the Vs planes with ld LDV during assembly)
        for (int i = tid; i < P.n_params; i += 128) S.par[i] = params[b * P.n_params + i];
        for (int i = tid; i < 32 * LDV; i += 128) { S.Vs_re[i] = 0.0; S.Vs_im[i] = 0.0; }
        __syncthreads();
        double* G_re = S.Vs_re;
        double* G_im = S.Vs_im;

        // ---- K1: thread r < 32 owns row r of G; threads < 40 own one jump weight each
        if (tid < D) {
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
                        G_re[r * LDV + c] += v * A.im;
                        G_im[r * LDV + c] -= v * A.re;
                    }
                } else if (tm.kind == EML_TERM_COUPLING) {
                    const int sb = pos(tm.site_b);
                    const int rb = (r >> sb) & 1;
                    const double v2 = v * v;
                    for (int a = 0; a < 2; ++a)
                        for (int bb = 0; bb < 2; ++bb) {
                            const cplx AB = cmul(ld_op(P.op_table, tm.op_a, ra, a), ld_op(P.op_table, tm.op_b, rb, bb));
                            const int c = (r & ~((1 << sa) | (1 << sb))) | (a << sa) | (bb << sb);
                            G_re[r * LDV + c] += v2 * AB.im;
                            G_im[r * LDV + c] -= v2 * AB.re;
                        }
                } else {
                    for (int a = 0; a < 2; ++a) {
                        cplx M = {0.0, 0.0};
                        for (int z = 0; z < 2; ++z) {
                            const cplx m = cmul(cconj(ld_op(P.op_table, tm.op_a, z, ra)), ld_op(P.op_table, tm.op_a, z, a));
                            M.re += m.re; M.im += m.im;
                        }
                        const int c = (r & ~(1 << sa)) | (a << sa);
                        G_re[r * LDV + c] -= 0.5 * v * M.re;
                        G_im[r * LDV + c] -= 0.5 * v * M.im;
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
        if (CHEB && tid >= 32 && tid < 42) {
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
            for (int i = 0; i < D; ++i) cs += hypot(G_re[i * LDV + tid], G_im[i * LDV + tid]);
            S.colsum[tid] = cs;
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
        if (!(nu * (P.times[T - 1] - P.times[0]) < 1e5)) {     // non-finite generator or absurd step count: do not hang
            if (tid == 0) {
                if (loss != nullptr) loss[b] = nan("");
                if (status != nullptr) status[b] = -1;
            }
            continue;
        }

        // ---- A fragments of this warp's rows (k index permuted: pi(ks,t) = 8(ks/2) + 2t + ks%2), stencil weights
        double a_re[8], a_im[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            // PARITY: only the column tiles kt = 2 (w >> 1) + {0, 1} are non-zero; they are kept in ks = 0..3
            const int kt = PARITY ? 2 * (w >> 1) + ((ks >> 1) & 1) : (ks >> 1);
            const int k = 8 * kt + 2 * t + (ks & 1);
            a_re[ks] = REAL_H ? 0.0 : G_re[phys(8 * w + g) * LDV + phys(k)];
            a_im[ks] = G_im[phys(8 * w + g) * LDV + phys(k)];
        }
        double W0[4][2];     // halved diagonal weights for (J, e)
#pragma unroll
        for (int J = 0; J < 4; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = phys(8 * w + g), col = phys(8 * J + 2 * t + e);      // physical indices
                double x = S.jw[4 * 8 + (row >> 4) * 2 + (col >> 4)] + S.jw[3 * 8 + (w & 1) * 2 + (J & 1)] +
                           S.jw[2 * 8 + (g >> 2) * 2 + (t >> 1)] + S.jw[1 * 8 + ((g >> 1) & 1) * 2 + (t & 1)] +
                           S.jw[0 * 8 + (g & 1) * 2 + e];
                if (REAL_H) x += G_re[row * LDV + row] + G_re[col * LDV + col];
                W0[J][e] = 0.5 * x;
            }
        // flip weight of the observed site: by J >> 1; PARITY: by (J, e), the observed bits being (row >> 4, col >> 4)
        constexpr int NF4 = PARITY ? 8 : 2;
        double F4[NF4];
        if constexpr (PARITY) {
#pragma unroll
            for (int J = 0; J < 4; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    F4[2 * J + e] = 0.5 * S.jw[4 * 8 + 4 + (phys(8 * w + g) >> 4) * 2 + (phys(8 * J + 2 * t + e) >> 4)];
        } else {
            F4[0] = 0.5 * S.jw[4 * 8 + 4 + (w >> 1) * 2 + 0]; F4[1] = 0.5 * S.jw[4 * 8 + 4 + (w >> 1) * 2 + 1];
        }
        double F3[2] = {0.5 * S.jw[3 * 8 + 4 + (w & 1) * 2 + 0], 0.5 * S.jw[3 * 8 + 4 + (w & 1) * 2 + 1]};     // by J&1
        double f2 = 0.5 * S.jw[2 * 8 + 4 + (g >> 2) * 2 + (t >> 1)];
        double f1 = 0.5 * S.jw[1 * 8 + 4 + ((g >> 1) & 1) * 2 + (t & 1)];
        double f0[2] = {0.5 * S.jw[0 * 8 + 4 + (g & 1) * 2 + 0], 0.5 * S.jw[0 * 8 + 4 + (g & 1) * 2 + 1]};
        // the generator is kept pre-multiplied by the current step length (see eml_warp_mma.cu): u_{j+1} = (h L) u_j
        double h_cur = 1.0;
        // ---- Chebyshev enclosure (see eml_cheb.cuh): G is still in the Vs planes, Xs and the trace table are free
        [[maybe_unused]] ChebPlan plan{};
        if constexpr (CHEB) {
            if (tid < D) {
                double rad = 0.0;
                for (int c = 0; c < D; ++c)
                    if (c != tid) rad += REAL_H ? fabs(G_im[tid * LDV + c]) : hypot(G_re[tid * LDV + c], G_im[tid * LDV + c]);
                const double ctr = -G_im[tid * LDV + tid];
                S.colsum[tid] = ctr + rad; S.rowlo[tid] = ctr - rad; S.diagh[tid] = ctr;
            }
            __syncthreads();
            double hi = -1e300, lo = 1e300, mu = 0.0;
            for (int i = 0; i < D; ++i) { hi = fmax(hi, S.colsum[i]); lo = fmin(lo, S.rowlo[i]); mu += S.diagh[i]; }
            mu *= 1.0 / D;
            double spread = hi - lo;
            // 2 rho(H - mu) <= 2 ||(H - mu)^8||_1^(1/8): three squarings, planes [re | im] of 32 x 32
            double* Ma_re = S.Xs_re; double* Ma_im = S.Xs_im;
            double* Mb_re = S.tr;    double* Mb_im = S.tr + D * D;
            for (int i = tid; i < D * D; i += 128) {
                const int r = i / D, c = i % D;
                Ma_re[i] = -G_im[r * LDV + c] - ((r == c) ? mu : 0.0);
                if (!REAL_H) Ma_im[i] = (r == c) ? 0.0 : G_re[r * LDV + c];
            }
            __syncthreads();
            for (int sq = 0; sq < 3; ++sq) {
                for (int i = tid; i < D * D; i += 128) {
                    const int r = i / D, c = i % D;
                    double sre = 0.0, sim = 0.0;
#pragma unroll 4
                    for (int k = 0; k < D; ++k) {
                        const double are = Ma_re[r * D + k], bre = Ma_re[k * D + c];
                        sre = fma(are, bre, sre);
                        if (!REAL_H) {
                            const double aim = Ma_im[r * D + k], bim = Ma_im[k * D + c];
                            sre = fma(-aim, bim, sre);
                            sim = fma(are, bim, fma(aim, bre, sim));
                        }
                    }
                    Mb_re[i] = sre;
                    if (!REAL_H) Mb_im[i] = sim;
                }
                __syncthreads();
                double* tmp = Ma_re; Ma_re = Mb_re; Mb_re = tmp;
                tmp = Ma_im; Ma_im = Mb_im; Mb_im = tmp;
            }
            if (tid < D) {
                double cs = 0.0;
                for (int r = 0; r < D; ++r) cs += REAL_H ? fabs(Ma_re[r * D + tid]) : hypot(Ma_re[r * D + tid], Ma_im[r * D + tid]);
                S.colsum[tid] = cs;
            }
            __syncthreads();
            double cs = 0.0;
            for (int i = 0; i < D; ++i) cs = fmax(cs, S.colsum[i]);
            spread = fmin(spread, 2.0 * sqrt(sqrt(sqrt(cs))) * (1.0 + 1e-12));
            double dlo, dhi, dasym;
            cheb_site_ranges(S.jw, S.kd, N, dlo, dhi, dasym);
            plan = cheb_make_plan(spread, dlo, dhi, dasym, P.times[T - 1] - P.times[0], cheb_zcap, Scratch::KCAP);
            // generator -> 2 (L + ac) / R  (stencil weights are stored halved)
            const double sc = 2.0 / plan.R;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) { a_re[ks] *= sc; a_im[ks] *= sc; }
#pragma unroll
            for (int J = 0; J < 4; ++J) { W0[J][0] = (W0[J][0] + 0.5 * plan.ac) * sc; W0[J][1] = (W0[J][1] + 0.5 * plan.ac) * sc; }
#pragma unroll
            for (int i = 0; i < NF4; ++i) F4[i] *= sc;
            F3[0] *= sc; F3[1] *= sc;
            f2 *= sc; f1 *= sc; f0[0] *= sc; f0[1] *= sc;
        }
        __syncthreads();     // G has been consumed: the Vs planes are free for the series terms

        // ---- state: this warp's rows of rho(t_k)
        double acc_re[4][2], acc_im[4][2];
#pragma unroll
        for (int J = 0; J < 4; ++J)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = orig_index(phys(8 * w + g)), c = orig_index(phys(8 * J + 2 * t + e));
                acc_re[J][e] = P.rho0[2 * (r * D + c)];
                acc_im[J][e] = P.rho0[2 * (r * D + c) + 1];
            }
        const int tset = (target_set != nullptr) ? target_set[b] : 0;
        const double* tgt = (P.targets != nullptr) ? P.targets + (size_t)tset * K * T * 2 : nullptr;
        double my_loss = 0.0;
        int n_apply = 0;
        bool capped = false;

        // per-warp partial trace of a term into S.trp[w]: diag lanes, tiles J = (w&1) + 2b
        auto partial_trace_store = [&](const double (&xr)[4][2], const double (&xi)[4][2]) {
            // tiles J = (w&1) and (w&1)+2, selected with predicates (a runtime index would put the arrays in local memory)
            const bool odd = w & 1;
            const double ar0 = odd ? xr[1][0] : xr[0][0], ar1 = odd ? xr[1][1] : xr[0][1];
            const double ai0 = odd ? xi[1][0] : xi[0][0], ai1 = odd ? xi[1][1] : xi[0][1];
            const double br0 = odd ? xr[3][0] : xr[2][0], br1 = odd ? xr[3][1] : xr[2][1];
            const double bi0 = odd ? xi[3][0] : xi[2][0], bi1 = odd ? xi[3][1] : xi[2][1];
            // only the 8 lanes 4g + (g >> 1) hold partial-trace elements: 3-stage reduce-scatter among exactly those
            // lanes (4 double shuffles); value 2*bit2(g) + bit1(g) ends up in the lanes with g even
            // this lane's contribution to (rho00, rho11, Re rho01, Im rho01) of the observed site: tile a has column bit
            // 0, tile b column bit 1, the row bit is w >> 1; PARITY: both bits are XORed with the parity p of the low bits
            const double are_ = gp ? ar1 : ar0, aim_ = gp ? ai1 : ai0, bre_ = gp ? br1 : br0, bim_ = gp ? bi1 : bi0;
            const bool top = w & 2;
            const bool pl = PARITY ? (((g ^ (g >> 1) ^ (g >> 2) ^ w) & 1) != 0) : false;
            double v0, v1, v2, v3;
            if (!top) { v0 = pl ? 0.0 : are_; v1 = pl ? are_ : 0.0; v2 = pl ? 0.0 : bre_; v3 = pl ? 0.0 : bim_; }
            else      { v0 = pl ? bre_ : 0.0; v1 = pl ? 0.0 : bre_; v2 = pl ? are_ : 0.0; v3 = pl ? aim_ : 0.0; }
            const bool hb = g & 4, mb = g & 2;
            const double r0 = __shfl_sync(0xffffffffu, hb ? v0 : v2, dl4), r1 = __shfl_sync(0xffffffffu, hb ? v1 : v3, dl4);
            const double k0 = (hb ? v2 : v0) + r0, k1 = (hb ? v3 : v1) + r1;
            const double kk = (mb ? k1 : k0) + __shfl_sync(0xffffffffu, mb ? k0 : k1, dl2);
            const double tot = kk + __shfl_sync(0xffffffffu, kk, dl1);
            if (diag_lane && !(g & 1)) S.trp[w][g >> 1] = tot;
        };
        // after a __syncthreads: the total of value index vi = lane >> 3 (rho00.re, rho11.re, rho01.re, rho01.im)
        auto trace_total = [&]() -> double {
            const int vi = lane >> 3;
            // rho_red[a][b]: a = w>>1 of the contributing warps, b = tile half
            return (S.trp[0][vi] + S.trp[1][vi]) + (S.trp[2][vi] + S.trp[3][vi]);
        };
        auto emit = [&](int q, int kbase, double y0, double y1) {
            const int p = lane;
            const int srcl = p & 7;
            double r[4];
#pragma unroll
            for (int vi = 0; vi < 4; ++vi) {
                const double lo = __shfl_sync(0xffffffffu, y0, vi * 8 + srcl), hi = __shfl_sync(0xffffffffu, y1, vi * 8 + srcl);
                r[vi] = (p & 8) ? hi : lo;
            }
            if (w == 0 && p < q) {
                const int k = kbase + p;
                for (int m = 0; m < K; ++m) {
                    const double* O = P.obs + m * 8;
                    const double ere = O[0] * r[0] + O[6] * r[1] + (O[2] + O[4]) * r[2] + (O[3] - O[5]) * r[3];
                    const double eim = O[1] * r[0] + O[7] * r[1] + (O[3] + O[5]) * r[2] + (O[4] - O[2]) * r[3];
                    if (expect != nullptr) {
                        double* e = expect + (((size_t)b * K + m) * T + k) * 2;
                        e[0] = ere; e[1] = eim;
                    }
                    if (tgt != nullptr) {
                        const double dre = ere - tgt[((size_t)m * T + k) * 2], dim = eim - tgt[((size_t)m * T + k) * 2 + 1];
                        my_loss += dre * dre + dim * dim;
                    }
                }
            }
        };

        // ---- generator application core, shared by both series.  Given x = (diagonal weight) * v (+ anything Hermitian / 2)
        //      and v published in Vs: add the site stencils (J(v)/2) and G v on the DMMA pipe, publish X' in Xs.
        auto apply_core = [&](const double (&v_re)[4][2], const double (&v_im)[4][2], double (&x_re)[4][2], double (&x_im)[4][2]) {
#pragma unroll
            for (int J = 0; J < 4; ++J) {
                // bit 4: (I^2, J^2); bit 3: (I^1, J^1); bit 2: (g^4, t^2); bit 1: (g^2, t^1); bit 0: (g^1, e^1)
                // PARITY: a flip of a site below the observed one also flips the top (parity) tile bit of row and column
                constexpr int PX = PARITY ? 2 : 0;
                const int o4 = (8 * (w ^ 2) + g) * LDV + 8 * (J ^ 2) + 2 * t;
                const int o3 = (8 * (w ^ 1 ^ PX) + g) * LDV + 8 * (J ^ 1 ^ PX) + 2 * t;
                const int o2 = (8 * (w ^ PX) + (g ^ 4)) * LDV + 8 * (J ^ PX) + 2 * (t ^ 2);
                const int o1 = (8 * (w ^ PX) + (g ^ 2)) * LDV + 8 * (J ^ PX) + 2 * (t ^ 1);
                const int o0 = (8 * (w ^ PX) + (g ^ 1)) * LDV + 8 * (J ^ PX) + 2 * t;
                const double2 r4 = *reinterpret_cast<const double2*>(&S.Vs_re[o4]), i4 = *reinterpret_cast<const double2*>(&S.Vs_im[o4]);
                const double2 r3 = *reinterpret_cast<const double2*>(&S.Vs_re[o3]), i3 = *reinterpret_cast<const double2*>(&S.Vs_im[o3]);
                const double2 r2 = *reinterpret_cast<const double2*>(&S.Vs_re[o2]), i2 = *reinterpret_cast<const double2*>(&S.Vs_im[o2]);
                const double2 r1 = *reinterpret_cast<const double2*>(&S.Vs_re[o1]), i1 = *reinterpret_cast<const double2*>(&S.Vs_im[o1]);
                const double2 r0 = *reinterpret_cast<const double2*>(&S.Vs_re[o0]), i0 = *reinterpret_cast<const double2*>(&S.Vs_im[o0]);
                const double w40 = PARITY ? F4[(2 * J) & (NF4 - 1)] : F4[J >> 1], w41 = PARITY ? F4[(2 * J + 1) & (NF4 - 1)] : F4[J >> 1];
                const double w3 = F3[J & 1];
                x_re[J][0] = fma(w40, r4.x, x_re[J][0]); x_re[J][1] = fma(w41, r4.y, x_re[J][1]);
                x_im[J][0] = fma(w40, i4.x, x_im[J][0]); x_im[J][1] = fma(w41, i4.y, x_im[J][1]);
                x_re[J][0] = fma(w3, r3.x, x_re[J][0]); x_re[J][1] = fma(w3, r3.y, x_re[J][1]);
                x_im[J][0] = fma(w3, i3.x, x_im[J][0]); x_im[J][1] = fma(w3, i3.y, x_im[J][1]);
                x_re[J][0] = fma(f2, r2.x, x_re[J][0]); x_re[J][1] = fma(f2, r2.y, x_re[J][1]);
                x_im[J][0] = fma(f2, i2.x, x_im[J][0]); x_im[J][1] = fma(f2, i2.y, x_im[J][1]);
                x_re[J][0] = fma(f1, r1.x, x_re[J][0]); x_re[J][1] = fma(f1, r1.y, x_re[J][1]);
                x_im[J][0] = fma(f1, i1.x, x_im[J][0]); x_im[J][1] = fma(f1, i1.y, x_im[J][1]);
                x_re[J][0] = fma(f0[0], r0.y, x_re[J][0]); x_re[J][1] = fma(f0[1], r0.x, x_re[J][1]);   // e^1
                x_im[J][0] = fma(f0[0], i0.y, x_im[J][0]); x_im[J][1] = fma(f0[1], i0.x, x_im[J][1]);
            }
            // B fragments: v[k][n] = conj(v[n][k]) with n = 8J+g (a row of warp J), k = pi(ks,t): read pairs
            // (ks even, ks odd) = columns (2t, 2t+1) of column tile ks/2 with one LDS.128
#pragma unroll
            for (int J = 0; J < 4; ++J)
#pragma unroll
                for (int kk = 0; kk < (PARITY ? 2 : 4); ++kk) {
                    const int kt = PARITY ? 2 * (w >> 1) + kk : kk;      // column tile of G = row tile of the B operand
                    const int o = (8 * J + g) * LDV + 8 * kt + 2 * t;
                    const double2 br = *reinterpret_cast<const double2*>(&S.Vs_re[o]);
                    const double2 bi = *reinterpret_cast<const double2*>(&S.Vs_im[o]);
                    // X.re += a.re*b_re + a.im*b_im ; X.im += -a.re*b_im + a.im*b_re   (B = conj(b))
                    if (!REAL_H) {
                        dmma_c(x_re[J][0], x_re[J][1], a_re[2 * kk], br.x);
                        dmma_c(x_im[J][0], x_im[J][1], a_re[2 * kk], neg_c(bi.x));
                        dmma_c(x_re[J][0], x_re[J][1], a_re[2 * kk + 1], br.y);
                        dmma_c(x_im[J][0], x_im[J][1], a_re[2 * kk + 1], neg_c(bi.y));
                    }
                    dmma_c(x_re[J][0], x_re[J][1], a_im[2 * kk], bi.x);
                    dmma_c(x_im[J][0], x_im[J][1], a_im[2 * kk], br.x);
                    dmma_c(x_re[J][0], x_re[J][1], a_im[2 * kk + 1], bi.y);
                    dmma_c(x_im[J][0], x_im[J][1], a_im[2 * kk + 1], br.y);
                }
            // publish X' column-major so that the transpose is a vector load
#pragma unroll
            for (int J = 0; J < 4; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    S.Xs_re[(8 * J + 2 * t + e) * LDX + 8 * w + g] = x_re[J][e];
                    S.Xs_im[(8 * J + 2 * t + e) * LDX + 8 * w + g] = x_im[J][e];
                }
        };
        // after a __syncthreads: w = X' + X'^dag,  X'^dag[8w+g][8J+2t+e] = conj(X'[8J+2t+e][8w+g])
        auto transpose_add = [&](const double (&x_re)[4][2], const double (&x_im)[4][2], double (&w_re)[4][2], double (&w_im)[4][2]) {
#pragma unroll
            for (int J = 0; J < 4; ++J) {
                const int o = (8 * w + g) * LDX + 8 * J + 2 * t;
                const double2 tr_ = *reinterpret_cast<const double2*>(&S.Xs_re[o]);
                const double2 ti_ = *reinterpret_cast<const double2*>(&S.Xs_im[o]);
                w_re[J][0] = x_re[J][0] + tr_.x; w_re[J][1] = x_re[J][1] + tr_.y;
                w_im[J][0] = x_im[J][0] - ti_.x; w_im[J][1] = x_im[J][1] - ti_.y;
            }
        };
        // observables and loss of one output time from the reduced density matrix r = (rho00, rho11, Re rho01, Im rho01)
        auto finish_output = [&](const int k, const double (&r)[4]) {
            for (int m = 0; m < K; ++m) {
                const double* O = P.obs + m * 8;
                const double ere = O[0] * r[0] + O[6] * r[1] + (O[2] + O[4]) * r[2] + (O[3] - O[5]) * r[3];
                const double eim = O[1] * r[0] + O[7] * r[1] + (O[3] + O[5]) * r[2] + (O[4] - O[2]) * r[3];
                if (expect != nullptr) {
                    double* e = expect + (((size_t)b * K + m) * T + k) * 2;
                    e[0] = ere; e[1] = eim;
                }
                if (tgt != nullptr) {
                    const double dre = ere - tgt[((size_t)m * T + k) * 2], dim = eim - tgt[((size_t)m * T + k) * 2 + 1];
                    my_loss += dre * dre + dim * dim;
                }
            }
        };

        partial_trace_store(acc_re, acc_im);
        __syncthreads();
        emit(1, 0, trace_total(), 0.0);

        if constexpr (CHEB) {
        const double Rb = plan.R, Sb = plan.S, ac = plan.ac, tau_max = plan.tau_max;
        int kidx = 0;
        while (kidx < T - 1) {           // every scalar below is computed identically by all threads
            const double t0 = P.times[kidx];
            int q = 0;      // output times within tau_max of t0: 32 candidates per ballot (every warp computes the same)
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
                __syncthreads();
                partial_trace_store(acc_re, acc_im);
                __syncthreads();
                double r[4];
#pragma unroll
                for (int vi = 0; vi < 4; ++vi) r[vi] = (S.trp[0][vi] + S.trp[1][vi]) + (S.trp[2][vi] + S.trp[3][vi]);
                for (int p = tid; p < q; p += 128) finish_output(kidx + 1 + p, r);
                kidx += q;
                continue;
            }
            int Kc = cheb_terms(tau_end * Sb);
            if (Kc > Scratch::KCAP) Kc = Scratch::KCAP;
            for (int sub = 0; sub < nsub; ++sub) {
                const bool last = (sub == nsub - 1);
                const bool need_acc = !(final_step && last);
                double cscale = 0.0;
                __syncthreads();     // the previous step's readers of tr / cJ / trp are done
                if (need_acc) {
                    const double tz = 2.0 / (tau_end * Rb);
                    double jn = 1.0, jn1 = 0.0, ssum = 0.0;
                    for (int k = Kc; k >= 1; --k) {
                        if ((k & 127) == tid) S.cJ[k] = jn;
                        if (!(k & 1)) ssum += jn;
                        const double jm = fma((double)k * tz, jn, -jn1);
                        jn1 = jn; jn = jm;
                    }
                    if (tid == 0) S.cJ[0] = jn;
                    cscale = exp(-ac * tau_end) / (2.0 * ssum + jn);
                }
                double v_re[4][2], v_im[4][2], ph_re[4][2], ph_im[4][2];
#pragma unroll
                for (int J = 0; J < 4; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        v_re[J][e] = acc_re[J][e]; v_im[J][e] = acc_im[J][e];
                        ph_re[J][e] = 0.0; ph_im[J][e] = 0.0;
                        acc_re[J][e] = 0.0; acc_im[J][e] = 0.0;
                    }
                auto run_terms = [&](auto na_tag) {
                    constexpr bool NA = decltype(na_tag)::value;     // accumulate the state at the step end
                    int hsub = 0;
                    for (int k = 0; k <= Kc; ++k) {
                        // ---- [A] publish term k: rows of this warp into Vs, per-warp partial trace
                        if (k < Kc) {
#pragma unroll
                            for (int J = 0; J < 4; ++J) {
                                *reinterpret_cast<double2*>(&S.Vs_re[(8 * w + g) * LDV + 8 * J + 2 * t]) = make_double2(v_re[J][0], v_re[J][1]);
                                *reinterpret_cast<double2*>(&S.Vs_im[(8 * w + g) * LDV + 8 * J + 2 * t]) = make_double2(v_im[J][0], v_im[J][1]);
                            }
                        }
                        partial_trace_store(v_re, v_im);
                        __syncthreads();
                        if (tid < 4)     // rho00, rho11, Re rho01, Im rho01
                            S.tr[k * 4 + tid] = (S.trp[0][tid] + S.trp[1][tid]) + (S.trp[2][tid] + S.trp[3][tid]);
                        if constexpr (NA) {
                            const double c = S.cJ[k] * cscale;
#pragma unroll
                            for (int J = 0; J < 4; ++J)
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    acc_re[J][e] = fma(c, v_re[J][e], acc_re[J][e]);
                                    acc_im[J][e] = fma(c, v_im[J][e], acc_im[J][e]);
                                }
                        }
                        if (k == Kc) break;
                        // ---- [B] accumulators = W0 v + (previous term) / 2, site stencils, G v; X' published in Xs
                        double x_re[4][2], x_im[4][2];
#pragma unroll
                        for (int J = 0; J < 4; ++J)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                x_re[J][e] = fma(W0[J][e], v_re[J][e], ph_re[J][e]);
                                x_im[J][e] = fma(W0[J][e], v_im[J][e], ph_im[J][e]);
                            }
                        apply_core(v_re, v_im, x_re, x_im);
                        __syncthreads();
                        // ---- [C] next term = X' + X'^dag
#pragma unroll
                        for (int J = 0; J < 4; ++J)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                ph_re[J][e] = scale_pow2(v_re[J][e], hsub); ph_im[J][e] = scale_pow2(v_im[J][e], hsub);
                            }
                        transpose_add(x_re, x_im, v_re, v_im);
                        hsub = 0x00100000;
                    }
                };
                if (need_acc) run_terms(std::true_type{}); else run_terms(std::false_type{});
                {   // a non-finite or absurd last term means the enclosure failed
                    const double chk = fabs(v_re[0][0]) + fabs(v_im[0][0]) + fabs(v_re[3][1]);
                    if (!(chk < 1e300)) capped = true;
                }
                n_apply += Kc;
                __syncthreads();     // the trace table is complete
                if (last) {
                    // ---- outputs of the step: warp w takes the blocks of 32 outputs w, w + 4, ...
                    for (int pbase = 32 * w; pbase < q; pbase += 128) {
                        const int p = pbase + lane;
                        const bool act = p < q;
                        const double tau = act ? ((nsub == 1) ? P.times[kidx + 1 + p] - t0 : tau_end) : 0.0;
                        const double z = tau * Rb;
                        const bool direct = !(z >= 1e-40);
                        int Kp = 0;
                        if (act && !direct) { Kp = cheb_terms(tau * Sb); if (Kp > Kc) Kp = Kc; }
                        const int kmax = __reduce_max_sync(0xffffffffu, Kp);
                        const double tz = direct ? 0.0 : 2.0 / z;
                        double a0 = 0.0, a2 = 0.0, a3 = 0.0, ssum = 0.0, jn = 0.0, jn1 = 0.0;
                        double kd = (double)kmax;
#pragma unroll 2
                        for (int k = kmax; k >= 1; --k, kd -= 1.0) {
                            if (k == Kp) jn = 1.0;
                            const double2 ta = *reinterpret_cast<const double2*>(&S.tr[k * 4]);
                            const double2 tb = *reinterpret_cast<const double2*>(&S.tr[k * 4 + 2]);
                            a0 = fma(jn, ta.x, a0); a2 = fma(jn, tb.x, a2); a3 = fma(jn, tb.y, a3);
                            if (!(k & 1)) ssum += jn;
                            const double jm = fma(kd * tz, jn, -jn1);
                            jn1 = jn; jn = jm;
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
                        if (act) finish_output(kidx + 1 + p, r);
                    }
                }
            }
            kidx += q;
        }
        // every warp produced outputs: combine the per-warp losses in a fixed order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
        __syncthreads();
        if (lane == 0) S.lossw[w] = my_loss;
        __syncthreads();
        my_loss = (S.lossw[0] + S.lossw[1]) + (S.lossw[2] + S.lossw[3]);
        if (tid == 0) {
            if (loss != nullptr) loss[b] = (tgt != nullptr) ? my_loss / ((double)T * (double)K) : 0.0;
            if (status != nullptr) status[b] = capped ? -1 : n_apply;
        }
        } else {
        int kidx = 0;
        while (CUNI(kidx < T - 1)) {
            int q, nsub;
            double htot;
            {
                const double t0 = P.times[kidx];
                q = 1;
                while (q < CQ && kidx + q < T - 1 && nu * (P.times[kidx + q + 1] - t0) <= P.theta_target) ++q;
                htot = P.times[kidx + q] - t0;
                nsub = 1;
                if (q == 1) {
                    const double th = nu * htot;
                    if (th > P.theta_target) nsub = (int)ceil(th / P.theta_target);
                }
            }
            if (CUNI(htot == 0.0)) {     // repeated time points: the state does not move
                __syncthreads();
                partial_trace_store(acc_re, acc_im);
                __syncthreads();
                const double tr = trace_total();
                emit(q, kidx + 1, tr, tr);
                kidx += q;
                continue;
            }
            const double h = htot / nsub;
            const double t0 = P.times[kidx];
            const int p0 = lane & 7, p1 = p0 + 8;
            const double xs0 = (p0 < q) ? ((nsub == 1) ? (P.times[kidx + 1 + p0] - t0) / htot : 1.0) : 0.0;
            const double xs1 = (p1 < q) ? (P.times[kidx + 1 + p1] - t0) / htot : 0.0;
            double y0 = 0.0, y1 = 0.0;

            for (int sub = 0; CUNI(sub < nsub); ++sub) {
                __syncthreads();     // everybody is done reading S.trp / S.mxs of the previous expansion
                const bool last = (sub == nsub - 1);
                double v_re[4][2], v_im[4][2];
#pragma unroll
                for (int J = 0; J < 4; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) { v_re[J][e] = acc_re[J][e]; v_im[J][e] = acc_im[J][e]; }
                if (CUNI(h != h_cur)) {
                    const double rs = h / h_cur;
                    h_cur = h;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) { a_re[ks] *= rs; a_im[ks] *= rs; }
#pragma unroll
                    for (int J = 0; J < 4; ++J) { W0[J][0] *= rs; W0[J][1] *= rs; }
                    F4[0] *= rs; F4[1] *= rs; F3[0] *= rs; F3[1] *= rs;
                    f2 *= rs; f1 *= rs; f0[0] *= rs; f0[1] *= rs;
                }
                double xp0 = 1.0, xp1 = 1.0;      // x^j / j!
                double cfac = 1.0;                // 1 / j!
                double thr = P.tol;               // tol * j!
                if (last) { y0 = 0.0; y1 = 0.0; }
                int small = 0;
                unsigned mx_term = 0x7fffffffu;     // size of the term about to be published (term 0: the state itself)
                unsigned thr_term = 0;
                int j = 0;
                for (;; ++j) {
                    // ---- [A] publish term j: rows of this warp into Vs, partial trace, size
#pragma unroll
                    for (int J = 0; J < 4; ++J) {
                        *reinterpret_cast<double2*>(&S.Vs_re[(8 * w + g) * LDV + 8 * J + 2 * t]) = make_double2(v_re[J][0], v_re[J][1]);
                        *reinterpret_cast<double2*>(&S.Vs_im[(8 * w + g) * LDV + 8 * J + 2 * t]) = make_double2(v_im[J][0], v_im[J][1]);
                    }
                    partial_trace_store(v_re, v_im);
                    if (lane == 0) S.mxs[w] = mx_term;
                    __syncthreads();
                    {
                        const double tr = trace_total();
                        y0 = fma(xp0, tr, y0);
                        y1 = fma(xp1, tr, y1);
                        const double inv = CTA_INV_TABLE[j];
                        xp0 *= xs0 * inv; xp1 *= xs1 * inv;
                    }
                    const unsigned mx = max(max(S.mxs[0], S.mxs[1]), max(S.mxs[2], S.mxs[3]));
                    small = (mx <= thr_term) ? small + 1 : 0;
                    if (CUNI(small >= 2 || j >= MCAP)) break;   // uniform over the CTA (values read from shared memory)

                    // ---- [B] accumulators = J(v)/2 (all five site stencils read from Vs), then += G v on the DMMA pipe
                    double x_re[4][2], x_im[4][2];
#pragma unroll
                    for (int J = 0; J < 4; ++J)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            x_re[J][e] = W0[J][e] * v_re[J][e];
                            x_im[J][e] = W0[J][e] * v_im[J][e];
                        }
                    apply_core(v_re, v_im, x_re, x_im);
                    __syncthreads();

                    // ---- [C] next term = f (X' + X'^dag):  X'^dag[8w+g][8J+2t+e] = conj(X'[8J+2t+e][8w+g])
                    cfac *= CTA_INV_TABLE[j];
                    thr *= (double)(j + 1);
                    thr_term = (unsigned)__double2hiint(thr) & 0x7fffffffu;
                    unsigned mxl = 0;
                    transpose_add(x_re, x_im, v_re, v_im);
#pragma unroll
                    for (int J = 0; J < 4; ++J)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            acc_re[J][e] = fma(cfac, v_re[J][e], acc_re[J][e]);
                            acc_im[J][e] = fma(cfac, v_im[J][e], acc_im[J][e]);
                            mxl = max(mxl, (unsigned)__double2hiint(v_re[J][e]) & 0x7fffffffu);
                            mxl = max(mxl, (unsigned)__double2hiint(v_im[J][e]) & 0x7fffffffu);
                        }
                    mx_term = __reduce_max_sync(0xffffffffu, mxl);
                }
                n_apply += j;
                if (j >= MCAP && small < 2) capped = true;
            }
            emit(q, kidx + 1, y0, y1);
            kidx += q;
        }

        if (w == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
            if (lane == 0) {
                if (loss != nullptr) loss[b] = (tgt != nullptr) ? my_loss / ((double)T * (double)K) : 0.0;
                if (status != nullptr) status[b] = capped ? -1 : n_apply;
            }
        }
        }
    }
}

#define EML_CTA_INST(RH, CH, PA) template __global__ void eval_cta_mma_kernel<RH, CH, PA>(const DevProblem, const long long, const double*, const int*, double*, double*, int*, unsigned int*, const int*, const double);
EML_CTA_INST(true, false, false) EML_CTA_INST(false, false, false) EML_CTA_INST(true, true, false) EML_CTA_INST(false, true, false)
EML_CTA_INST(true, true, true) EML_CTA_INST(false, true, true)

bool cta_mma_supported(const DevProblem& P, bool hermitian) { return hermitian && P.n_tls == 5 && P.n_params <= CMAXP; }

// eml_cta_orbit.cu: the second-generation kernel for parity-conserving Hamiltonians (Chebyshev series)
cudaError_t launch_cta_orbit(const DevProblem& P, long long n_models, const double* params, const int* target_set, double* expect,
                             double* loss, int* status, unsigned int* counter, int* order, bool real_h, double* prep_ws,
                             long long prep_cap, cudaStream_t stream);
bool cta_orbit_selected() {
    static const bool tilerow = [] { const char* e = std::getenv("EML_D32_TILEROW"); return e && e[0] == '1'; }();
    return !tilerow;
}

// declared in eml_warp_mma.cu
void launch_model_order(const DevProblem& P, long long n_models, const double* params, int* order, const int* cost_hint,
                        cudaStream_t stream);

cudaError_t launch_cta_mma(const DevProblem& P, long long n_models, const double* params, const int* target_set,
                           double* expect, double* loss, int* status, unsigned int* counter, int* order, bool real_h,
                           bool chebyshev, bool parity_h, const int* cost_hint, double* prep_ws, long long prep_cap, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    {
        static bool done[64] = {false};
        if (dev >= 64 || !done[dev]) {
            double hinv[MCAP + 2];
            for (int j = 0; j < MCAP + 2; ++j) hinv[j] = 1.0 / (double)(j + 1);
            if ((e = cudaMemcpyToSymbol(CTA_INV_TABLE, hinv, sizeof(hinv))) != cudaSuccess) return e;
            if (dev < 64) done[dev] = true;
        }
    }
    // parity-conserving H: the orbit kernel with its own prepare kernel (EML_D32_TILEROW=1 keeps the first-generation
    // tile-row kernel, for A/B measurements)
    if (parity_h && chebyshev && cta_orbit_selected())
        return launch_cta_orbit(P, n_models, params, target_set, expect, loss, status, counter, order, real_h, prep_ws, prep_cap, stream);
    if ((e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream)) != cudaSuccess) return e;
    const int* use_order = nullptr;
    if (order != nullptr && n_models > 2LL * sms && n_models <= (1 << 18)) {
        launch_model_order(P, n_models, params, order, cost_hint, stream);
        use_order = order;
    }
    using Kern = void (*)(const DevProblem, const long long, const double*, const int*, double*, double*, int*, unsigned int*,
                          const int*, const double);
    const bool parity = parity_h && chebyshev;
    const Kern kern = parity ? (real_h ? (Kern)eval_cta_mma_kernel<true, true, true> : (Kern)eval_cta_mma_kernel<false, true, true>)
                    : chebyshev ? (real_h ? (Kern)eval_cta_mma_kernel<true, true, false> : (Kern)eval_cta_mma_kernel<false, true, false>)
                                : (real_h ? (Kern)eval_cta_mma_kernel<true, false, false> : (Kern)eval_cta_mma_kernel<false, false, false>);
    const size_t smem = chebyshev ? sizeof(CtaScratch<true>) : sizeof(CtaScratch<false>);
    static int per_sm_cache[64][8] = {{0}};
    const int variant = (parity ? 4 : 0) + (chebyshev ? 2 : 0) + (real_h ? 1 : 0);
    int per_sm = (dev < 64) ? per_sm_cache[dev][variant] : 0;
    if (per_sm == 0) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem)) != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        if (dev < 64) per_sm_cache[dev][variant] = per_sm;
    }
    long long ctas = n_models;
    if (ctas > (long long)per_sm * sms) ctas = (long long)per_sm * sms;
    const double zcap = chebyshev ? cheb_zcap_for(CTA_CHEB_KCAP) : 0.0;
    kern<<<(unsigned)ctas, 128, smem, stream>>>(P, n_models, params, target_set, expect, loss, status, counter, use_order, zcap);
    return cudaGetLastError();
}

}  // namespace eml
