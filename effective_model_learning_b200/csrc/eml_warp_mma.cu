// d = 8 and d = 16 (n_tls = 3, 4; smaller models are embedded) evaluator: ONE WARP PER MODEL, density matrix and
// series terms resident in registers in the FP64 mma.m8n8k4 accumulator layout, generator product on the DMMA pipe.
//
//   L rho = X + X^dag + J(rho),  X = G rho  (rho Hermitian => rho G^dag = X^dag),  G = -iH - 1/2 sum c^dag c
//   J(rho)[r,c] = W0[r,c] rho[r,c] + sum_sites F_s[r_s,c_s] rho[r ^ m_s, c ^ m_s]   (real weights)
//
// Layout ("C layout"): lane l = 4g + t (g = l/4, t = l%4) holds, for tile (I,J) and e in {0,1}, the element
// (r, c) = (8I + g, 8J + 2t + e) -- exactly the m8n8 accumulator fragment, so X lands where rho lives.
// Because rho is Hermitian the B operand rho[k][n] = conj(rho[n][k]) is ALSO in the lane's own registers when
// the k index of k-step ks is permuted to pi(ks,t) = 8(ks/2) + 2t + (ks%2); G's A fragments are loaded once per
// model with the same permutation.  X^dag is gathered through a swizzled, conflict-free shared-memory round trip
// (or 32 shuffles), the site stencils cost 32 shuffles per site (none for the site mapped to the tile bits).
// The accumulators of the DMMA start from J(u)/2 (weights stored halved), so X' + X'^dag = G u + u G^dag + J(u).
//
// Two series turn generator applications into the propagator action (template flag CHEB):
//   CHEB = true (default, kernel ids "warp_cheb*"): Chebyshev / Bessel series, see eml_cheb.cuh -- one three-term
//     recurrence serves every output time of a macro step; only the 2x2 partial trace of each term is kept (shared
//     memory) and each output is a Bessel-weighted sum of them (Miller's backward recurrence, one lane per output).
//   CHEB = false ("warp_mma*", the independent cross-check): Taylor series with dense output -- the generator is kept
//     pre-multiplied by the step length h, u_{j+1} = (hL) u_j, rho(t_s + x h) = sum_j (x^j / j!) u_j, NY dense-output
//     accumulators per lane, a macro step is limited to nu h <= 16 by the e^theta cancellation hump.
// PARITY (d = 16, CHEB): H block diagonal by the parity of the basis index -> half the DMMAs (see the kernel body).
//
// Control flow is made provably warp-uniform with votes (UNI), which keeps ptxas from guarding every shuffle
// against divergence and lets it schedule SHFL / DFMA / DMMA of one application in a single basic block.
//
// TL = d/8 tiles per side: d = 16 has 2x2 tiles per lane (the observed site, or the parity, on the tile bits),
// d = 8 a single tile (all three site stencils go through shuffles).
//
// Requirements (checked at plan creation, else the generic kernel runs): n_tls == 3 or 4, Hermitian H and rho0,
// every jump operator A has A (x) conj(A) real with only "diagonal" and "both bits flipped" entries
// (sigmax/y/z/p/m, exc, gnd, id2, mm all qualify).
//
// Replaces qutip.mesolve + e_ops + loss for one model (reference basic_model.py:381-388, learning_chain.py:738-742).
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "eml_common.cuh"
#include "eml_cheb.cuh"

namespace eml {

// outputs per expansion: NY per lane (p = (lane&7) + 8n); d = 16 has registers for 2 (16 outputs), d = 8 for 4
// (32 outputs -- small models have small nu * dt, so long expansions need many outputs to reach theta)
#ifndef EML_WARPS_PER_CTA
#define EML_WARPS_PER_CTA 4
#endif
constexpr int WARPS_PER_CTA = EML_WARPS_PER_CTA;
#ifndef EML_CTAS_PER_SM
#define EML_CTAS_PER_SM 2
#endif
constexpr int CTAS_PER_SM = EML_CTAS_PER_SM;   // 2 -> 8 warps per SM (<= 255 regs); 3 -> 12 warps (<= 168 regs)
constexpr int MAXP = WARP_MAXP; // parameter-vector cap for this kernel

// Chebyshev series: terms per macro step that fit the per-warp trace buffer (8 warps x 26.4 KB at d = 16,
// 16 warps x 13 KB at d <= 8), and the largest  tau * (A + B)  whose series fits (cheb_terms(z) <= cap)
#ifndef EML_CHEB_KCAP_TL2
#define EML_CHEB_KCAP_TL2 512
#endif
// kernel-variant switches for experiments (tools/build_variant.sh); the defaults are the measured best
#ifndef EML_CHEB_UNROLL
#define EML_CHEB_UNROLL 1          // unroll factor of the Chebyshev recurrence loop
#endif
#ifndef EML_HALVE_ON_FP64
#define EML_HALVE_ON_FP64 2        // halving of the stored term: 0 integer-pipe exponent decrement, 1 DMUL, 2 DMUL in the
#endif                             // parity-blocked variant only (its FP64 pipe has room, its issue slots do not): -2.7 %
#ifndef EML_TL2_SMEM_STENCIL
#define EML_TL2_SMEM_STENCIL 0     // 1: d = 16 lane-bit site stencils read the published term from shared memory
#endif                             // (8 STS.128 + 24 LDS.128) instead of 96 SHFL
#ifndef EML_TL2_SMEM_TRANSPOSE
#define EML_TL2_SMEM_TRANSPOSE 1   // d = 16 conjugate transpose through shared memory (as d = 8) instead of 32 SHFL + 64 FSEL:
#endif                             // -2.3 % with the parity-blocked variant (was +1 % when the kernel was FP64-pipe bound)
constexpr int CHEB_KCAP_TL2 = EML_CHEB_KCAP_TL2, CHEB_KCAP_TL1 = 256;
constexpr int CHEB_UNROLL = EML_CHEB_UNROLL;
#ifndef EML_CHEB_KCAP_SB
#define EML_CHEB_KCAP_SB 488
#endif
#ifndef EML_OUTPUT_CHAINS
#define EML_OUTPUT_CHAINS 1        // output times per lane and pass of the Bessel-weighted sums (measured: 2 chains +-0, 4 chains -6 %, 8 chains -22 %: the phase is bound by the FP64 pipe it shares with the recurrences of the other warps, not by the latency of its dependent FMA chain)
#endif
#ifndef EML_LAT_OUTPUT_CHAINS
#define EML_LAT_OUTPUT_CHAINS 4    // output times per lane and pass in the latency variant (launches with fewer models than warp slots)
#endif
#ifndef EML_LAT_WARPS_PER_SM
#define EML_LAT_WARPS_PER_SM 8     // resident warps per SM of the latency variant (<= 255 registers)
#endif
#ifndef EML_WARPS_PER_SM_SB
#define EML_WARPS_PER_SM_SB 16
#endif
#ifndef EML_WARPS_PER_SM_PAR
#define EML_WARPS_PER_SM_PAR 8      // full-state parity variant (d = 16): resident warps per SM (8: <= 255 regs, 12: <= 168 regs)
#endif
#ifndef EML_CHEB_KCAP_PAR
#define EML_CHEB_KCAP_PAR 512
#endif
#ifndef EML_PAR_XT_BUFFERS
#define EML_PAR_XT_BUFFERS 2        // transposition buffers of the full-state parity variant (1: a second __syncwarp per application)
#endif
constexpr int WARPS_PER_SM_PAR = EML_WARPS_PER_SM_PAR, CHEB_KCAP_PAR = EML_CHEB_KCAP_PAR, PAR_XT_BUFFERS = EML_PAR_XT_BUFFERS;
constexpr int CHEB_KCAP_SB = EML_CHEB_KCAP_SB;      // sector variant: 16 warps per SM (<= 128 regs, 14 KB shared memory per warp), two trace values per term
constexpr int WARPS_PER_SM_SB = EML_WARPS_PER_SM_SB;
static_assert(WARPS_PER_SM_SB % EML_WARPS_PER_CTA == 0, "resident warps per SM must be a whole number of CTAs");

// Per-warp shared memory.  The Taylor variants assemble the generator themselves (G, parameters, norm-bound scratch);
// the Chebyshev variants receive it from the prepare kernel and keep the trace table, the coefficient table of a
// non-final macro step and the transposition tile instead.
template <int TL, bool CHEB, bool SB = false, bool PA = false>
struct alignas(16) WarpScratch;
template <int TL>
struct alignas(16) WarpScratch<TL, false, false, false> {
    static constexpr int D = 8 * TL, KCAP = 0, NTR = 4;
    double G_re[D][D + 1];
    double G_im[D][D + 1];
    double par[MAXP];
    double jw[32];      // [bitpos][type(0 diag,1 flip)][ra][ca]
    double kd[8];       // [bitpos][ra]: diagonal of sum rate * A^dag A per site
    double colsum[16];
    double misc[4];
    static constexpr int XT = TL == 1 ? 128 : (EML_TL2_SMEM_STENCIL ? 16 * 24 : (EML_TL2_SMEM_TRANSPOSE ? 256 : 2));
    alignas(16) double xt_re[XT];
    alignas(16) double xt_im[XT];   // X' staged column-major (swizzled) for the conjugate transpose
    alignas(16) double tr[4];
};
template <int TL, bool SB, bool PA>
struct alignas(16) WarpScratch<TL, true, SB, PA> {
    static constexpr int D = 8 * TL;
    static constexpr int KCAP = TL == 2 ? (SB ? CHEB_KCAP_SB : (PA ? CHEB_KCAP_PAR : CHEB_KCAP_TL2)) : CHEB_KCAP_TL1;
    static constexpr int NTR = SB ? 2 : 4;      // partial-trace values kept per term
    double jw[32];      // [bitpos][type(0 diag,1 flip)][ra][ca]          (from the prepare kernel's record)
    double gdiag[16];   // real H: diagonal of G.re = -1/2 sum c^dag c     (contiguous with plan: one record read)
    double plan[4];     // R, S, ac, tau_max
    double half_tr[2];  // sector variant: Tr(rho0) / 2 (rho00 = rho11 of the observed site never move)
    static constexpr int XT = TL == 1 ? 128 : (EML_TL2_SMEM_STENCIL ? 16 * 24 : (EML_TL2_SMEM_TRANSPOSE ? 256 : 2));
    // SB: one 8 x 8 complex tile, planes re | im of 64 doubles, two buffers (256 doubles in xt_re; xt_im unused)
    // (parity variants: 8 x 8 tiles, planes re | im of 64 doubles, two buffers -- one tile in the sector variant, three otherwise)
    alignas(16) double xt_re[SB ? 256 : (PA ? 384 * PAR_XT_BUFFERS : XT)];
    alignas(16) double xt_im[(SB || PA) ? 2 : XT];   // X' staged column-major (swizzled) for the conjugate transpose
    double cj[KCAP + 2];                     // Bessel coefficients of the step end (non-final macro steps only)
    alignas(16) double tr[(KCAP + 4) * NTR]; // partial traces of the Chebyshev terms (+3 zero entries: groups of four)
};
// The prepare kernel's per-warp scratch: generator, parameters, and the two matrices of the spread-bound squarings.
template <int TL, bool REAL_H>
struct alignas(16) PrepScratch {
    static constexpr int D = 8 * TL;
    double G_re[D][D + 1];
    double G_im[D][D + 1];
    double par[MAXP];
    double jw[32];
    double kd[8];
    double colsum[16];
    double misc[4];
    alignas(16) double mat[REAL_H ? 2 * D * (D + 4) : 4 * D * D];      // real H: row stride D + 4
};

// Phase clocks (profiling builds only, -DEML_PROFILE_PHASES): warp-cycles spent in set-up, recurrence, outputs, epilogue,
// summed over all warps, plus the cycles between a warp's last model and the end of the launch are NOT included.
__device__ unsigned long long g_phase_clk[8];
#ifdef EML_PROFILE_PHASES
#define PHASE_MARK(slot) do { const long long _now = clock64(); if (lane == 0) atomicAdd(&g_phase_clk[slot], (unsigned long long)(_now - phase_t)); phase_t = _now; } while (0)
#else
#define PHASE_MARK(slot) do { } while (0)
#endif
cudaError_t read_phase_clocks(unsigned long long* out, bool reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_phase_clk, sizeof(g_phase_clk));
    if (e == cudaSuccess && reset) {
        const unsigned long long zero[8] = {0};
        e = cudaMemcpyToSymbol(g_phase_clk, zero, sizeof(zero));
    }
    return e;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
// 1/(j+1) for the Taylor recurrences
__constant__ double INV_TABLE[MCAP + 2];

// warp-uniform predicate the compiler can SEE is uniform (vote result): keeps shuffles free of divergence guards
#define UNI(c) (__all_sync(0xffffffffu, (c)) != 0)

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double neg(double v) { return __hiloint2double(__double2hiint(v) ^ 0x80000000, __double2loint(v)); }

// host-side structural test of the jump operators: A (x) conj(A) restricted to diag / both-flipped, real
bool jump_op_is_diag_flip_real(const double* op) {
    for (int ra = 0; ra < 2; ++ra)
        for (int ca = 0; ca < 2; ++ca)
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    const double xr = op[(ra * 2 + a) * 2], xi = op[(ra * 2 + a) * 2 + 1];
                    const double yr = op[(ca * 2 + b) * 2], yi = -op[(ca * 2 + b) * 2 + 1];
                    const double re = xr * yr - xi * yi, im = xr * yi + xi * yr;
                    const bool diag = (a == ra && b == ca), flip = (a != ra && b != ca);
                    if (fabs(im) > 1e-15) return false;
                    if (!diag && !flip && fabs(re) > 1e-15) return false;
                }
    return true;
}

// Longest-processing-time-first ordering: the work of a model grows with the size of its generator, so the
// persistent warps fetch models in order of decreasing  sum |energy| + sum strength^2 + sum rate  (a proxy for
// the norm bound): a cost kernel (one thread per model) and a one-CTA counting sort.  Without it the last warps to finish
// determine the kernel time (3.5 models per warp at 4096 chains).  These two kernels serve the Taylor variants (proxy cost)
// and Chebyshev launches beyond ORDER_LIST_CAP models (exact cost from the prepare kernel); Chebyshev launches up to
// ORDER_LIST_CAP models are ordered by the prepare kernel's cost-bin lists instead (order_list_put / order_list_get below).
constexpr int ORDER_BINS = 256;
constexpr int ORDER_MAX_MODELS = 1 << 18;

// pass 1: one thread per model, cost proxy into scratch
__global__ void __launch_bounds__(256) model_cost_kernel(const DevProblem P, const int n_models,
                                                         const double* __restrict__ params, float* __restrict__ cost) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_models) return;
    float c = 0.f;
    for (int ti = 0; ti < P.n_terms; ++ti) {
        const EmlTerm tm = P.terms[ti];
        const float v = (float)params[(size_t)b * P.n_params + tm.param];
        c += (tm.kind == EML_TERM_COUPLING) ? v * v : fabsf(v);
    }
    cost[b] = c;
}

// pass 2: one CTA, 256-bin counting sort of the costs, largest first
__global__ void __launch_bounds__(1024) order_models_kernel(const int n_models, const float* __restrict__ cost,
                                                            int* __restrict__ order) {
    __shared__ unsigned int hist[ORDER_BINS];
    __shared__ unsigned int start[ORDER_BINS];
    __shared__ float smax;
    const int tid = threadIdx.x;
    if (tid < ORDER_BINS) hist[tid] = 0;
    if (tid == 0) smax = 0.f;
    __syncthreads();
    float mymax = 0.f;
    for (int b = tid; b < n_models; b += blockDim.x) mymax = fmaxf(mymax, cost[b]);
    atomicMax((int*)&smax, __float_as_int(mymax));      // non-negative floats order like ints
    __syncthreads();
    const float scale = (smax > 0.f) ? (ORDER_BINS - 1) / smax : 0.f;
    for (int b = tid; b < n_models; b += blockDim.x) atomicAdd(&hist[ORDER_BINS - 1 - (int)(cost[b] * scale)], 1u);
    __syncthreads();
    if (tid < 32) {     // exclusive scan of the 256 bins by one warp: 8 bins per lane, then a shuffle scan of the lane totals
        unsigned int loc[ORDER_BINS / 32], tot = 0;
#pragma unroll
        for (int i = 0; i < ORDER_BINS / 32; ++i) { loc[i] = tot; tot += hist[tid * (ORDER_BINS / 32) + i]; }
        unsigned int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += up;
        }
        const unsigned int base = inc - tot;
#pragma unroll
        for (int i = 0; i < ORDER_BINS / 32; ++i) start[tid * (ORDER_BINS / 32) + i] = base + loc[i];
    }
    __syncthreads();
    for (int b = tid; b < n_models; b += blockDim.x) {
        const unsigned int slot = atomicAdd(&start[ORDER_BINS - 1 - (int)(cost[b] * scale)], 1u);
        order[slot] = b;
    }
}

// ---- K1 (reference basic_model.py:152-255): G = -iH - 1/2 sum c^dag c of one model into shared memory (lane r < D owns row r),
//      the real jump weights jw (one entry per lane) and, for the Chebyshev enclosure, the diagonal of sum rate * A^dag A per
//      site; returns the norm bound  nu >= ||L||_1.  Shared by the prepare kernel and the Taylor evaluators.
template <int TL, bool WANT_KD, class SC>
__device__ __forceinline__ double assemble_generator(SC& S, const DevProblem& P, const double* __restrict__ prow, const int lane) {
    constexpr int N = 2 + TL, D = 8 * TL;
    const int obs_site = P.obs_site;
    auto pos = [&](int s) { return (s == obs_site) ? N - 1 : (s < obs_site ? N - 2 - s : N - 1 - s); };
    // ---- parameters, clear G
    for (int i = lane; i < P.n_params; i += 32) S.par[i] = prow[i];
    for (int i = lane; i < D * (D + 1); i += 32) { (&S.G_re[0][0])[i] = 0.0; (&S.G_im[0][0])[i] = 0.0; }
    __syncwarp();

    // ---- K1: assemble G (lane r < 16 owns row r) and the real jump weights (one entry per lane)
    if (lane < D) {
        const int r = lane;
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
                    S.G_re[r][c] += v * A.im;        // -i h
                    S.G_im[r][c] -= v * A.re;
                }
            } else if (tm.kind == EML_TERM_COUPLING) {
                const int sb = pos(tm.site_b);
                const int rb = (r >> sb) & 1;
                const double v2 = v * v;
                for (int a = 0; a < 2; ++a)
                    for (int bb = 0; bb < 2; ++bb) {
                        const cplx AB = cmul(ld_op(P.op_table, tm.op_a, ra, a), ld_op(P.op_table, tm.op_b, rb, bb));
                        const int c = (r & ~((1 << sa) | (1 << sb))) | (a << sa) | (bb << sb);
                        S.G_re[r][c] += v2 * AB.im;
                        S.G_im[r][c] -= v2 * AB.re;
                    }
            } else {
                for (int a = 0; a < 2; ++a) {
                    cplx M = {0.0, 0.0};
                    for (int z = 0; z < 2; ++z) {
                        const cplx m = cmul(cconj(ld_op(P.op_table, tm.op_a, z, ra)), ld_op(P.op_table, tm.op_a, z, a));
                        M.re += m.re; M.im += m.im;
                    }
                    const int c = (r & ~(1 << sa)) | (a << sa);
                    S.G_re[r][c] -= 0.5 * v * M.re;
                    S.G_im[r][c] -= 0.5 * v * M.im;
                }
            }
        }
    }
    {
        const int bp = lane >> 3, type = (lane >> 2) & 1, ra = (lane >> 1) & 1, ca = lane & 1;
        const int a = type ? 1 - ra : ra, bb = type ? 1 - ca : ca;
        // the lanes (type 0, ca 0) also collect the diagonal of sum rate * A^dag A of their site and row bit (the
        // anticommutator part of the site generator, needed by the Chebyshev enclosure)
        const bool kd_lane = WANT_KD && type == 0 && ca == 0;
        double w = 0.0, kdv = 0.0;
        for (int ti = 0; ti < P.n_terms && bp < N; ++ti) {
            const EmlTerm tm = P.terms[ti];
            if (tm.kind != EML_TERM_LINDBLAD || pos(tm.site_a) != bp) continue;
            const double v = S.par[tm.param];
            if (v == 0.0) continue;
            const cplx x = cmul(ld_op(P.op_table, tm.op_a, ra, a), cconj(ld_op(P.op_table, tm.op_a, ca, bb)));
            w += v * x.re;
            if (kd_lane) {
                const cplx m0 = ld_op(P.op_table, tm.op_a, 0, ra), m1 = ld_op(P.op_table, tm.op_a, 1, ra);
                kdv += v * ((m0.re * m0.re + m0.im * m0.im) + (m1.re * m1.re + m1.im * m1.im));
            }
        }
        S.jw[lane] = w;
        if (kd_lane) S.kd[bp * 2 + ra] = kdv;
    }
    __syncwarp();

    // ---- norm bound: 2 * max column sum |G| (G^dag supplies the row-sum term) + sum_s ||S_s||_1
    if (lane < D) {
        double cs = 0.0;
        for (int i = 0; i < D; ++i) cs += hypot(S.G_re[i][lane], S.G_im[i][lane]);
        S.colsum[lane] = cs;   // vec(rho G^dag) = (conj(G) (x) I) vec(rho) has the same 1-norm: factor 2 below
    }
    __syncwarp();
    if (lane == 0) {
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
    __syncwarp();
    return S.misc[0];
}

// ---- prepare kernel (Chebyshev variants): one warp per model, high occupancy.  Assembles the generator, bounds its spectrum
//      (Gershgorin and 2 ||(H - mean diag)^8||_1^(1/8) for the commutator, per-site ranges for the dissipator), chooses the
//      Chebyshev plan and writes ONE RECORD per model for the evaluator:
//        [NA][32]  A fragments of G per lane (im plane, then re plane for a complex H), k index permuted as the evaluator
//                  multiplies them;  [32] jw;  [16] diagonal of G.re;  [4] R, S, ac, tau_max  (R = NaN: not evaluable)
//      and the number of generator applications the evaluation will need (exact for a single macro step): the cost the
//      persistent warps of the evaluator are ordered by.  Persistent CTAs: the plan's linear program (eml_common.cuh) is
//      staged in shared memory once per CTA, then every warp takes models gw, gw + (warps of the grid), ...; the scalar
//      tail (site ranges, the six enclosure candidates) is spread over lanes (eml_cheb.cuh: *_warp).
// ---- Longest-first order WITHOUT an ordering pass (launches of at most ORDER_LIST_CAP models): the prepare kernel drops
//      every model into the list of its cost bin (bins of (kcap + 1) / ORDER_BINS generator applications, most expensive
//      first; multi-step models beyond kcap applications share bin 0); an evaluator warp turns its fetch ticket into
//      (bin, position) with a prefix sum over the bin counts (one 32-byte load per lane and a shuffle scan per fetch).
//      Workspace words: [0] the fetch counter, [ORDER_HIST_OFF ...] the bin counts -- zeroed by ONE memset node per launch.
constexpr int ORDER_LIST_CAP = 4096;
constexpr int ORDER_HIST_OFF = 4;      // 16-byte aligned behind the counter
static_assert(ORDER_BINS * ORDER_LIST_CAP <= 2 * ORDER_MAX_MODELS * 2, "the bin lists live in the ordering buffer (2^20 ints)");
static_assert(ORDER_BINS == 256, "eight bins per lane");
__device__ __forceinline__ void order_list_put(unsigned int* __restrict__ fetch_ws, int* __restrict__ lists, const float cost,
                                               const float bin_scale, const int b) {
    const int bin = ORDER_BINS - 1 - min(ORDER_BINS - 1, (int)(cost * bin_scale));
    const unsigned int at = atomicAdd(&fetch_ws[ORDER_HIST_OFF + bin], 1u);
    EML_CHECK(at < (unsigned int)ORDER_LIST_CAP);
    lists[bin * ORDER_LIST_CAP + (int)at] = b;
}
// model of fetch ticket `rank` (< the number of models of the launch); every lane returns it
__device__ __forceinline__ int order_list_get(const unsigned int* __restrict__ fetch_ws, const int* __restrict__ lists,
                                              const unsigned int rank, const int lane) {
    const uint4* const h = reinterpret_cast<const uint4*>(fetch_ws + ORDER_HIST_OFF) + 2 * lane;
    const uint4 a = __ldcg(h), c = __ldcg(h + 1);
    const unsigned int cnt[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    unsigned int tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += cnt[i];
    unsigned int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    const unsigned int before = inc - tot;
    int found = -1;
    if (rank >= before && rank < inc) {
        unsigned int r = rank - before;
        int i = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (i == j && r >= cnt[j]) { r -= cnt[j]; i = j + 1; }
        found = ((8 * lane + i) * ORDER_LIST_CAP) + (int)r;
    }
    const unsigned int who = __ballot_sync(0xffffffffu, found >= 0);
    EML_CHECK(who != 0u);
    found = __shfl_sync(0xffffffffu, found, __ffs(who) - 1);
    return lists[found];
}

// The fetch workspace cleans itself: each warp of an evaluator grid draws exactly one ticket >= n_units (its exit ticket), so
// the warp that draws ticket n_units + (warps of the grid) - 1 is the last one to touch the counter and the bin counts
// and zeroes them for the next launch (no memset node per launch; the workspace is zero when it is allocated).
__device__ __forceinline__ void fetch_ws_release(unsigned int* __restrict__ fetch_ws, const unsigned int ticket, const long long n_units,
                                                 const int binned, const int lane) {
    const long long last = n_units + (long long)gridDim.x * (blockDim.x >> 5) - 1;
    if ((long long)ticket != last) return;
    if (binned) {
        uint4* const h = reinterpret_cast<uint4*>(fetch_ws + ORDER_HIST_OFF) + 2 * lane;
        h[0] = make_uint4(0u, 0u, 0u, 0u); h[1] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (lane == 0) fetch_ws[0] = 0u;
}
// programmatic dependent launch: the evaluator may become resident while the prepare kernel drains and waits here for its
// records (a no-op for launches without the attribute)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int TL, bool REAL_H, bool PARITY>
__host__ __device__ constexpr int prep_fragments() { return (REAL_H ? 1 : 2) * TL * (PARITY ? 2 : 2 * TL); }
template <int TL, bool REAL_H, bool PARITY>
__host__ __device__ constexpr int prep_record_doubles() { return prep_fragments<TL, REAL_H, PARITY>() * 32 + 32 + 16 + 4; }
constexpr int PREP_WARPS = 4;
constexpr int PREP_LIN_STAGE_MAX = 40 * 1024;      // largest linear program staged in shared memory (else it is read through L1/L2)

template <int TL, bool REAL_H, bool PARITY>
__global__ void __launch_bounds__(PREP_WARPS * 32)
prepare_models_kernel(const DevProblem P, const long long n_models, const double* __restrict__ params, double* __restrict__ rec,
                      float* __restrict__ cost, int* __restrict__ lists, unsigned int* __restrict__ fetch_ws,
                      const double cheb_zcap, const int kcap, const int stage_lin) {
    // cost: per-model cost for the ordering kernel (launches beyond ORDER_LIST_CAP models), or lists: the cost-bin lists
    constexpr int N = 2 + TL, D = 8 * TL, KS = 2 * TL, KSE = PARITY ? 2 : KS;
    constexpr int NA = prep_fragments<TL, REAL_H, PARITY>(), REC = prep_record_doubles<TL, REAL_H, PARITY>();
    auto phys = [](const int x) { return PARITY ? x ^ (((x ^ (x >> 1) ^ (x >> 2)) & 1) << 3) : x; };
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Scratch = PrepScratch<TL, REAL_H>;
    Scratch& S = reinterpret_cast<Scratch*>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    pdl_launch_dependents();      // the evaluator's CTAs may take the place of this grid's CTAs as they exit (it waits for the records)
    // the linear program, staged behind the per-warp scratch (rows, then entries) when the launch reserved room for it
    const LinRow* lin_row = P.lin_row;
    const LinEntry* lin_ent = P.lin_ent;
    if (stage_lin) {
        int4* dst = reinterpret_cast<int4*>(smem_raw + PREP_WARPS * sizeof(Scratch));
        const int4* src_r = reinterpret_cast<const int4*>(P.lin_row);
        const int4* src_e = reinterpret_cast<const int4*>(P.lin_ent);
        for (int i = threadIdx.x; i < P.lin_rows; i += blockDim.x) dst[i] = src_r[i];
        for (int i = threadIdx.x; i < P.lin_nnz; i += blockDim.x) dst[P.lin_rows + i] = src_e[i];
        lin_row = reinterpret_cast<const LinRow*>(dst);
        lin_ent = reinterpret_cast<const LinEntry*>(dst + P.lin_rows);
        __syncthreads();
    }
    const double span = P.times[P.n_times - 1] - P.times[0];
    const float bin_scale = (float)ORDER_BINS / (float)(kcap + 1);
    const long long warps_in_grid = (long long)gridDim.x * PREP_WARPS;
    for (long long b = (long long)blockIdx.x * PREP_WARPS + (threadIdx.x >> 5); b < n_models; b += warps_in_grid) {
    double* const R = rec + (size_t)b * REC;
    double nu;
    if (P.lin_rows > 0) {
        // ---- K1 as a linear program (compiled once per plan, eml_capi.cu): every entry of G, jw, kd is a gather over the
        //      transformed parameters c_p (coupling strengths squared, reference basic_model.py:245,247) -- independent
        //      accumulations, rows of similar length in one pass of the warp
        const double* const prow = params + b * P.n_params;
        for (int i = lane; i < P.n_params; i += 32) { const double v = prow[i]; S.par[i] = P.lin_sq[i] ? v * v : v; }
        double* const Y = &S.G_re[0][0];
        static_assert(offsetof(Scratch, G_im) == sizeof(double) * D * (D + 1) && offsetof(Scratch, par) == sizeof(double) * 2 * D * (D + 1) &&
                      offsetof(Scratch, jw) == sizeof(double) * prep_off_jw(D, 0) && offsetof(Scratch, kd) == sizeof(double) * prep_off_kd(D, 0),
                      "the linear program addresses the scratch by these offsets");
        for (int i = lane; i < 2 * D * (D + 1); i += 32) Y[i] = 0.0;
        S.jw[lane] = 0.0;
        if (lane < 8) S.kd[lane] = 0.0;
        __syncwarp();
        for (int i = lane; i < P.lin_rows; i += 32) {
            const LinRow rw = lin_row[i];
            EML_CHECK(rw.offset >= 0 && rw.offset < (int)(sizeof(Scratch) / sizeof(double)) - (int)((REAL_H ? 2 : 4) * D * D) &&
                      rw.begin >= 0 && rw.end <= P.lin_nnz && rw.begin <= rw.end);
            double acc = 0.0;
            for (int k = rw.begin; k < rw.end; ++k) {
                const LinEntry e = lin_ent[k];
                acc = fma(e.val, S.par[e.col], acc);
            }
            Y[rw.offset] = acc;
        }
        __syncwarp();
        // norm bound for the guard below: 2 max row sum (|re| + |im|) of G + the jump part, NaNs carried along
        double rs = 0.0;
        if (lane < D)
            for (int c = 0; c < D; ++c) rs += fabs(S.G_re[lane][c]) + fabs(S.G_im[lane][c]);
        double jp = 0.0;
        if (lane < N) {
            for (int a = 0; a < 2; ++a)
                for (int bb = 0; bb < 2; ++bb)
                    jp = fmax(jp, fabs(S.jw[lane * 8 + a * 2 + bb]) + fabs(S.jw[lane * 8 + 4 + (1 - a) * 2 + (1 - bb)]));
        }
        double mx = rs, sum = rs + jp;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            jp += __shfl_xor_sync(0xffffffffu, jp, o);
        }
        nu = 2.0 * mx + jp + 0.0 * sum;
    } else {
        nu = assemble_generator<TL, true>(S, P, params + b * P.n_params, lane);
    }
    // ---- record: A fragments with the permuted k index (PARITY: only the diagonal tiles, kept in ks = 0, 1), weights
#pragma unroll
    for (int I = 0; I < TL; ++I)
#pragma unroll
        for (int ks = 0; ks < KSE; ++ks) {
            const int k = PARITY ? 8 * I + 2 * t + (ks & 1) : 8 * (ks >> 1) + 2 * t + (ks & 1);
            R[((0 * TL + I) * KSE + ks) * 32 + lane] = S.G_im[phys(8 * I + g)][phys(k)];
            if (!REAL_H) R[((1 * TL + I) * KSE + ks) * 32 + lane] = S.G_re[phys(8 * I + g)][phys(k)];
        }
    R[NA * 32 + lane] = S.jw[lane];
    if (lane < 16) R[NA * 32 + 32 + lane] = (lane < D) ? S.G_re[lane][lane] : 0.0;
    // ---- spectral enclosure
    double spread;
    {
        double hi = -1e300, lo = 1e300;
        if (lane < D) {
            double rad = 0.0;
            for (int c = 0; c < D; ++c)
                if (c != lane) rad += REAL_H ? fabs(S.G_im[lane][c]) : hypot(S.G_re[lane][c], S.G_im[lane][c]);
            const double ctr = -S.G_im[lane][lane];
            hi = ctr + rad; lo = ctr - rad;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        }
        spread = hi - lo;
        // Gershgorin overestimates the spread by ~1.5x at d = 16; the spectral radius of Hc = H - mean(diag) is bounded far
        // more tightly by ||Hc^8||_1^(1/8) (three squarings): lambda_max - lambda_min <= 2 rho(Hc), measured 1.09x the true
        // spread on the benchmark library.
        double mu = (lane < D) ? -S.G_im[lane][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mu += __shfl_xor_sync(0xffffffffu, mu, o);
        mu *= 1.0 / D;
        // (real H: row stride D + 4, which makes the fragment loads of the DMMA squarings bank-conflict free: with stride D
        //  the eight rows of an A fragment hit the same banks -- 56 % of the kernel's shared-memory wavefronts were conflicts)
        constexpr int LDM = REAL_H ? D + 4 : D;
        double* Ma = S.mat;                                   // [re | im] planes of D x D (real H: re only)
        double* Mb = S.mat + (REAL_H ? D * LDM : 2 * D * D);
        for (int i = lane; i < D * D; i += 32) {
            const int r = i / D, c = i % D;
            Ma[r * LDM + c] = -S.G_im[r][c] - ((r == c) ? mu : 0.0);
            if (!REAL_H) Ma[D * D + i] = (r == c) ? 0.0 : S.G_re[r][c];
        }
        __syncwarp();
        for (int sq = 0; sq < 3; ++sq) {
            if constexpr (REAL_H) {
                // real symmetric D x D squaring on the DMMA pipe: fragments straight from shared memory
#pragma unroll
                for (int I = 0; I < TL; ++I)
#pragma unroll
                    for (int J = 0; J < TL; ++J) {
                        double d0 = 0.0, d1 = 0.0;
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks)
                            dmma(d0, d1, Ma[(8 * I + g) * LDM + 4 * ks + t], Ma[(4 * ks + t) * LDM + 8 * J + g]);
                        *reinterpret_cast<double2*>(&Mb[(8 * I + g) * LDM + 8 * J + 2 * t]) = make_double2(d0, d1);
                    }
            } else
            for (int i = lane; i < D * D; i += 32) {
                const int r = i / D, c = i % D;
                double sre = 0.0, sim = 0.0;
#pragma unroll 4
                for (int k = 0; k < D; ++k) {
                    const double are = Ma[r * D + k], bre = Ma[k * D + c];
                    const double aim = Ma[D * D + r * D + k], bim = Ma[D * D + k * D + c];
                    sre = fma(are, bre, fma(-aim, bim, sre));
                    sim = fma(are, bim, fma(aim, bre, sim));
                }
                Mb[i] = sre;
                Mb[D * D + i] = sim;
            }
            __syncwarp();
            double* tmp = Ma; Ma = Mb; Mb = tmp;
        }
        double cs = 0.0;
        if (lane < D)
            for (int r = 0; r < D; ++r)
                cs += REAL_H ? fabs(Ma[r * LDM + lane]) : hypot(Ma[r * D + lane], Ma[D * D + r * D + lane]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cs = fmax(cs, __shfl_xor_sync(0xffffffffu, cs, o));
        spread = fmin(spread, 2.0 * sqrt(sqrt(sqrt(cs))) * (1.0 + 1e-12));
    }
    // a non-finite generator (NaN / inf parameters) or an absurd step count must not hang the device
    if (!(nu * span < 1e5)) {
        if (lane < 4) R[NA * 32 + 48 + lane] = nan("");
        if (lane == 0 && cost != nullptr) cost[b] = 0.f;
        if (lane == 0 && lists != nullptr) order_list_put(fetch_ws, lists, 0.f, bin_scale, (int)b);
    } else {
        double dlo, dhi, dasym;
        cheb_site_ranges_warp(S.jw, S.kd, N, lane, dlo, dhi, dasym);
        const ChebPlan plan = cheb_make_plan_warp(spread, dlo, dhi, dasym, span, cheb_zcap, kcap, lane);
        if (lane == 0) {
            R[NA * 32 + 48] = plan.R; R[NA * 32 + 49] = plan.S; R[NA * 32 + 50] = plan.ac; R[NA * 32 + 51] = plan.tau_max;
            if (cost != nullptr || lists != nullptr) {
                // generator applications: one macro step when the span fits (exact), else about span / tau_max full steps
                const double steps = ceil(span / plan.tau_max);
                const int k1 = min(cheb_terms(fmin(span, plan.tau_max) * plan.S), kcap);
                const float c = (float)(fmax(steps, 1.0) * (double)k1);
                if (cost != nullptr) cost[b] = c;
                else order_list_put(fetch_ws, lists, c, bin_scale, (int)b);
            }
        }
    }
    __syncwarp();      // the scratch is reused by this warp's next model
    }
}

// LAT ("latency variant", Chebyshev kernels): chosen when a launch has no more models than this variant has warp slots, i.e.
// when the FP64 pipe is mostly idle and a model's time is the latency of its own dependent chains: the Bessel-weighted
// output sums run EML_LAT_OUTPUT_CHAINS independent output times per lane (the same phase is bound by the pipe, and slower
// with several chains, when all 16 warps of an SM are busy), with the registers of 8 warps per SM.
template <int TL, bool REAL_H, bool CHEB, bool PARITY, bool SB, bool LAT = false>
// resident warps per SM: 16 (d = 8), 16 (d = 16 sector variant), 8 (d = 16); CTAS_PER_SM is quoted for four-warp CTAs
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, LAT ? EML_LAT_WARPS_PER_SM / WARPS_PER_CTA : (SB ? WARPS_PER_SM_SB : PARITY ? WARPS_PER_SM_PAR : 4 * ((TL == 1) ? 4 : CTAS_PER_SM)) / WARPS_PER_CTA)
eval_warp_mma_kernel(const DevProblem P, const long long n_models, const double* __restrict__ params,
                     const int* __restrict__ target_set, double* __restrict__ expect, double* __restrict__ loss,
                     int* __restrict__ status, unsigned int* __restrict__ counter, const int* __restrict__ order,
                     const double* __restrict__ prep, const int binned) {
    constexpr int N = 2 + TL, D = 8 * TL, KS = 2 * TL;      // sites, Hilbert dimension, k-steps of 4
    constexpr int NY = (TL == 1) ? 4 : 2, WQ = 8 * NY;      // dense-output accumulators per lane, outputs per expansion
    static_assert(!PARITY || (CHEB && TL == 2), "the parity-blocked variant exists for the d = 16 Chebyshev kernel");
    static_assert(!SB || PARITY, "the sector variant builds on the parity coordinates");
    // SB ("sector B only"): in the parity coordinates the generator maps tile (I, J) to itself (commutator) or to
    // (I ^ 1, J ^ 1) (site flips), so the parity-even tiles (0,0), (1,1) and the parity-odd tiles (0,1), (1,0) evolve
    // INDEPENDENTLY.  When the parity-even part of rho0 is a multiple of the identity and every jump operator of the
    // library is unital (c c^dag = c^dag c: the Paulis), that part is a steady state: rho00 and rho11 of the observed
    // site stay at Tr(rho0) / 2 and only the two odd tiles -- half the state, half of every stencil, shuffle and DMMA --
    // are evolved.  Both conditions are properties of the plan (checked at plan creation): the production initial state
    // of the reference (qubit |+><+|, defects maximally mixed, execute_learning_full.py:151-164) with any configs.py library.
    auto live = [](const int I, const int J) { return !SB || (((I ^ J) & 1) != 0); };
    // PARITY (every Hamiltonian term conserves the parity of the computational-basis index: sigma_z energies,
    // diag (x) diag or flip (x) flip couplings -- all libraries of configs.py): H is block diagonal in the coordinates
    // x' = x ^ (parity(x & 7) << 3), i.e. the top (tile) bit of the kernel's row / column index is the PARITY of the
    // physical index instead of the observed site's bit.  Then G has only diagonal tiles and G v needs HALF the DMMAs;
    // a flip of any site also flips the tile bit of row and column (pure register renaming, as the observed site's
    // flip already was), the lane shuffles are unchanged.
    auto phys = [](const int x) { return PARITY ? x ^ (((x ^ (x >> 1) ^ (x >> 2)) & 1) << 3) : x; };
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Scratch = WarpScratch<TL, CHEB, SB, PARITY>;
    constexpr int NTR = Scratch::NTR;
    Scratch& S = reinterpret_cast<Scratch*>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int T = P.n_times, K = P.n_obs;

    // site -> bit position: the observable's site sits on the top bit (for d = 16 the tile bits I,J), the
    // others follow in order
    const int obs_site = P.obs_site;
    auto pos = [&](int s) { return (s == obs_site) ? N - 1 : (s < obs_site ? N - 2 - s : N - 1 - s); };
    // permuted index -> original (site 0 most significant) index, for reading rho0
    auto orig_index = [&](int x) {
        int o = 0;
#pragma unroll
        for (int s = 0; s < N; ++s) o |= ((x >> pos(s)) & 1) << (N - 1 - s);
        return o;
    };

    // transposition sources (see header): round A serves dest e = g&1, round B dest e = 1-(g&1)
    const int gp = g & 1;
    const int srcA = 4 * (2 * t + gp) + (g >> 1);
    const int srcB = 4 * (2 * t + 1 - gp) + (g >> 1);
    // lanes holding partial-trace elements (row and column agree on every bit except the observed one), at e = g&1:
    // d = 16: (g2,g1,g0) == (t1,t0,e); d = 8: (g1,g0) == (t0,e) and the observed bit is (g2 | t1)
    const bool diag_lane = (TL == 2) ? (t == (g >> 1)) : (((g >> 1) & 1) == (t & 1));

    pdl_wait();
    for (;;) {
        // ---- dynamic model fetch (one atomic per warp)
#ifdef EML_PROFILE_PHASES
        long long phase_t = clock64();
#endif
        unsigned int bidx = 0;
        if (lane == 0) bidx = atomicAdd(counter, 1u);
        bidx = __shfl_sync(0xffffffffu, bidx, 0);
        if (UNI((long long)bidx >= n_models)) { fetch_ws_release(counter, bidx, n_models, binned, lane); break; }
        // (binned: `order` holds the cost-bin lists of the prepare kernel and the ticket is resolved here)
        const long long b = (order == nullptr) ? (long long)bidx : binned ? (long long)order_list_get(counter, order, bidx, lane) : (long long)order[bidx];

        double a_re[TL][KS], a_im[TL][KS];
        if constexpr (CHEB) {
            // ---- the prepare kernel's record of this model: A fragments (coalesced), weights and plan through shared memory
            constexpr int KSE = PARITY ? 2 : KS;
            constexpr int NA = prep_fragments<TL, REAL_H, PARITY>(), REC = prep_record_doubles<TL, REAL_H, PARITY>();
            EML_CHECK(b >= 0 && b < n_models);
            const double* const R = prep + (size_t)b * REC;
#pragma unroll
            for (int I = 0; I < TL; ++I)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    a_im[I][ks] = (ks < KSE) ? R[((0 * TL + I) * KSE + ks) * 32 + lane] : 0.0;
                    a_re[I][ks] = (!REAL_H && ks < KSE) ? R[((1 * TL + I) * KSE + ks) * 32 + lane] : 0.0;
                }
            S.jw[lane] = R[NA * 32 + lane];
            if (lane < 20) S.gdiag[lane] = R[NA * 32 + 32 + lane];      // gdiag[16] and plan[4] are contiguous
            __syncwarp();
            // not evaluable (non-finite parameters or an absurd step count): reported, not attempted
            if (UNI(!(S.plan[0] == S.plan[0]))) {
                if (lane == 0) {
                    if (loss != nullptr) loss[b] = nan("");
                    if (status != nullptr) status[b] = -1;
                }
                __syncwarp();
                continue;
            }
        } else {
            const double nu_ = assemble_generator<TL, false>(S, P, params + b * P.n_params, lane);
            // a non-finite generator (NaN / inf parameters) or an absurd step count must not hang the device
            if (UNI(!(nu_ * (P.times[T - 1] - P.times[0]) < 1e5))) {
                if (lane == 0) {
                    if (loss != nullptr) loss[b] = nan("");
                    if (status != nullptr) status[b] = -1;
                }
                __syncwarp();
                continue;
            }
        }
        [[maybe_unused]] double nu = 0.0;
        if constexpr (!CHEB) nu = S.misc[0];
        // ---- A fragments of G with the permuted k index.  REAL_H: H is real, so G.re = -K/2 is diagonal and is
        //      folded into the diagonal stencil weight; only G.im = -H goes through the DMMA pipe.
        if constexpr (!CHEB) {
#pragma unroll
            for (int I = 0; I < TL; ++I)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int k = 8 * (ks >> 1) + 2 * t + (ks & 1);
                    a_re[I][ks] = REAL_H ? 0.0 : S.G_re[8 * I + g][k];
                    a_im[I][ks] = S.G_im[8 * I + g][k];
                }
        }
        // Stencil weights are stored HALVED: the DMMA accumulators start from J(v)/2, so that
        // X' + X'^dag = G v + v G^dag + J(v)  (J(v) is Hermitian).
        double W0[TL][TL][2];
        constexpr int FE = PARITY ? 2 : 1;      // PARITY: the observed site's column bit depends on e
        double F3[TL][TL][FE];
#pragma unroll
        for (int I = 0; I < TL; ++I)
#pragma unroll
            for (int J = 0; J < TL; ++J) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int row = phys(8 * I + g), col = phys(8 * J + 2 * t + e);     // physical indices
                    const int r3 = (row >> 3) & 1, c3 = (col >> 3) & 1;                 // the observed site's bits (d = 16)
                    double w = S.jw[2 * 8 + (g >> 2) * 2 + (t >> 1)] + S.jw[1 * 8 + ((g >> 1) & 1) * 2 + (t & 1)] +
                               S.jw[0 * 8 + (g & 1) * 2 + e];
                    if (TL == 2) w += S.jw[3 * 8 + r3 * 2 + c3];
                    if constexpr (REAL_H) {
                        if constexpr (CHEB) w += S.gdiag[row] + S.gdiag[col]; else w += S.G_re[row][row] + S.G_re[col][col];
                    }
                    W0[I][J][e] = 0.5 * w;
                    if (e < FE) F3[I][J][e] = (TL == 2) ? 0.5 * S.jw[3 * 8 + 4 + r3 * 2 + c3] : 0.0;
                }
            }
        double f2 = 0.5 * S.jw[2 * 8 + 4 + (g >> 2) * 2 + (t >> 1)];
        double f1 = 0.5 * S.jw[1 * 8 + 4 + ((g >> 1) & 1) * 2 + (t & 1)];
        double f0[2] = {0.5 * S.jw[0 * 8 + 4 + (g & 1) * 2 + 0], 0.5 * S.jw[0 * 8 + 4 + (g & 1) * 2 + 1]};
        // The generator is kept PRE-MULTIPLIED by the current step length h (rescaled in place when h changes), so
        // that u_{j+1} = (h L) u_j needs no per-element scaling; the Taylor term is u_j / j!.
        double h_cur = 1.0;

        // ---- state: acc = rho(t_k) in C layout (rho0 permuted to the kernel's bit order)
        double acc_re[TL][TL][2], acc_im[TL][TL][2];
#pragma unroll
        for (int I = 0; I < TL; ++I)
#pragma unroll
            for (int J = 0; J < TL; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r = orig_index(phys(8 * I + g)), c = orig_index(phys(8 * J + 2 * t + e));
                    acc_re[I][J][e] = P.rho0[2 * (r * D + c)];
                    acc_im[I][J][e] = P.rho0[2 * (r * D + c) + 1];
                }

        const int tset = (target_set != nullptr) ? target_set[b] : 0;
        const double* tgt = (P.targets != nullptr) ? P.targets + (size_t)tset * K * T * 2 : nullptr;
        double my_loss = 0.0;
        int n_apply = 0;
        bool capped = false;

        // partial trace over the non-observed sites of a C-layout operand, reduce-scattered over the warp:
        // afterwards every lane holds the total of value index vi = lane >> 3 (0: rho00.re, 1: rho11.re,
        // 2: rho01.re, 3: rho01.im)
        auto partial_trace = [&](const double (&xr)[TL][TL][2], const double (&xi)[TL][TL][2]) -> double {
            double v0, v1, v2, v3;
            if (TL == 2) {
                v0 = diag_lane ? (gp ? xr[0][0][1] : xr[0][0][0]) : 0.0;
                v1 = diag_lane ? (gp ? xr[TL - 1][TL - 1][1] : xr[TL - 1][TL - 1][0]) : 0.0;
                v2 = diag_lane ? (gp ? xr[0][TL - 1][1] : xr[0][TL - 1][0]) : 0.0;
                v3 = diag_lane ? (gp ? xi[0][TL - 1][1] : xi[0][TL - 1][0]) : 0.0;
            } else {
                const double er = gp ? xr[0][0][1] : xr[0][0][0], ei = gp ? xi[0][0][1] : xi[0][0][0];
                const int a = g >> 2, bb = t >> 1;      // observed bit of the row / of the column
                v0 = (diag_lane && a == 0 && bb == 0) ? er : 0.0;
                v1 = (diag_lane && a == 1 && bb == 1) ? er : 0.0;
                v2 = (diag_lane && a == 0 && bb == 1) ? er : 0.0;
                v3 = (diag_lane && a == 0 && bb == 1) ? ei : 0.0;
            }
            const bool up16 = lane & 16;
            // lanes with bit4 = 0 keep (v0, v1), bit4 = 1 keep (v2, v3)
            double s0 = up16 ? v0 : v2, s1 = up16 ? v1 : v3;
            s0 = __shfl_xor_sync(0xffffffffu, s0, 16);
            s1 = __shfl_xor_sync(0xffffffffu, s1, 16);
            double k0 = (up16 ? v2 : v0) + s0, k1 = (up16 ? v3 : v1) + s1;
            const bool up8 = lane & 8;
            double s = up8 ? k0 : k1;
            s = __shfl_xor_sync(0xffffffffu, s, 8);
            double k = (up8 ? k1 : k0) + s;
            k += __shfl_xor_sync(0xffffffffu, k, 4);
            k += __shfl_xor_sync(0xffffffffu, k, 2);
            k += __shfl_xor_sync(0xffffffffu, k, 1);
            return k;   // value index = 2*bit4 + bit3 = lane >> 3
        };

        // observables and loss of one output time from the reduced density matrix r = (rho00, rho11, Re rho01, Im rho01)
        auto finish_output = [&](const int k, const double (&r)[4]) {
            for (int m = 0; m < K; ++m) {
                const double* O = P.obs + m * 8;
                // Tr(O rho_red) = O00 r00 + O11 r11 + O01 r10 + O10 r01,  r10 = conj(r01)
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

        // emit q outputs starting at time index kbase; y[n] is this lane's accumulator for output
        // p = (lane&7) + 8n of value index lane>>3.  Lane l finalises outputs l, l + 32, ...
        auto emit = [&](int q, int kbase, const double (&y)[NY]) {
          for (int pbase = 0; UNI(pbase < q); pbase += 32) {
            const int p = pbase + lane;
            const int srcl = p & 7;
            double r[4];
#pragma unroll
            for (int vi = 0; vi < 4; ++vi) {
                double val = 0.0;
#pragma unroll
                for (int n = 0; n < NY; ++n) {
                    // the source lane vi*8 + srcl holds output 8n + srcl in y[n]; this lane wants n == p >> 3
                    if ((n >> 2) == (pbase >> 5)) {      // outputs of this 32-block live in y[4*(pbase/32) .. +3]
                        const double got = shfl(y[n], vi * 8 + srcl);
                        if (n == (p >> 3)) val = got;
                    }
                }
                r[vi] = val;
            }
            if (p < q) finish_output(kbase + p, r);
          }
        };

        if constexpr (!CHEB) {
            const double v = partial_trace(acc_re, acc_im);
            double y1st[NY];
#pragma unroll
            for (int n = 0; n < NY; ++n) y1st[n] = v;
            emit(1, 0, y1st);
        }

        // ---- generator application core, shared by both series: given x = (diagonal stencil weight) * v (+ anything
        //      Hermitian / 2), add the remaining site stencils (J(v)/2) and G v on the DMMA pipe -> X'
        auto apply_rest = [&](const double (&v_re)[TL][TL][2], const double (&v_im)[TL][TL][2], double (&x_re)[TL][TL][2], double (&x_im)[TL][TL][2]) {
            constexpr int PX = PARITY ? 1 : 0;     // PARITY: a flip of a lane-bit site also flips the tile (parity) bit
            {
#pragma unroll
                for (int I = 0; I < TL; ++I)
#pragma unroll
                    for (int J = 0; J < TL; ++J)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (!live(I, J)) continue;
                            if (TL == 2) {
                                x_re[I][J][e] = fma(F3[I][J][e & (FE - 1)], v_re[(I ^ 1) & (TL - 1)][(J ^ 1) & (TL - 1)][e], x_re[I][J][e]);
                                x_im[I][J][e] = fma(F3[I][J][e & (FE - 1)], v_im[(I ^ 1) & (TL - 1)][(J ^ 1) & (TL - 1)][e], x_im[I][J][e]);
                            }
                        }
            }
            if constexpr (TL == 2 && EML_TL2_SMEM_STENCIL) {
                // the term is published row-major (ld 24: conflict-free 16-byte accesses over (g, t)) in the buffer the
                // transposition uses later; every lane-bit site then costs one LDS.128 per tile and plane
                constexpr int LD = 24;
#pragma unroll
                for (int I = 0; I < TL; ++I)
#pragma unroll
                    for (int J = 0; J < TL; ++J) {
                        *reinterpret_cast<double2*>(&S.xt_re[(8 * I + g) * LD + 8 * J + 2 * t]) = make_double2(v_re[I][J][0], v_re[I][J][1]);
                        *reinterpret_cast<double2*>(&S.xt_im[(8 * I + g) * LD + 8 * J + 2 * t]) = make_double2(v_im[I][J][0], v_im[I][J][1]);
                    }
                __syncwarp();
#pragma unroll
                for (int I = 0; I < TL; ++I)
#pragma unroll
                    for (int J = 0; J < TL; ++J) {
                        const int ro = 8 * (I ^ PX), co = 8 * (J ^ PX);
                        const int o2 = (ro + (g ^ 4)) * LD + co + 2 * (t ^ 2);
                        const int o1 = (ro + (g ^ 2)) * LD + co + 2 * (t ^ 1);
                        const int o0 = (ro + (g ^ 1)) * LD + co + 2 * t;
                        const double2 r2 = *reinterpret_cast<const double2*>(&S.xt_re[o2]), i2 = *reinterpret_cast<const double2*>(&S.xt_im[o2]);
                        const double2 r1 = *reinterpret_cast<const double2*>(&S.xt_re[o1]), i1 = *reinterpret_cast<const double2*>(&S.xt_im[o1]);
                        const double2 r0 = *reinterpret_cast<const double2*>(&S.xt_re[o0]), i0 = *reinterpret_cast<const double2*>(&S.xt_im[o0]);
                        x_re[I][J][0] = fma(f2, r2.x, x_re[I][J][0]); x_re[I][J][1] = fma(f2, r2.y, x_re[I][J][1]);
                        x_im[I][J][0] = fma(f2, i2.x, x_im[I][J][0]); x_im[I][J][1] = fma(f2, i2.y, x_im[I][J][1]);
                        x_re[I][J][0] = fma(f1, r1.x, x_re[I][J][0]); x_re[I][J][1] = fma(f1, r1.y, x_re[I][J][1]);
                        x_im[I][J][0] = fma(f1, i1.x, x_im[I][J][0]); x_im[I][J][1] = fma(f1, i1.y, x_im[I][J][1]);
                        x_re[I][J][0] = fma(f0[0], r0.y, x_re[I][J][0]); x_re[I][J][1] = fma(f0[1], r0.x, x_re[I][J][1]);   // e ^ 1
                        x_im[I][J][0] = fma(f0[0], i0.y, x_im[I][J][0]); x_im[I][J][1] = fma(f0[1], i0.x, x_im[I][J][1]);
                    }
                __syncwarp();      // the buffer is reused by the transposition
            } else {
            {
#pragma unroll
                for (int I = 0; I < TL; ++I)
#pragma unroll
                    for (int J = 0; J < TL; ++J)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (!live(I, J)) continue;
                            x_re[I][J][e] = fma(f2, __shfl_xor_sync(0xffffffffu, v_re[I ^ PX][J ^ PX][e], 18), x_re[I][J][e]);
                            x_im[I][J][e] = fma(f2, __shfl_xor_sync(0xffffffffu, v_im[I ^ PX][J ^ PX][e], 18), x_im[I][J][e]);
                        }
            }
            {
#pragma unroll
                for (int I = 0; I < TL; ++I)
#pragma unroll
                    for (int J = 0; J < TL; ++J)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (!live(I, J)) continue;
                            x_re[I][J][e] = fma(f1, __shfl_xor_sync(0xffffffffu, v_re[I ^ PX][J ^ PX][e], 9), x_re[I][J][e]);
                            x_im[I][J][e] = fma(f1, __shfl_xor_sync(0xffffffffu, v_im[I ^ PX][J ^ PX][e], 9), x_im[I][J][e]);
                        }
            }
            {
#pragma unroll
                for (int I = 0; I < TL; ++I)
#pragma unroll
                    for (int J = 0; J < TL; ++J)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (!live(I, J)) continue;
                            x_re[I][J][e] = fma(f0[e], __shfl_xor_sync(0xffffffffu, v_re[I ^ PX][J ^ PX][e ^ 1], 4), x_re[I][J][e]);
                            x_im[I][J][e] = fma(f0[e], __shfl_xor_sync(0xffffffffu, v_im[I ^ PX][J ^ PX][e ^ 1], 4), x_im[I][J][e]);
                        }
            }
            }
            // ---- X' = J(v)/2 + G v on the DMMA pipe (B fragments = conj of own registers):
            //      X.re += a.re*b_re + a.im*b_im ; X.im += -a.re*b_im + a.im*b_re   (B = conj(b))
            if constexpr (PARITY) {
                // only the diagonal tiles of G: X[I][J] = G[I][I] v[I][J], B = conj of the own registers v[J][I]
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int J = 0; J < TL; ++J)
#pragma unroll
                        for (int I = 0; I < TL; ++I) {
                            if (!live(I, J)) continue;
                            const double b_re = v_re[J][I][ks], b_im = v_im[J][I][ks];
                            if (!REAL_H) {
                                dmma(x_re[I][J][0], x_re[I][J][1], a_re[I][ks], b_re);
                                dmma(x_im[I][J][0], x_im[I][J][1], a_re[I][ks], neg(b_im));
                            }
                            dmma(x_re[I][J][0], x_re[I][J][1], a_im[I][ks], b_im);
                            dmma(x_im[I][J][0], x_im[I][J][1], a_im[I][ks], b_re);
                        }
            } else {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                if (!REAL_H) {
#pragma unroll
                    for (int J = 0; J < TL; ++J) {
                        const double b_re = v_re[J][ks >> 1][ks & 1];
                        const double nb_im = neg(v_im[J][ks >> 1][ks & 1]);
#pragma unroll
                        for (int I = 0; I < TL; ++I) {
                            dmma(x_re[I][J][0], x_re[I][J][1], a_re[I][ks], b_re);
                            dmma(x_im[I][J][0], x_im[I][J][1], a_re[I][ks], nb_im);
                        }
                    }
                }
#pragma unroll
                for (int J = 0; J < TL; ++J) {
                    const double b_re = v_re[J][ks >> 1][ks & 1];
                    const double b_im = v_im[J][ks >> 1][ks & 1];
#pragma unroll
                    for (int I = 0; I < TL; ++I) {
                        dmma(x_re[I][J][0], x_re[I][J][1], a_im[I][ks], b_im);
                        dmma(x_im[I][J][0], x_im[I][J][1], a_im[I][ks], b_re);
                    }
                }
            }
            }
        };
        // w = X' + X'^dag (the conjugate transpose gathered with shuffles, d = 8: through shared memory)
        auto symmetrise = [&](const double (&x_re)[TL][TL][2], const double (&x_im)[TL][TL][2], double (&w_re)[TL][TL][2], double (&w_im)[TL][TL][2]) {
            if constexpr (TL == 2 && !EML_TL2_SMEM_TRANSPOSE) {     // d = 16: shuffles (measured faster than the shared-memory route)
#pragma unroll
            for (int I = 0; I < TL; ++I)
#pragma unroll
                for (int J = 0; J < TL; ++J) {
                    if (!live(I, J)) continue;
                    // X^dag[I][J][e] = conj(X[8J+2t+e][8I+g]) gathered from tile (J,I) with shuffles
                    const double sAr = gp ? x_re[J][I][1] : x_re[J][I][0], sBr = gp ? x_re[J][I][0] : x_re[J][I][1];
                    const double sAi = gp ? x_im[J][I][1] : x_im[J][I][0], sBi = gp ? x_im[J][I][0] : x_im[J][I][1];
                    const double rAr = shfl(sAr, srcA), rBr = shfl(sBr, srcB);
                    const double rAi = shfl(sAi, srcA), rBi = shfl(sBi, srcB);
                    const double t_re[2] = {gp ? rBr : rAr, gp ? rAr : rBr};
                    const double t_im[2] = {gp ? rBi : rAi, gp ? rAi : rBi};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double wr = x_re[I][J][e] + t_re[e];
                        const double wi = x_im[I][J][e] - t_im[e];
                        w_re[I][J][e] = wr; w_im[I][J][e] = wi;
                    }
                }
            } else {                     // d = 8: issue-bound, the shared-memory route saves ~90 instructions (+5 %)
            // X'^dag through shared memory: element (R, C) of X' is stored at C*16 + (R ^ m(C)) with the swizzle
            // m(C) = 4*bit1(C) + 8*(bit2(C) ^ bit0(C)), which makes both the STS.64 of the accumulator
            // fragments and the LDS.128 of the transposed pairs bank-conflict free.
#pragma unroll
            for (int I = 0; I < TL; ++I)
#pragma unroll
                for (int J = 0; J < TL; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (!live(I, J)) continue;
                        const int Cc = 8 * J + 2 * t + e;
                        const int m = 4 * (t & 1) + 8 * (((t >> 1) & 1) ^ e);
                        S.xt_re[Cc * 16 + ((8 * I + g) ^ m)] = x_re[I][J][e];
                        S.xt_im[Cc * 16 + ((8 * I + g) ^ m)] = x_im[I][J][e];
                    }
            __syncwarp();
#pragma unroll
            for (int I = 0; I < TL; ++I)
#pragma unroll
                for (int J = 0; J < TL; ++J) {
                    if (!live(I, J)) continue;
                    const int m = 4 * ((g >> 1) & 1) + 8 * (((g >> 2) & 1) ^ (g & 1));
                    const int o = (8 * I + g) * 16 + ((8 * J + 2 * t) ^ m);
                    const double2 tr_ = *reinterpret_cast<const double2*>(&S.xt_re[o]);
                    const double2 ti_ = *reinterpret_cast<const double2*>(&S.xt_im[o]);
                    const double wr0 = x_re[I][J][0] + tr_.x, wr1 = x_re[I][J][1] + tr_.y;
                    const double wi0 = x_im[I][J][0] - ti_.x, wi1 = x_im[I][J][1] - ti_.y;
                    w_re[I][J][0] = wr0; w_re[I][J][1] = wr1; w_im[I][J][0] = wi0; w_im[I][J][1] = wi1;
                }
            __syncwarp();
            }
        };

        if constexpr (CHEB) {
        // =====================================================================================================
        // Chebyshev series of the propagator (Jacobi-Anger):  with  L = L_s - ac  and  At = L_s / R,
        //     exp(tau L) rho = e^{-ac tau} [ J_0(tau R) phi_0 + 2 sum_{k>=1} J_k(tau R) phi_k ],
        //     phi_0 = rho,  phi_1 = At rho,  phi_{k+1} = 2 At phi_k + phi_{k-1}
        // (phi_k = i^k T_k(At / i): a REAL three-term recurrence).  The phi_k do not depend on tau, so ONE recurrence
        // serves every output time of a macro step; only the 2x2 partial trace of each phi_k is kept (shared
        // memory) and each output is a Bessel-weighted sum of them, the weights generated per output by Miller's
        // backward recurrence.  There is no e^theta cancellation hump as in the Taylor series, so a macro step can
        // span hundreds of output times: about  tau (A + B) + 9.25 (tau (A + B))^(1/3) + 5  generator applications
        // instead of ~4 per unit of the norm bound.  [-iR, iR] must enclose the imaginary extent of the numerical
        // range: Gershgorin on H for the commutator, per-site Gershgorin for the dissipator (real range [dlo, dhi],
        // centred by the shift ac; antisymmetric part added to the imaginary extent); the series is truncated for
        // the confocal ellipse (semi-axes A along the imaginary axis, B real) through the corner of that rectangle.
        PHASE_MARK(0);      // record of the prepare kernel, stencil weights, initial state
        // the plan (R, S, ac, tau_max) and Tr(rho0)/2 stay in shared memory and are re-read where they are needed, so that
        // none of them occupies registers across the recurrence loop
#define Rb (S.plan[0])
#define Sb (S.plan[1])
#define ac (S.plan[2])
#define tau_max (S.plan[3])
        {   // generator -> 2 (L + ac) / R  (stencil weights are stored halved)
            const double sc = 2.0 / Rb;
            const double sw = SB ? 2.0 * sc : sc;      // the sector variant applies J(v) in one piece: weights unhalved
#pragma unroll
            for (int I = 0; I < TL; ++I) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) { a_re[I][ks] *= sc; a_im[I][ks] *= sc; }
#pragma unroll
                for (int J = 0; J < TL; ++J) {
#pragma unroll
                    for (int e = 0; e < FE; ++e) F3[I][J][e] *= sw;
                    W0[I][J][0] = (W0[I][J][0] + 0.5 * ac) * sw; W0[I][J][1] = (W0[I][J][1] + 0.5 * ac) * sw;
                }
            }
            f2 *= sw; f1 *= sw; f0[0] *= sw; f0[1] *= sw;
        }
        // 2x2 partial trace of a term straight into shared memory (dst[0..3] = rho00, rho11, Re rho01, Im rho01).  d = 16:
        // only the 8 lanes 4g + (g >> 1) hold partial-trace elements; a 3-stage reduce-scatter among exactly those lanes
        // (4 double shuffles) leaves value 2*bit2(g) + bit1(g) in the lanes with g even.
        const int dl4 = 4 * (g ^ 4) + ((g ^ 4) >> 1), dl2 = 4 * (g ^ 2) + ((g ^ 2) >> 1), dl1 = 4 * (g ^ 1) + ((g ^ 1) >> 1);
        // (the table entry is addressed by term index so that the store stays a predicated STS with a shared-window address)
        // (d = 8: the four lanes of observed-bit pair (a, bb) are 16 a + 2 bb + 9 g1 + 4 g0; after the two-stage butterfly all of
        //  them hold the totals: lane 0 stores rho00, lane 18 rho11, lane 2 Re rho01, lane 6 Im rho01)
        const bool tr_lane = SB ? (diag_lane && (g & 3) == 0) : (TL == 2) ? (diag_lane && !(g & 1))
                                : (lane == 0 || lane == 18 || lane == 2 || lane == 6);
        const int tr_off = SB ? (g >> 2) : (TL == 2) ? (g >> 1) : (lane == 0 ? 0 : lane == 18 ? 1 : lane == 2 ? 2 : 3);
        const unsigned tr_base = (unsigned)__cvta_generic_to_shared(&S.tr[tr_off]);
        // predicated shared store of one table entry (kept out of a branch so that the recurrence loop stays one basic block)
        auto tr_put = [&](const int kterm, const double val) {
            EML_CHECK(kterm >= 0 && kterm <= Scratch::KCAP);
            asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.shared.f64 [%0], %1; }"
                         :: "r"(tr_base + (unsigned)(kterm * NTR * 8)), "d"(val), "r"((int)tr_lane) : "memory");
        };
        // SB: only rho01 of the observed site moves.  It lives in tile (p, 1 ^ p), p = parity of the (equal) low bits of row
        // and column; tile (1,0) is the conjugate transpose of tile (0,1) and a partial-trace element sits on the diagonal
        // of its tile, so tile (0,1) alone supplies it (conjugated where p = 1); dst[0..1]
        auto trace_store_sb = [&](const double (&mr)[2], const double (&mi)[2], const int kterm) {
            const bool pl = (g ^ (g >> 1) ^ (g >> 2)) & 1;
            const double a2 = gp ? mr[1] : mr[0], a3 = gp ? mi[1] : mi[0];
            const double v2 = a2, v3 = pl ? neg(a3) : a3;
            const bool hb = g & 4;
            double kk = (hb ? v3 : v2) + shfl(hb ? v2 : v3, dl4);
            kk += shfl(kk, dl2);
            kk += shfl(kk, dl1);
            tr_put(kterm, kk);
        };
        auto trace_store = [&](const double (&xr)[TL][TL][2], const double (&xi)[TL][TL][2], const int kterm) {
            if constexpr (SB) {
                trace_store_sb(xr[0][1], xi[0][1], kterm);
            } else if constexpr (TL == 2) {
                double v0 = gp ? xr[0][0][1] : xr[0][0][0], v1 = gp ? xr[1][1][1] : xr[1][1][0];
                double v2 = gp ? xr[0][1][1] : xr[0][1][0], v3 = gp ? xi[0][1][1] : xi[0][1][0];
                if constexpr (PARITY) {
                    // the observed site's bits are (I ^ p, J ^ p), p = parity of the (equal) low bits of row and column
                    const bool pl = (g ^ (g >> 1) ^ (g >> 2)) & 1;
                    // (tile (1,0) is the conjugate transpose of tile (0,1); a partial-trace element sits on its tile's diagonal)
                    const double a0 = v0, a1 = v1;
                    v0 = pl ? a1 : a0; v1 = pl ? a0 : a1; v3 = pl ? neg(v3) : v3;
                }
                const bool hb = g & 4, mb = g & 2;
                const double r0 = shfl(hb ? v0 : v2, dl4), r1 = shfl(hb ? v1 : v3, dl4);
                const double k0 = (hb ? v2 : v0) + r0, k1 = (hb ? v3 : v1) + r1;
                const double kk = (mb ? k1 : k0) + shfl(mb ? k0 : k1, dl2);
                const double tot = kk + shfl(kk, dl1);
                tr_put(kterm, tot);
            } else {
                // d = 8: a partial-trace element (row and column agree on the two non-observed bits) sits at e = g & 1 of the
                // lanes with t0 = g1; flipping g0 (lane ^ 4) or g1 together with t0 (lane ^ 9) stays among them
                const double er = gp ? xr[0][0][1] : xr[0][0][0], ei = gp ? xi[0][0][1] : xi[0][0][0];
                double sr = er + __shfl_xor_sync(0xffffffffu, er, 4), si = ei + __shfl_xor_sync(0xffffffffu, ei, 4);
                sr += __shfl_xor_sync(0xffffffffu, sr, 9);
                si += __shfl_xor_sync(0xffffffffu, si, 9);
                tr_put(kterm, lane == 6 ? si : sr);
            }
        };
        double* const cJ = S.cj;     // coefficient table of the step end

        // SB: rho00 = rho11 = Tr(rho0) / 2 of the observed site never move
        if constexpr (SB) {
            double h = (lane < D) ? P.rho0[2 * (lane * D + lane)] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
            if (lane == 0) S.half_tr[0] = 0.5 * h;
            __syncwarp();
        }
#define half_tr (S.half_tr[0])
        {   // the first output time carries the initial state
            trace_store(acc_re, acc_im, 0);
            __syncwarp();
            const double r[4] = {SB ? half_tr : S.tr[0], SB ? half_tr : S.tr[1], S.tr[NTR - 2], S.tr[NTR - 1]};
            if (lane == 0) finish_output(0, r);
            __syncwarp();
        }
        int kidx = 0;
        while (UNI(kidx < T - 1)) {
            const double t0 = P.times[kidx];
            int q = 0;      // output times within tau_max of t0: 32 candidates per ballot
            for (;;) {
                const int cand = kidx + q + 1 + lane;
                const unsigned in = __ballot_sync(0xffffffffu, cand <= T - 1 && P.times[min(cand, T - 1)] - t0 <= tau_max);
                const int run = __ffs(~in) - 1;       // leading run of accepted candidates (-1: all 32)
                if (run >= 0) { q += run; break; }
                q += 32;
            }
            int nsub = 1;
            if (UNI(q == 0)) { nsub = (int)ceil((P.times[kidx + 1] - t0) / tau_max); q = 1; }
            const double tau_end = (P.times[kidx + q] - t0) / nsub;
            const bool final_step = (kidx + q == T - 1);
            if (UNI(!(tau_end * Rb >= 1e-40))) {     // repeated time points: the state does not move
                __syncwarp();
                trace_store(acc_re, acc_im, 0);
                __syncwarp();
                const double r[4] = {SB ? half_tr : S.tr[0], SB ? half_tr : S.tr[1], S.tr[NTR - 2], S.tr[NTR - 1]};
                for (int pbase = 0; UNI(pbase < q); pbase += 32)
                    if (pbase + lane < q) finish_output(kidx + 1 + pbase + lane, r);
                __syncwarp();
                kidx += q;
                continue;
            }
            int Kc = cheb_terms(tau_end * Sb);
            if (Kc > Scratch::KCAP) Kc = Scratch::KCAP;
            for (int sub = 0; UNI(sub < nsub); ++sub) {
                const bool last = (sub == nsub - 1);
                const bool need_acc = UNI(!(final_step && last));     // the state at the step end is needed
                PHASE_MARK(1);      // plan, rescaling, first output, step scan
                double cscale = 0.0;
                if (need_acc) {
                    // J_k(tau_end R), k = 0..Kc: Miller's backward recurrence, every lane redundantly, normalised by
                    // J_0 + 2 sum J_2k = 1
                    __syncwarp();
                    const double tz = 2.0 / (tau_end * Rb);
                    double jn = 1.0, jn1 = 0.0, ssum = 0.0;
                    for (int k = Kc; k >= 1; --k) {
                        EML_CHECK(k <= Scratch::KCAP + 1);
                        if ((k & 31) == lane) cJ[k] = jn;
                        if (!(k & 1)) ssum += jn;
                        const double jm = fma((double)k * tz, jn, -jn1);
                        jn1 = jn; jn = jm;
                    }
                    if (lane == 0) cJ[0] = jn;
                    cscale = exp(-ac * tau_end) / (2.0 * ssum + jn);
                    __syncwarp();
                }
                if constexpr (PARITY) {
                // ---- parity coordinates: rho is Hermitian, so the odd tile (1,0) is the conjugate transpose of tile (0,1) and
                //      the even tiles (0,0), (1,1) are Hermitian themselves.  The STATE is  m = tile (0,1)  and, unless the
                //      even sector is steady (SB),  d[I] = tile (I,I);  n = m^dag is kept only as an operand.
                //      Odd tile:  w01 = J(v)01 + G00 m + m G11^dag + (previous term).  Every site flip maps tile (0,1) <-> (1,0),
                //      so the flip stencils read n; G00 m takes B = conj(n) from the lane's own registers, m G11^dag takes
                //      A = m from the lane's own registers (k index permuted as for G's A fragments, whose registers are also
                //      the B fragments of G11^dag): no X' + X'^dag pass, the new n = w01^dag goes through shared memory.
                //      Even tiles:  w[I] = X' + X'^dag,  X' = J(v)[I][I]/2 + G[I][I] d[I]  (B = conj(d[I]) = own registers; the flips
                //      read d[I ^ 1]); X'^dag comes back through the same shared-memory transposition.
                double m_re[2] = {acc_re[0][1][0], acc_re[0][1][1]}, m_im[2] = {acc_im[0][1][0], acc_im[0][1][1]};
                double n_re[2], nn_im[2], p_re[2] = {0.0, 0.0}, p_im[2] = {0.0, 0.0};      // n = n_re - i nn_im
                acc_re[0][1][0] = acc_re[0][1][1] = acc_im[0][1][0] = acc_im[0][1][1] = 0.0;
                [[maybe_unused]] double d_re[2][2], d_im[2][2], q_re[2][2], q_im[2][2];     // [I][e]; q = previous term, halved
                if constexpr (!SB) {
#pragma unroll
                    for (int I = 0; I < 2; ++I)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            d_re[I][e] = acc_re[I][I][e]; d_im[I][e] = acc_im[I][I][e];
                            q_re[I][e] = q_im[I][e] = 0.0;
                            acc_re[I][I][e] = acc_im[I][I][e] = 0.0;
                        }
                }
                // transposition tile: element (R, C) of a plane at 2 pi + (R & 1), pi = 8 h + (a & 1) + 2 ((C >> 1) ^ h),
                // a = R >> 1, h = 2 (C & 1) + (a >> 1): the STS.64 of the accumulator fragments (R = g, C = 2t + e) and the
                // LDS.128 of the transposed pairs (R = 2t, 2t + 1; C = g) are both bank-conflict free.  Planes re | im of 64
                // doubles per tile (odd tile, then the two even tiles), two buffers (term k uses buffer k & 1, so one
                // __syncwarp per application suffices).
                constexpr int XBUF = SB ? 128 : 384;
                constexpr int NBUF = SB ? 2 : PAR_XT_BUFFERS;
                auto xt_addr = [](const int R, const int C) {
                    const int a = R >> 1, h = 2 * (C & 1) + (a >> 1);
                    return 2 * (8 * h + (a & 1) + 2 * ((C >> 1) ^ h)) + (R & 1);
                };
                double* const xs0 = S.xt_re + xt_addr(g, 2 * t), * const xs1 = S.xt_re + xt_addr(g, 2 * t + 1);
                const double* const xl = S.xt_re + xt_addr(2 * t, g);
                EML_CHECK(xt_addr(g, 2 * t + 1) < 64 && xt_addr(2 * t, g) + 1 < 64 && (2 * XBUF) * sizeof(double) <= sizeof(S.xt_re));
                auto transpose_conj = [&](const int buf) {      // n = m^dag
                    EML_CHECK(buf == 0 || buf == 1);
                    xs0[buf * XBUF] = m_re[0]; xs1[buf * XBUF] = m_re[1];
                    xs0[buf * XBUF + 64] = m_im[0]; xs1[buf * XBUF + 64] = m_im[1];
                    __syncwarp();
                    const double2 cr = *reinterpret_cast<const double2*>(xl + buf * XBUF);
                    const double2 ci = *reinterpret_cast<const double2*>(xl + buf * XBUF + 64);
                    n_re[0] = cr.x; n_re[1] = cr.y; nn_im[0] = ci.x; nn_im[1] = ci.y;
                };
                __syncwarp();
                transpose_conj(NBUF - 1);
                if constexpr (NBUF == 1) __syncwarp();
                // flip stencils of the lane-bit sites and of the observed site onto (ar, ai): sources (sr, -si), weights F, f2, f1, f0
                auto flips = [&](double (&ar)[2], double (&ai)[2], const double (&sr)[2], const double (&si)[2], const double (&F)[FE],
                                 const bool neg_im) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        ar[e] = fma(F[e], sr[e], ar[e]);
                        ai[e] = fma(F[e], neg_im ? -si[e] : si[e], ai[e]);
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double hr = __shfl_xor_sync(0xffffffffu, sr[e], 18), hi = __shfl_xor_sync(0xffffffffu, si[e], 18);
                        ar[e] = fma(f2, hr, ar[e]); ai[e] = fma(f2, neg_im ? -hi : hi, ai[e]);
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double hr = __shfl_xor_sync(0xffffffffu, sr[e], 9), hi = __shfl_xor_sync(0xffffffffu, si[e], 9);
                        ar[e] = fma(f1, hr, ar[e]); ai[e] = fma(f1, neg_im ? -hi : hi, ai[e]);
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double hr = __shfl_xor_sync(0xffffffffu, sr[e ^ 1], 4), hi = __shfl_xor_sync(0xffffffffu, si[e ^ 1], 4);
                        ar[e] = fma(f0[e], hr, ar[e]); ai[e] = fma(f0[e], neg_im ? -hi : hi, ai[e]);
                    }
                };
                auto trace_put = [&](const int k) {
                    if constexpr (SB) {
                        trace_store_sb(m_re, m_im, k);
                    } else {
                        // rho00, rho11 from the even tiles, rho01 from tile (0,1) (conjugated where the parity of the low bits is odd)
                        const bool pl = (g ^ (g >> 1) ^ (g >> 2)) & 1;
                        const double a0 = gp ? d_re[0][1] : d_re[0][0], a1 = gp ? d_re[1][1] : d_re[1][0];
                        const double v0 = pl ? a1 : a0, v1 = pl ? a0 : a1;
                        const double v2 = gp ? m_re[1] : m_re[0], a3 = gp ? m_im[1] : m_im[0];
                        const double v3 = pl ? neg(a3) : a3;
                        const bool hb = g & 4, mb = g & 2;
                        const double r0 = shfl(hb ? v0 : v2, dl4), r1 = shfl(hb ? v1 : v3, dl4);
                        const double k0 = (hb ? v2 : v0) + r0, k1 = (hb ? v3 : v1) + r1;
                        const double kk = (mb ? k1 : k0) + shfl(mb ? k0 : k1, dl2);
                        tr_put(k, kk + shfl(kk, dl1));
                    }
                };
                auto accumulate = [&](const int k) {
                    const double c = cJ[k] * cscale;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        acc_re[0][1][e] = fma(c, m_re[e], acc_re[0][1][e]);
                        acc_im[0][1][e] = fma(c, m_im[e], acc_im[0][1][e]);
                        if constexpr (!SB) {
#pragma unroll
                            for (int I = 0; I < 2; ++I) {
                                acc_re[I][I][e] = fma(c, d_re[I][e], acc_re[I][I][e]);
                                acc_im[I][I][e] = fma(c, d_im[I][e], acc_im[I][I][e]);
                            }
                        }
                    }
                };
                auto application = [&](const int k, auto na_tag, auto first_tag) {
                    constexpr bool NA = decltype(na_tag)::value, FIRST = decltype(first_tag)::value;
                    trace_put(k);
                    if constexpr (NA) accumulate(k);
                    // ---- odd tile
                    double x_re[2], x_im[2];
                    if constexpr (SB) {      // weights unhalved: the stencils accumulate straight into x
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            x_re[e] = fma(W0[0][1][e], m_re[e], p_re[e]);
                            x_im[e] = fma(W0[0][1][e], m_im[e], p_im[e]);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 2; ++e) { x_re[e] = p_re[e]; x_im[e] = p_im[e]; }
                    }
                    // m G11^dag: A = m (own registers), B = conj of G11's A fragment -- independent of n, issued first
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        if (!REAL_H) {
                            dmma(x_re[0], x_re[1], m_re[ks], a_re[1][ks]);
                            dmma(x_im[0], x_im[1], m_im[ks], a_re[1][ks]);
                        }
                        dmma(x_re[0], x_re[1], m_im[ks], a_im[1][ks]);
                        dmma(x_im[0], x_im[1], m_re[ks], -a_im[1][ks]);
                    }
                    if constexpr (SB) {
                        flips(x_re, x_im, n_re, nn_im, F3[0][1], true);
                    } else {                 // weights halved (the even tiles need them so): J(v)01 = 2 s
                        double s_re[2], s_im[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) { s_re[e] = W0[0][1][e] * m_re[e]; s_im[e] = W0[0][1][e] * m_im[e]; }
                        flips(s_re, s_im, n_re, nn_im, F3[0][1], true);
#pragma unroll
                        for (int e = 0; e < 2; ++e) { x_re[e] = fma(2.0, s_re[e], x_re[e]); x_im[e] = fma(2.0, s_im[e], x_im[e]); }
                    }
                    // G00 m: B = conj(n) from the lane's own registers
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        if (!REAL_H) {
                            dmma(x_re[0], x_re[1], a_re[0][ks], n_re[ks]);
                            dmma(x_im[0], x_im[1], a_re[0][ks], nn_im[ks]);
                        }
                        dmma(x_re[0], x_re[1], a_im[0][ks], -nn_im[ks]);
                        dmma(x_im[0], x_im[1], a_im[0][ks], n_re[ks]);
                    }
                    // ---- even tiles
                    [[maybe_unused]] double y_re[2][2], y_im[2][2];
                    if constexpr (!SB) {
#pragma unroll
                        for (int I = 0; I < 2; ++I) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                y_re[I][e] = fma(W0[I][I][e], d_re[I][e], q_re[I][e]);
                                y_im[I][e] = fma(W0[I][I][e], d_im[I][e], q_im[I][e]);
                            }
                            flips(y_re[I], y_im[I], d_re[I ^ 1], d_im[I ^ 1], F3[I][I], false);
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                if (!REAL_H) {
                                    dmma(y_re[I][0], y_re[I][1], a_re[I][ks], d_re[I][ks]);
                                    dmma(y_im[I][0], y_im[I][1], a_re[I][ks], -d_im[I][ks]);
                                }
                                dmma(y_re[I][0], y_re[I][1], a_im[I][ks], d_im[I][ks]);
                                dmma(y_im[I][0], y_im[I][1], a_im[I][ks], d_re[I][ks]);
                            }
                        }
                    }
                    // T_2 = (2 At) T_1 + 2 T_0, afterwards T_{k+1} = (2 At) T_k + T_{k-1}  (q is kept halved: X' + X'^dag doubles it)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        p_re[e] = FIRST ? 2.0 * m_re[e] : m_re[e]; p_im[e] = FIRST ? 2.0 * m_im[e] : m_im[e];
                        m_re[e] = x_re[e]; m_im[e] = x_im[e];
                    }
                    const int buf = (NBUF == 2) ? (k & 1) : 0;
                    if constexpr (!SB) {
#pragma unroll
                        for (int I = 0; I < 2; ++I) {
                            xs0[buf * XBUF + 128 * (1 + I)] = y_re[I][0]; xs1[buf * XBUF + 128 * (1 + I)] = y_re[I][1];
                            xs0[buf * XBUF + 128 * (1 + I) + 64] = y_im[I][0]; xs1[buf * XBUF + 128 * (1 + I) + 64] = y_im[I][1];
                        }
                    }
                    transpose_conj(buf);
                    if constexpr (!SB) {
#pragma unroll
                        for (int I = 0; I < 2; ++I) {
                            const double2 tr_ = *reinterpret_cast<const double2*>(xl + buf * XBUF + 128 * (1 + I));
                            const double2 ti_ = *reinterpret_cast<const double2*>(xl + buf * XBUF + 128 * (1 + I) + 64);
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                q_re[I][e] = FIRST ? d_re[I][e] : 0.5 * d_re[I][e]; q_im[I][e] = FIRST ? d_im[I][e] : 0.5 * d_im[I][e];
                            }
                            d_re[I][0] = y_re[I][0] + tr_.x; d_re[I][1] = y_re[I][1] + tr_.y;
                            d_im[I][0] = y_im[I][0] - ti_.x; d_im[I][1] = y_im[I][1] - ti_.y;
                        }
                    }
                    if constexpr (NBUF == 1) __syncwarp();      // the single buffer is rewritten by the next application
                };
                auto run_terms_par = [&](auto na_tag) {
                    application(0, na_tag, std::true_type{});
#pragma unroll 2
                    for (int k = 1; UNI(k < Kc); ++k) application(k, na_tag, std::false_type{});
                    trace_put(Kc);          // the last term
                    if constexpr (decltype(na_tag)::value) accumulate(Kc);
                };
                if (need_acc) run_terms_par(std::true_type{}); else run_terms_par(std::false_type{});
                {   // a non-finite or absurd last term means the enclosure failed: report instead of returning garbage
                    double chk = fabs(m_re[0]) + fabs(m_im[0]) + fabs(m_re[1]);
                    if constexpr (!SB) chk += fabs(d_re[0][0]) + fabs(d_re[1][1]);
                    if (!UNI(chk < 1e300)) capped = true;
                }
                } else {
                    double v_re[TL][TL][2], v_im[TL][TL][2], ph_re[TL][TL][2], ph_im[TL][TL][2];
#pragma unroll
                    for (int I = 0; I < TL; ++I)
#pragma unroll
                        for (int J = 0; J < TL; ++J)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                if (!live(I, J)) continue;
                                v_re[I][J][e] = acc_re[I][J][e]; v_im[I][J][e] = acc_im[I][J][e];
                                ph_re[I][J][e] = 0.0; ph_im[I][J][e] = 0.0;
                                acc_re[I][J][e] = 0.0; acc_im[I][J][e] = 0.0;
                            }
                    // Iteration 0 takes v = phi_0 with nothing added and gives 2 phi_1; from then on the stored input is halved
                    // (exponent decrement on the integer pipe) to supply phi_{k-1} / 2 ... so every term k >= 1 is carried
                    // DOUBLED, which is exactly the factor 2 of the series: the weights are plain J_k for all k.
                    auto run_terms = [&](auto na_tag) {
                        constexpr bool NA = decltype(na_tag)::value;     // accumulate the state at the step end
                        int hsub = 0;
                        [[maybe_unused]] double hmul = 1.0;
#pragma unroll CHEB_UNROLL
                        for (int k = 0; UNI(k < Kc); ++k) {
                            trace_store(v_re, v_im, k);       // partial trace of term k (overlaps the DMMAs below)
                            if constexpr (NA) {
                                const double c = cJ[k] * cscale;
#pragma unroll
                                for (int I = 0; I < TL; ++I)
#pragma unroll
                                    for (int J = 0; J < TL; ++J)
#pragma unroll
                                        for (int e = 0; e < 2; ++e) {
                                            if (!live(I, J)) continue;
                                            acc_re[I][J][e] = fma(c, v_re[I][J][e], acc_re[I][J][e]);
                                            acc_im[I][J][e] = fma(c, v_im[I][J][e], acc_im[I][J][e]);
                                        }
                            }
                            double x_re[TL][TL][2], x_im[TL][TL][2], w_re[TL][TL][2], w_im[TL][TL][2];
#pragma unroll
                            for (int I = 0; I < TL; ++I)
#pragma unroll
                                for (int J = 0; J < TL; ++J)
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        if (!live(I, J)) continue;
                                        x_re[I][J][e] = fma(W0[I][J][e], v_re[I][J][e], ph_re[I][J][e]);
                                        x_im[I][J][e] = fma(W0[I][J][e], v_im[I][J][e], ph_im[I][J][e]);
                                    }
                            apply_rest(v_re, v_im, x_re, x_im);
                            symmetrise(x_re, x_im, w_re, w_im);       // (2 At) v + (previous term)
#pragma unroll
                            for (int I = 0; I < TL; ++I)
#pragma unroll
                                for (int J = 0; J < TL; ++J)
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        if (!live(I, J)) continue;
                                        if constexpr (EML_HALVE_ON_FP64 == 1 || (EML_HALVE_ON_FP64 == 2 && PARITY)) {
                                            ph_re[I][J][e] = hmul * v_re[I][J][e]; ph_im[I][J][e] = hmul * v_im[I][J][e];
                                        } else {
                                            ph_re[I][J][e] = scale_pow2(v_re[I][J][e], hsub); ph_im[I][J][e] = scale_pow2(v_im[I][J][e], hsub);
                                        }
                                        v_re[I][J][e] = w_re[I][J][e]; v_im[I][J][e] = w_im[I][J][e];
                                    }
                            hsub = 0x00100000; hmul = 0.5;
                        }
                        trace_store(v_re, v_im, Kc);          // the last term
                        if constexpr (NA) {
                            const double c = cJ[Kc] * cscale;
#pragma unroll
                            for (int I = 0; I < TL; ++I)
#pragma unroll
                                for (int J = 0; J < TL; ++J)
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        if (!live(I, J)) continue;
                                        acc_re[I][J][e] = fma(c, v_re[I][J][e], acc_re[I][J][e]);
                                        acc_im[I][J][e] = fma(c, v_im[I][J][e], acc_im[I][J][e]);
                                    }
                        }
                    };
                    if (need_acc) run_terms(std::true_type{}); else run_terms(std::false_type{});
                    {   // a non-finite or absurd last term means the enclosure failed: report instead of returning garbage
                        const double chk = fabs(v_re[0][TL - 1][0]) + fabs(v_im[0][TL - 1][0]) + fabs(v_re[TL - 1][0][1]);
                        if (!UNI(chk < 1e300)) capped = true;
                    }
                }
                n_apply += Kc;
                EML_CHECK(Kc >= 1 && Kc <= Scratch::KCAP);
                if (lane < 3 * NTR) S.tr[(Kc + 1) * NTR + lane] = 0.0;      // the output sums read the table in groups of four terms
                __syncwarp();
                PHASE_MARK(2);      // recurrence (and the coefficient table of a non-final step)
                if (UNI(last)) {
                    // ---- outputs of the step: Bessel weights J_k(tau_p R) by Miller's backward recurrence, NCH independent
                    //      output times per lane (the recurrence is one dependent FMA chain per output: several chains in
                    //      flight hide its latency; all chains of a pass share the loads of the trace table).  A pass costs
                    //      as many terms as its LATEST output needs, so the passes are cut from the end of the list: the
                    //      partial pass is the one with the earliest (cheapest) output times.  Terms go in groups of four,
                    //      k a multiple of 4 (the table is zero beyond Kc): the multipliers 2k/z of a group cost one product
                    //      and three subtractions, the even terms of the normalisation sum J_0 + 2 sum J_2k = 1 are known
                    //      statically, and the test "this chain starts here" is only needed down to the smallest start index
                    //      of the pass (a chain that has not started carries exact zeros).
                    //      The per-output scalar work (term count, 2 / z, e^{-ac tau}, observables) is written without branches
                    //      so that the chains of a pass interleave: reciprocals by __drcp_rn instead of divisions, the term count
                    //      from a single-precision cube root with an upward margin (a start index one too high costs one term),
                    //      the exponential issued before the term loop, the observables' coefficients loaded once per pass.
                    constexpr int NCH = LAT ? EML_LAT_OUTPUT_CHAINS : EML_OUTPUT_CHAINS;
                    const double t0o = P.times[kidx], tau_e = (P.times[kidx + q] - t0o) / nsub;
                    const double tzs = 2.0 / Rb;
                    for (int pbase = 0; UNI(pbase < q); pbase += 32 * NCH) {
                        int Kp[NCH], pidx[NCH];
                        double tz[NCH], tau[NCH], ex[NCH];
                        bool direct[NCH];
                        int kmax = 0, kmin = 0x7fffffff;
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            pidx[c] = q - 1 - pbase - 32 * c - lane;
                            const bool act = pidx[c] >= 0;
                            tau[c] = act ? ((nsub == 1) ? P.times[kidx + 1 + max(pidx[c], 0)] - t0o : tau_e) : 0.0;
                        }
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            const double z = tau[c] * Rb, zs = tau[c] * Sb;
                            direct[c] = !(z >= 1e-40);
                            const int kraw = (int)ceil(fma(CHEB_C1 * (1.0 + 1e-6), (double)cbrtf((float)zs), zs) + (CHEB_C0 + 1e-3));
                            const bool on = pidx[c] >= 0 && !direct[c];
                            Kp[c] = on ? min(kraw, Kc) : 0;
                            kmin = min(kmin, on ? Kp[c] : 0x7fffffff);
                            kmax = max(kmax, Kp[c]);
                            tz[c] = direct[c] ? 0.0 : tzs * __drcp_rn(tau[c]);
                            ex[c] = exp(-ac * tau[c]);
                        }
                        kmax = __reduce_max_sync(0xffffffffu, kmax);
                        kmin = __reduce_min_sync(0xffffffffu, kmin);
                        double a0[NCH], a2[NCH], a3[NCH], ssum[NCH], jn[NCH], jn1[NCH];
#pragma unroll
                        for (int c = 0; c < NCH; ++c) { a0[c] = a2[c] = a3[c] = ssum[c] = jn[c] = jn1[c] = 0.0; }
                        int k = (kmax + 3) & ~3;
                        double kd = (double)k;
                        auto four_terms = [&](auto chk_tag) {
                            constexpr bool CHK = decltype(chk_tag)::value;
                            double cm[NCH];
#pragma unroll
                            for (int c = 0; c < NCH; ++c) cm[c] = kd * tz[c];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int kk = k - j;
                                EML_CHECK(kk >= 1 && kk <= Kc + 3 && kk <= Scratch::KCAP + 3);
                                double2 ta = make_double2(0.0, 0.0);
                                if constexpr (!SB) ta = *reinterpret_cast<const double2*>(&S.tr[kk * NTR]);
                                const double2 tb = *reinterpret_cast<const double2*>(&S.tr[kk * NTR + NTR - 2]);
#pragma unroll
                                for (int c = 0; c < NCH; ++c) {
                                    if constexpr (CHK)      // jn is exactly 0.0 before the start: setting the high word makes it 1.0
                                        jn[c] = __hiloint2double(kk == Kp[c] ? 0x3ff00000 : __double2hiint(jn[c]), __double2loint(jn[c]));
                                    if constexpr (!SB) a0[c] = fma(jn[c], ta.x, a0[c]);
                                    a2[c] = fma(jn[c], tb.x, a2[c]); a3[c] = fma(jn[c], tb.y, a3[c]);
                                    if (!(j & 1)) ssum[c] += jn[c];
                                    const double jm = fma(cm[c], jn[c], -jn1[c]);
                                    jn1[c] = jn[c]; jn[c] = jm;
                                    if (j < 3) cm[c] -= tz[c];
                                }
                            }
                            k -= 4; kd -= 4.0;
                        };
                        while (UNI(k >= 4 && k >= kmin)) four_terms(std::true_type{});
                        while (UNI(k >= 4)) four_terms(std::false_type{});
                        const double2 ta = SB ? make_double2(half_tr, half_tr) : *reinterpret_cast<const double2*>(&S.tr[0]);
                        const double2 tb = *reinterpret_cast<const double2*>(&S.tr[NTR - 2]);
                        double r[NCH][4];
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            const double nrm = ex[c] * __drcp_rn(2.0 * ssum[c] + jn[c]);
                            // the generator is trace preserving: rho11 = Tr rho(step start) - rho00 saves a quarter of the sums
                            const double r0 = SB ? half_tr : fma(jn[c], ta.x, a0[c]) * nrm;
                            r[c][0] = direct[c] ? ta.x : r0; r[c][1] = direct[c] ? ta.y : (ta.x + ta.y) - r0;
                            r[c][2] = direct[c] ? tb.x : fma(jn[c], tb.x, a2[c]) * nrm; r[c][3] = direct[c] ? tb.y : fma(jn[c], tb.y, a3[c]) * nrm;
                        }
                        double lsum[NCH];       // loss terms of one output time are summed first (the same order whatever NCH is)
#pragma unroll
                        for (int c = 0; c < NCH; ++c) lsum[c] = 0.0;
                        for (int m = 0; m < K; ++m) {      // as finish_output, the coefficients of an observable shared by the chains
                            const double* O = P.obs + m * 8;
                            const double c0 = O[0], c1 = O[6], c2 = O[2] + O[4], c3 = O[3] - O[5];
                            const double d0 = O[1], d1 = O[7], d2 = O[3] + O[5], d3 = O[4] - O[2];
#pragma unroll
                            for (int c = 0; c < NCH; ++c) {
                                const int kk = kidx + 1 + max(pidx[c], 0);
                                const double ere = c0 * r[c][0] + c1 * r[c][1] + c2 * r[c][2] + c3 * r[c][3];
                                const double eim = d0 * r[c][0] + d1 * r[c][1] + d2 * r[c][2] + d3 * r[c][3];
                                if (pidx[c] >= 0) {
                                    if (expect != nullptr) {      // (caller-owned: only 8-byte alignment is assumed)
                                        double* eo = expect + (((size_t)b * K + m) * T + kk) * 2;
                                        eo[0] = ere; eo[1] = eim;
                                    }
                                    if (tgt != nullptr) {
                                        const double2 tg = *reinterpret_cast<const double2*>(tgt + ((size_t)m * T + kk) * 2);
                                        const double dre = ere - tg.x, dim = eim - tg.y;
                                        lsum[c] += dre * dre + dim * dim;
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int c = 0; c < NCH; ++c) my_loss += lsum[c];
                    }
                    __syncwarp();
                    PHASE_MARK(3);      // Bessel-weighted output sums, observables, loss terms
                }
            }
            kidx += q;
        }
#undef Rb
#undef Sb
#undef ac
#undef tau_max
#undef half_tr
        } else {
        int kidx = 0;
        while (UNI(kidx < T - 1)) {
            int q, nsub;
            double htot;
            {
                // same rule as choose_step() with this kernel's output capacity
                const double t0 = P.times[kidx];
                q = 1;
                while (UNI(q < WQ && kidx + q < T - 1 && nu * (P.times[kidx + q + 1] - t0) <= P.theta_target)) ++q;
                htot = P.times[kidx + q] - t0;
                nsub = 1;
                if (q == 1) {
                    const double th = nu * htot;
                    if (th > P.theta_target) nsub = (int)ceil(th / P.theta_target);
                }
            }
            if (UNI(htot == 0.0)) {      // repeated time points: the state does not move
                const double tr = partial_trace(acc_re, acc_im);
                double yz[NY];
#pragma unroll
                for (int n = 0; n < NY; ++n) yz[n] = tr;
                emit(q, kidx + 1, yz);
                kidx += q;
                continue;
            }
            const double h = htot / nsub;
            const double t0 = P.times[kidx];
            double xs[NY], y[NY];
#pragma unroll
            for (int n = 0; n < NY; ++n) {
                const int p = (lane & 7) + 8 * n;
                xs[n] = (p < q) ? ((nsub == 1) ? (P.times[kidx + 1 + p] - t0) / htot : 1.0) : 0.0;
                y[n] = 0.0;
            }

            for (int sub = 0; UNI(sub < nsub); ++sub) {
                const bool last = (sub == nsub - 1);
                double v_re[TL][TL][2], v_im[TL][TL][2];
#pragma unroll
                for (int I = 0; I < TL; ++I)
#pragma unroll
                    for (int J = 0; J < TL; ++J)
#pragma unroll
                        for (int e = 0; e < 2; ++e) { v_re[I][J][e] = acc_re[I][J][e]; v_im[I][J][e] = acc_im[I][J][e]; }
                if (UNI(h != h_cur)) {
                    const double rs = h / h_cur;
                    h_cur = h;
#pragma unroll
                    for (int I = 0; I < TL; ++I) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) { a_re[I][ks] *= rs; a_im[I][ks] *= rs; }
#pragma unroll
                        for (int J = 0; J < TL; ++J) {
                            F3[I][J][0] *= rs;
                            W0[I][J][0] *= rs; W0[I][J][1] *= rs;
                        }
                    }
                    f2 *= rs; f1 *= rs; f0[0] *= rs; f0[1] *= rs;
                }
                double xp[NY];                    // x^j / j! for this lane's outputs
#pragma unroll
                for (int n = 0; n < NY; ++n) xp[n] = 1.0;
                double cfac = 1.0;                // 1 / j!
                double thr = P.tol;               // tol * j!: |u_j| <= thr  <=>  |term_j| <= tol
                if (last) {
#pragma unroll
                    for (int n = 0; n < NY; ++n) y[n] = 0.0;
                }
                int small = 0;
                int j = 0;
                for (; UNI(j < MCAP); ++j) {
                    // ---- dense output: y += x^j Tr_rest(v_j) for the term that is the INPUT of this application
                    //      (independent of the DMMAs below, so the shuffles overlap with them)
                    {
                        const double tr = partial_trace(v_re, v_im);
                        const double inv = INV_TABLE[j];
#pragma unroll
                        for (int n = 0; n < NY; ++n) {
                            y[n] = fma(xp[n], tr, y[n]);
                            xp[n] *= xs[n] * inv;
                        }
                    }
                    // ---- accumulators start from J(v)/2: real-weighted site stencils (depend only on v)
                    double x_re[TL][TL][2], x_im[TL][TL][2];
#pragma unroll
                    for (int I = 0; I < TL; ++I)
#pragma unroll
                        for (int J = 0; J < TL; ++J)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                x_re[I][J][e] = W0[I][J][e] * v_re[I][J][e];
                                x_im[I][J][e] = W0[I][J][e] * v_im[I][J][e];
                            }
                    apply_rest(v_re, v_im, x_re, x_im);
                    // ---- u_{j+1} = X' + X'^dag;  term_{j+1} = u_{j+1} / (j+1)!
                    cfac *= INV_TABLE[j];
                    thr *= (double)(j + 1);
                    unsigned mx = 0;
                    double w_re[TL][TL][2], w_im[TL][TL][2];
                    symmetrise(x_re, x_im, w_re, w_im);
#pragma unroll
                    for (int I = 0; I < TL; ++I)
#pragma unroll
                        for (int J = 0; J < TL; ++J)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                mx = max(mx, (unsigned)__double2hiint(w_re[I][J][e]) & 0x7fffffffu);
                                mx = max(mx, (unsigned)__double2hiint(w_im[I][J][e]) & 0x7fffffffu);
                            }
#pragma unroll
                    for (int I = 0; I < TL; ++I)
#pragma unroll
                        for (int J = 0; J < TL; ++J)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                v_re[I][J][e] = w_re[I][J][e]; v_im[I][J][e] = w_im[I][J][e];
                                acc_re[I][J][e] = fma(cfac, w_re[I][J][e], acc_re[I][J][e]);
                                acc_im[I][J][e] = fma(cfac, w_im[I][J][e], acc_im[I][J][e]);
                            }
                    mx = __reduce_max_sync(0xffffffffu, mx);
                    small = (mx <= ((unsigned)__double2hiint(thr) & 0x7fffffffu)) ? small + 1 : 0;   // mx is warp-uniform (REDUX)
                    if (UNI(small >= 2)) { ++j; break; }
                }
                {   // the final term computed above has not been traced yet
                    const double tr = partial_trace(v_re, v_im);
#pragma unroll
                    for (int n = 0; n < NY; ++n) y[n] = fma(xp[n], tr, y[n]);
                }
                n_apply += j;
                if (j >= MCAP && small < 2) capped = true;
            }
            emit(q, kidx + 1, y);
            kidx += q;
        }

        }

        // ---- K5: loss, fixed-order warp reduction
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
        if (lane == 0) {
            if (loss != nullptr) loss[b] = (tgt != nullptr) ? my_loss / ((double)T * (double)K) : 0.0;
            if (status != nullptr) status[b] = capped ? -1 : n_apply;
        }
        __syncwarp();
        if constexpr (CHEB) { PHASE_MARK(4); }     // loss reduction, result stores
    }
}

#define EML_WARP_INST(TL, RH, CH, PA, SB) template __global__ void eval_warp_mma_kernel<TL, RH, CH, PA, SB>(const DevProblem, const long long, const double*, const int*, double*, double*, int*, unsigned int*, const int*, const double*, const int);
EML_WARP_INST(1, true, false, false, false) EML_WARP_INST(1, false, false, false, false) EML_WARP_INST(2, true, false, false, false) EML_WARP_INST(2, false, false, false, false)
EML_WARP_INST(1, true, true, false, false) EML_WARP_INST(1, false, true, false, false) EML_WARP_INST(2, true, true, false, false) EML_WARP_INST(2, false, true, false, false)
EML_WARP_INST(2, true, true, true, false) EML_WARP_INST(2, false, true, true, false)
EML_WARP_INST(2, true, true, true, true) EML_WARP_INST(2, false, true, true, true)
#define EML_WARP_INST_LAT(TL, RH, PA, SB) template __global__ void eval_warp_mma_kernel<TL, RH, true, PA, SB, true>(const DevProblem, const long long, const double*, const int*, double*, double*, int*, unsigned int*, const int*, const double*, const int);
EML_WARP_INST_LAT(1, true, false, false) EML_WARP_INST_LAT(1, false, false, false) EML_WARP_INST_LAT(2, true, false, false) EML_WARP_INST_LAT(2, false, false, false)
EML_WARP_INST_LAT(2, true, true, false) EML_WARP_INST_LAT(2, false, true, false) EML_WARP_INST_LAT(2, true, true, true) EML_WARP_INST_LAT(2, false, true, true)

// =====================================================================================================================
// d = 4 (and d = 2): TWO MODELS PER WARP.  Models with one or two sites are embedded into three sites by spectator sites in
// |0><0| (identity dynamics), so only the block of the 8 x 8 tile with spectator bit 0 in row and column is alive.  The
// lowest bit of the kernel's index is such a spectator bit: here it is the MODEL index instead -- rho = rho_A (+) rho_B,
// G = G_A (+) G_B (the records of the prepare kernel already hold G_4 (x) 1, so lane (g, t) simply reads the fragments of
// model g & 1), weights, shift and scale of the Chebyshev recurrence per row model.  No flip connects the two blocks, so
// one recurrence (the same DMMAs, half the stencils: the spectator's flip is dropped) advances both models; the macro steps
// follow the shorter tau_max, the number of terms the longer series.  Partial traces are kept per model (8 values per
// term), the Bessel-weighted output sums, observables and losses are per model as in eval_warp_mma_kernel.
// Used for launches with more models than warp slots (below that the FP64 pipe is idle and two warps are faster than one).
struct alignas(16) Pack2Scratch {
    static constexpr int KCAP = 124;     // terms per macro step: the trace table holds 8 values per term
    double jw[2][32];
    double gdiag[2][16];
    double plan[2][4];                   // R, S, ac, tau_max per model
    alignas(16) double xt_re[128];
    alignas(16) double xt_im[128];
    double cj[2][KCAP + 2];
    alignas(16) double tr[(KCAP + 4) * 8];
};

template <bool REAL_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 16 / WARPS_PER_CTA)
eval_pack2_kernel(const DevProblem P, const long long n_models, const int* __restrict__ target_set, double* __restrict__ expect,
                  double* __restrict__ loss, int* __restrict__ status, unsigned int* __restrict__ counter,
                  const int* __restrict__ order, const double* __restrict__ prep, const int binned) {
    constexpr int N = 3, D = 8, KCAP = Pack2Scratch::KCAP;
    constexpr int NA = prep_fragments<1, REAL_H, false>(), REC = prep_record_doubles<1, REAL_H, false>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Pack2Scratch& S = reinterpret_cast<Pack2Scratch*>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, gp = g & 1;
    const int mr = g & 1;                 // the model this lane's row belongs to
    const int T = P.n_times, K = P.n_obs;
    const int obs_site = P.obs_site;
    auto pos = [&](int s) { return (s == obs_site) ? N - 1 : (s < obs_site ? N - 2 - s : N - 1 - s); };
    auto orig_index = [&](int x) {
        int o = 0;
#pragma unroll
        for (int s = 0; s < N; ++s) o |= ((x >> pos(s)) & 1) << (N - 1 - s);
        return o;
    };
    // partial-trace lanes (see trace_put below)
    const int tl = lane & ~4;
    const bool tr_lane = (tl == 0 || tl == 18 || tl == 2 || tl == 11);
    const int tr_off = 4 * mr + (tl == 0 ? 0 : tl == 18 ? 1 : tl == 2 ? 2 : 3);
    const long long n_units = (n_models + 1) / 2;

    pdl_wait();
    for (;;) {
        unsigned int bidx = 0;
        if (lane == 0) bidx = atomicAdd(counter, 1u);
        bidx = __shfl_sync(0xffffffffu, bidx, 0);
        if (UNI((long long)bidx >= n_units)) { fetch_ws_release(counter, bidx, n_units, binned, lane); break; }
        long long bm[2];
        bool live[2] = {true, 2LL * bidx + 1 < n_models};
        auto ranked = [&](const long long r) -> long long {
            return (order == nullptr) ? r : binned ? (long long)order_list_get(counter, order, (unsigned int)r, lane) : (long long)order[r];
        };
        bm[0] = ranked(2LL * bidx);
        bm[1] = live[1] ? ranked(2LL * bidx + 1) : bm[0];
        // a model that cannot be evaluated (non-finite parameters or an absurd step count: R = NaN) is reported and its slot
        // runs a copy of its partner, so that nothing non-finite enters the shared tile
        long long src[2] = {bm[0], bm[1]};
        {
            const double r0 = prep[(size_t)bm[0] * REC + NA * 32 + 48], r1 = prep[(size_t)bm[1] * REC + NA * 32 + 48];
            const bool bad0 = !(r0 == r0), bad1 = !(r1 == r1);
            if (lane == 0) {
                if (bad0) { if (loss != nullptr) loss[bm[0]] = nan(""); if (status != nullptr) status[bm[0]] = -1; }
                if (bad1 && live[1]) { if (loss != nullptr) loss[bm[1]] = nan(""); if (status != nullptr) status[bm[1]] = -1; }
            }
            if (bad0) live[0] = false;
            if (bad1) live[1] = false;
            if (UNI(!live[0] && !live[1])) { __syncwarp(); continue; }
            if (bad0) src[0] = bm[1];
            if (bad1 || !live[1]) src[1] = live[0] ? bm[0] : bm[1];
        }
        __syncwarp();      // the previous unit's readers of the scratch are done
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const double* const R = prep + (size_t)src[m] * REC;
            S.jw[m][lane] = R[NA * 32 + lane];
            if (lane < 16) S.gdiag[m][lane] = R[NA * 32 + 32 + lane];
            if (lane < 4) S.plan[m][lane] = R[NA * 32 + 48 + lane];
        }
        double a_re[2], a_im[2];
        {
            const double* const R = prep + (size_t)src[mr] * REC;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                a_im[ks] = R[ks * 32 + lane];
                a_re[ks] = REAL_H ? 0.0 : R[(2 + ks) * 32 + lane];
            }
        }
        __syncwarp();
        // ---- weights of this lane's row model: generator -> 2 (L + ac) / R, stencil weights halved (X' + X'^dag doubles them)
        const double* const jwp = S.jw[mr];
        const double sc = 2.0 / S.plan[mr][0], acl = S.plan[mr][2];
        double W0[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double w = jwp[2 * 8 + (g >> 2) * 2 + (t >> 1)] + jwp[1 * 8 + ((g >> 1) & 1) * 2 + (t & 1)] + jwp[0 * 8 + (g & 1) * 2 + e];
            if (REAL_H) w += S.gdiag[mr][g] + S.gdiag[mr][2 * t + e];
            W0[e] = (0.5 * w + 0.5 * acl) * sc;
        }
        const double f2 = 0.5 * jwp[2 * 8 + 4 + (g >> 2) * 2 + (t >> 1)] * sc;
        const double f1 = 0.5 * jwp[1 * 8 + 4 + ((g >> 1) & 1) * 2 + (t & 1)] * sc;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) { a_re[ks] *= sc; a_im[ks] *= sc; }
        // ---- state: block (model, model) of the tile carries rho0 of the embedded problem (spectator bits cleared)
        double acc_re[2], acc_im[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int r = orig_index(g & ~1), c = orig_index(2 * t);
            const bool alive = (e == mr);
            acc_re[e] = alive ? P.rho0[2 * (r * D + c)] : 0.0;
            acc_im[e] = alive ? P.rho0[2 * (r * D + c) + 1] : 0.0;
        }
        const double* tgt[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int tset = (target_set != nullptr) ? target_set[bm[m]] : 0;
            tgt[m] = (P.targets != nullptr) ? P.targets + (size_t)tset * K * T * 2 : nullptr;
        }
        double my_loss[2] = {0.0, 0.0};
        int n_apply = 0;
        bool capped = false;

        // 2x2 partial trace per model: a partial-trace element (row and column agree on the non-observed bits, the model bit
        // included) sits at e = g & 1 of the lanes with t0 = g1; summing over g1 (lane ^ 9) leaves model g0's totals of the
        // observed-bit pair (g2, t1) in lanes 16 g2 + 2 t1 + 4 g0 (and their partners ^ 9): rho00 <- lane 4m, rho11 <- 18 + 4m,
        // Re rho01 <- 2 + 4m, Im rho01 <- 11 + 4m
        const unsigned tr_base = (unsigned)__cvta_generic_to_shared(&S.tr[tr_off]);
        auto trace_put = [&](const double (&xr)[2], const double (&xi)[2], const int kterm) {
            const double er = gp ? xr[1] : xr[0], ei = gp ? xi[1] : xi[0];
            const double sr = er + __shfl_xor_sync(0xffffffffu, er, 9), si = ei + __shfl_xor_sync(0xffffffffu, ei, 9);
            EML_CHECK(kterm >= 0 && kterm <= KCAP);
            asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.shared.f64 [%0], %1; }"
                         :: "r"(tr_base + (unsigned)(kterm * 64)), "d"(tl == 11 ? si : sr), "r"((int)tr_lane) : "memory");
        };
        auto finish_output = [&](const int m, const int k, const double (&r)[4]) {
            double ls = 0.0;
            for (int o = 0; o < K; ++o) {
                const double* O = P.obs + o * 8;
                const double ere = O[0] * r[0] + O[6] * r[1] + (O[2] + O[4]) * r[2] + (O[3] - O[5]) * r[3];
                const double eim = O[1] * r[0] + O[7] * r[1] + (O[3] + O[5]) * r[2] + (O[4] - O[2]) * r[3];
                if (expect != nullptr) {
                    double* e = expect + (((size_t)bm[m] * K + o) * T + k) * 2;
                    e[0] = ere; e[1] = eim;
                }
                if (tgt[m] != nullptr) {
                    const double dre = ere - tgt[m][((size_t)o * T + k) * 2], dim = eim - tgt[m][((size_t)o * T + k) * 2 + 1];
                    ls += dre * dre + dim * dim;
                }
            }
            my_loss[m] += ls;
        };
        auto first_like_output = [&](const int kfirst, const int count) {     // the state itself at `count` output times from kfirst
            __syncwarp();
            trace_put(acc_re, acc_im, 0);
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (!live[m]) continue;
                const double r[4] = {S.tr[4 * m], S.tr[4 * m + 1], S.tr[4 * m + 2], S.tr[4 * m + 3]};
                for (int pb = 0; UNI(pb < count); pb += 32)
                    if (pb + lane < count) finish_output(m, kfirst + pb + lane, r);
            }
            __syncwarp();
        };
        first_like_output(0, 1);

        const double tau_max = fmin(S.plan[0][3], S.plan[1][3]);
        int kidx = 0;
        while (UNI(kidx < T - 1)) {
            const double t0 = P.times[kidx];
            int q = 0;
            for (;;) {
                const int cand = kidx + q + 1 + lane;
                const unsigned in = __ballot_sync(0xffffffffu, cand <= T - 1 && P.times[min(cand, T - 1)] - t0 <= tau_max);
                const int run = __ffs(~in) - 1;
                if (run >= 0) { q += run; break; }
                q += 32;
            }
            int nsub = 1;
            if (UNI(q == 0)) { nsub = (int)ceil((P.times[kidx + 1] - t0) / tau_max); q = 1; }
            const double tau_end = (P.times[kidx + q] - t0) / nsub;
            const bool final_step = (kidx + q == T - 1);
            if (UNI(!(tau_end * fmin(S.plan[0][0], S.plan[1][0]) >= 1e-40))) {     // repeated time points: the state does not move
                first_like_output(kidx + 1, q);
                kidx += q;
                continue;
            }
            int Kcm[2], Kc = 0;
#pragma unroll
            for (int m = 0; m < 2; ++m) { Kcm[m] = min(cheb_terms(tau_end * S.plan[m][1]), KCAP); Kc = max(Kc, Kcm[m]); }
            for (int sub = 0; UNI(sub < nsub); ++sub) {
                const bool last = (sub == nsub - 1);
                const bool need_acc = UNI(!(final_step && last));
                double cscale = 0.0;       // of this lane's row model
                if (need_acc) {
                    __syncwarp();
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        const double tz = 2.0 / (tau_end * S.plan[m][0]);
                        double jn = 1.0, jn1 = 0.0, ssum = 0.0;
                        for (int k = Kc; k >= 1; --k) {
                            if ((k & 31) == lane) S.cj[m][k] = jn;
                            if (!(k & 1)) ssum += jn;
                            const double jm = fma((double)k * tz, jn, -jn1);
                            jn1 = jn; jn = jm;
                        }
                        if (lane == 0) S.cj[m][0] = jn;
                        const double cs = exp(-S.plan[m][2] * tau_end) / (2.0 * ssum + jn);
                        if (m == mr) cscale = cs;
                    }
                    __syncwarp();
                }
                double v_re[2] = {acc_re[0], acc_re[1]}, v_im[2] = {acc_im[0], acc_im[1]}, ph_re[2] = {0.0, 0.0}, ph_im[2] = {0.0, 0.0};
                acc_re[0] = acc_re[1] = acc_im[0] = acc_im[1] = 0.0;
                auto run_terms = [&](auto na_tag) {
                    constexpr bool NAC = decltype(na_tag)::value;
                    double hmul = 1.0;
                    for (int k = 0; UNI(k < Kc); ++k) {
                        trace_put(v_re, v_im, k);
                        if constexpr (NAC) {
                            const double c = S.cj[mr][k] * cscale;
#pragma unroll
                            for (int e = 0; e < 2; ++e) { acc_re[e] = fma(c, v_re[e], acc_re[e]); acc_im[e] = fma(c, v_im[e], acc_im[e]); }
                        }
                        double x_re[2], x_im[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) { x_re[e] = fma(W0[e], v_re[e], ph_re[e]); x_im[e] = fma(W0[e], v_im[e], ph_im[e]); }
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            x_re[e] = fma(f2, __shfl_xor_sync(0xffffffffu, v_re[e], 18), x_re[e]);
                            x_im[e] = fma(f2, __shfl_xor_sync(0xffffffffu, v_im[e], 18), x_im[e]);
                        }
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            x_re[e] = fma(f1, __shfl_xor_sync(0xffffffffu, v_re[e], 9), x_re[e]);
                            x_im[e] = fma(f1, __shfl_xor_sync(0xffffffffu, v_im[e], 9), x_im[e]);
                        }
                        // X' = J(v)/2 + G v on the DMMA pipe, B = conj of the lane's own registers
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            if (!REAL_H) {
                                dmma(x_re[0], x_re[1], a_re[ks], v_re[ks]);
                                dmma(x_im[0], x_im[1], a_re[ks], neg(v_im[ks]));
                            }
                            dmma(x_re[0], x_re[1], a_im[ks], v_im[ks]);
                            dmma(x_im[0], x_im[1], a_im[ks], v_re[ks]);
                        }
                        // w = X' + X'^dag through the swizzled tile (as the d = 8 kernel)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int Cc = 2 * t + e;
                            const int m_ = 4 * (t & 1) + 8 * (((t >> 1) & 1) ^ e);
                            S.xt_re[Cc * 16 + (g ^ m_)] = x_re[e];
                            S.xt_im[Cc * 16 + (g ^ m_)] = x_im[e];
                        }
                        __syncwarp();
                        {
                            const int m_ = 4 * ((g >> 1) & 1) + 8 * (((g >> 2) & 1) ^ (g & 1));
                            const int o = g * 16 + ((2 * t) ^ m_);
                            const double2 tr_ = *reinterpret_cast<const double2*>(&S.xt_re[o]);
                            const double2 ti_ = *reinterpret_cast<const double2*>(&S.xt_im[o]);
#pragma unroll
                            for (int e = 0; e < 2; ++e) { ph_re[e] = hmul * v_re[e]; ph_im[e] = hmul * v_im[e]; }
                            v_re[0] = x_re[0] + tr_.x; v_re[1] = x_re[1] + tr_.y;
                            v_im[0] = x_im[0] - ti_.x; v_im[1] = x_im[1] - ti_.y;
                        }
                        __syncwarp();
                        hmul = 0.5;
                    }
                    trace_put(v_re, v_im, Kc);
                    if constexpr (NAC) {
                        const double c = S.cj[mr][Kc] * cscale;
#pragma unroll
                        for (int e = 0; e < 2; ++e) { acc_re[e] = fma(c, v_re[e], acc_re[e]); acc_im[e] = fma(c, v_im[e], acc_im[e]); }
                    }
                };
                if (need_acc) run_terms(std::true_type{}); else run_terms(std::false_type{});
                {   // a non-finite or absurd last term means an enclosure failed.  Both models of the pair are reported: 0 x inf in the
                    // block-diagonal products lets a diverging model poison its partner's rows, so neither result can be trusted
                    const double chk = fabs(v_re[0]) + fabs(v_im[0]) + fabs(v_re[1]) + fabs(v_im[1]);
                    if (!UNI(chk < 1e300)) capped = true;
                }
                n_apply += Kc;
                if (lane < 24) S.tr[(Kc + 1) * 8 + lane] = 0.0;      // the output sums read the table in groups of four terms
                __syncwarp();
                if (UNI(last)) {
                    // ---- outputs of the step, per model: as eval_warp_mma_kernel (one output time per lane and pass, passes cut
                    //      from the end of the list, terms in groups of four)
                    const double t0o = P.times[kidx], tau_e = (P.times[kidx + q] - t0o) / nsub;
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        if (!live[m]) continue;
                        const double Rb = S.plan[m][0], Sb = S.plan[m][1], acm = S.plan[m][2];
                        const double tzs = 2.0 / Rb;
                        for (int pbase = 0; UNI(pbase < q); pbase += 32) {
                            const int pidx = q - 1 - pbase - lane;
                            const bool act = pidx >= 0;
                            const double tau = act ? ((nsub == 1) ? P.times[kidx + 1 + max(pidx, 0)] - t0o : tau_e) : 0.0;
                            const double z = tau * Rb, zs = tau * Sb;
                            const bool direct = !(z >= 1e-40);
                            const int kraw = (int)ceil(fma(CHEB_C1 * (1.0 + 1e-6), (double)cbrtf((float)zs), zs) + (CHEB_C0 + 1e-3));
                            const bool on = act && !direct;
                            const int Kp = on ? min(kraw, Kcm[m]) : 0;
                            const int kmax = __reduce_max_sync(0xffffffffu, Kp);
                            const int kmin = __reduce_min_sync(0xffffffffu, on ? Kp : 0x7fffffff);
                            const double tz = direct ? 0.0 : tzs * __drcp_rn(tau);
                            const double ex = exp(-acm * tau);
                            double a0 = 0.0, a2 = 0.0, a3 = 0.0, ssum = 0.0, jn = 0.0, jn1 = 0.0;
                            int k = (kmax + 3) & ~3;
                            double kd = (double)k;
                            auto four_terms = [&](auto chk_tag) {
                                constexpr bool CHK = decltype(chk_tag)::value;
                                double cm = kd * tz;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int kk = k - j;
                                    EML_CHECK(kk >= 1 && kk <= Kc + 3 && kk <= KCAP + 3);
                                    const double2 ta = *reinterpret_cast<const double2*>(&S.tr[kk * 8 + 4 * m]);
                                    const double2 tb = *reinterpret_cast<const double2*>(&S.tr[kk * 8 + 4 * m + 2]);
                                    if constexpr (CHK)
                                        jn = __hiloint2double(kk == Kp ? 0x3ff00000 : __double2hiint(jn), __double2loint(jn));
                                    a0 = fma(jn, ta.x, a0); a2 = fma(jn, tb.x, a2); a3 = fma(jn, tb.y, a3);
                                    if (!(j & 1)) ssum += jn;
                                    const double jm = fma(cm, jn, -jn1);
                                    jn1 = jn; jn = jm;
                                    if (j < 3) cm -= tz;
                                }
                                k -= 4; kd -= 4.0;
                            };
                            while (UNI(k >= 4 && k >= kmin)) four_terms(std::true_type{});
                            while (UNI(k >= 4)) four_terms(std::false_type{});
                            const double2 ta = *reinterpret_cast<const double2*>(&S.tr[4 * m]);
                            const double2 tb = *reinterpret_cast<const double2*>(&S.tr[4 * m + 2]);
                            const double nrm = ex * __drcp_rn(2.0 * ssum + jn);
                            const double r0 = fma(jn, ta.x, a0) * nrm;
                            const double r[4] = {direct ? ta.x : r0, direct ? ta.y : (ta.x + ta.y) - r0,
                                                 direct ? tb.x : fma(jn, tb.x, a2) * nrm, direct ? tb.y : fma(jn, tb.y, a3) * nrm};
                            if (act) finish_output(m, kidx + 1 + pidx, r);
                        }
                    }
                    __syncwarp();
                }
            }
            kidx += q;
        }
        // ---- losses, fixed-order warp reductions
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_loss[m] += __shfl_xor_sync(0xffffffffu, my_loss[m], o);
            if (lane == 0 && live[m]) {
                if (loss != nullptr) loss[bm[m]] = (tgt[m] != nullptr) ? my_loss[m] / ((double)T * (double)K) : 0.0;
                if (status != nullptr) status[bm[m]] = capped ? -1 : n_apply;
            }
        }
        __syncwarp();
    }
}
template __global__ void eval_pack2_kernel<true>(const DevProblem, const long long, const int*, double*, double*, int*, unsigned int*, const int*, const double*, const int);
template __global__ void eval_pack2_kernel<false>(const DevProblem, const long long, const int*, double*, double*, int*, unsigned int*, const int*, const double*, const int);

__global__ void __launch_bounds__(256) hint_cost_kernel(const int n_models, const int* __restrict__ hint, float* __restrict__ cost) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_models) cost[b] = (float)max(hint[b], 0);
}

// longest-first model order into `order` (capacity 2 * ORDER_MAX_MODELS ints: order + float cost scratch).
// cost_hint (device, optional): a measured cost per model, e.g. the generator applications the same chain's previous
// proposal needed; without it a proxy computed from the parameters is used.
void launch_model_order(const DevProblem& P, long long n_models, const double* params, int* order, const int* cost_hint,
                        cudaStream_t stream) {
    float* cost = reinterpret_cast<float*>(order + ORDER_MAX_MODELS);
    if (cost_hint != nullptr)
        hint_cost_kernel<<<(unsigned)((n_models + 255) / 256), 256, 0, stream>>>((int)n_models, cost_hint, cost);
    else
        model_cost_kernel<<<(unsigned)((n_models + 255) / 256), 256, 0, stream>>>(P, (int)n_models, params, cost);
    order_models_kernel<<<1, 1024, 0, stream>>>((int)n_models, cost, order);
}

// one-CTA counting sort of per-model costs, largest first (used by the d = 32 orbit kernel's launcher)
void launch_order_from_cost(int n_models, const float* cost, int* order, cudaStream_t stream) {
    order_models_kernel<<<1, 1024, 0, stream>>>(n_models, cost, order);
}

bool warp_mma_supported(const DevProblem& P, bool hermitian) {
    return hermitian && (P.n_tls == 3 || P.n_tls == 4) && P.n_params <= MAXP;
}

static cudaError_t init_inv_table() {
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && done[dev]) return cudaSuccess;
    double h[MCAP + 2];
    for (int j = 0; j < MCAP + 2; ++j) h[j] = 1.0 / (double)(j + 1);
    e = cudaMemcpyToSymbol(INV_TABLE, h, sizeof(h));
    if (e == cudaSuccess && dev < 64) done[dev] = true;
    return e;
}

using WarpKern = void (*)(const DevProblem, const long long, const double*, const int*, double*, double*, int*,
                          unsigned int*, const int*, const double*, const int);
using PrepKern = void (*)(const DevProblem, const long long, const double*, double*, float*, int*, unsigned int*, const double, const int, const int);

struct WarpVariant {
    WarpKern eval = nullptr;
    WarpKern eval_lat = nullptr; // latency variant (Chebyshev kernels): launches with no more models than it has warp slots
    PrepKern prep = nullptr;     // Chebyshev variants only
    size_t smem = 0, prep_smem = 0;
    int rec = 0, kcap = 0, id = 0;
};
template <int TL, bool RH, bool CH, bool PA, bool SB>
static WarpVariant make_variant() {
    WarpVariant v;
    v.eval = (WarpKern)eval_warp_mma_kernel<TL, RH, CH, PA, SB>;
    v.smem = WARPS_PER_CTA * sizeof(WarpScratch<TL, CH, SB, PA>);
    v.kcap = WarpScratch<TL, CH, SB, PA>::KCAP;
    v.id = (SB ? 16 : 0) + (PA ? 8 : 0) + (TL == 2 ? 4 : 0) + (CH ? 2 : 0) + (RH ? 1 : 0);
    if constexpr (CH) {
        v.eval_lat = (WarpKern)eval_warp_mma_kernel<TL, RH, CH, PA, SB, true>;
        v.prep = (PrepKern)prepare_models_kernel<TL, RH, PA>;
        v.prep_smem = PREP_WARPS * sizeof(PrepScratch<TL, RH>);
        v.rec = prep_record_doubles<TL, RH, PA>();
    }
    return v;
}
template <int TL, bool CH, bool PA, bool SB>
static WarpVariant make_variant_rh(bool real_h) {
    return real_h ? make_variant<TL, true, CH, PA, SB>() : make_variant<TL, false, CH, PA, SB>();
}
static WarpVariant pick_variant(bool tl2, bool real_h, bool chebyshev, bool parity, bool sector) {
    if (!tl2) return chebyshev ? make_variant_rh<1, true, false, false>(real_h) : make_variant_rh<1, false, false, false>(real_h);
    if (!chebyshev) return make_variant_rh<2, false, false, false>(real_h);
    if (sector) return make_variant_rh<2, true, true, true>(real_h);
    if (parity) return make_variant_rh<2, true, true, false>(real_h);
    return make_variant_rh<2, true, false, false>(real_h);
}

// bytes of prepare-kernel workspace per model of the Chebyshev variant a plan runs (0: Taylor variant, no prepare kernel)
size_t warp_prep_bytes_per_model(const DevProblem& P, bool real_h, bool chebyshev, bool parity_h, bool sector_b) {
    const bool tl2 = (P.n_tls == 4), parity = parity_h && chebyshev && tl2, sector = parity && sector_b;
    return (size_t)pick_variant(tl2, real_h, chebyshev, parity, sector).rec * sizeof(double);
}

// prep_ws: workspace of the prepare kernel for prep_cap models (larger batches run in chunks of prep_cap)
cudaError_t launch_warp_mma(const DevProblem& P, long long n_models, const double* params, const int* target_set,
                            double* expect, double* loss, int* status, unsigned int* counter, int* order,
                            bool real_h, bool chebyshev, bool parity_h, bool sector_b, const int* cost_hint,
                            double* prep_ws, long long prep_cap, bool pack2_ok, bool pdl_ok, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = init_inv_table()) != cudaSuccess) return e;
    const bool tl2 = (P.n_tls == 4);
    const bool parity = parity_h && chebyshev && tl2;     // H block diagonal by parity: half the DMMAs (d = 16)
    const bool sector = parity && sector_b;               // parity-even part of rho0 is a steady state: evolve the odd tiles only
    const WarpVariant v = pick_variant(tl2, real_h, chebyshev, parity, sector);
    // per device and kernel variant: dynamic shared memory opt-in and the resident CTAs per SM
    static int per_sm_cache[64][32] = {{0}};
    static int per_sm_lat_cache[64][32] = {{0}};
    int per_sm = (dev < 64) ? per_sm_cache[dev][v.id] : 0;
    int per_sm_lat = (dev < 64) ? per_sm_lat_cache[dev][v.id] : 0;
    if (per_sm == 0) {
        if (v.eval_lat != nullptr) {
            if ((e = cudaFuncSetAttribute(v.eval_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem)) != cudaSuccess) return e;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_lat, v.eval_lat, WARPS_PER_CTA * 32, v.smem)) != cudaSuccess) return e;
            if (per_sm_lat < 1) per_sm_lat = 1;
            if (dev < 64) per_sm_lat_cache[dev][v.id] = per_sm_lat;
        }
        if ((e = cudaFuncSetAttribute(v.eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem)) != cudaSuccess) return e;
        if (v.prep != nullptr &&
            (e = cudaFuncSetAttribute(v.prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(v.prep_smem + PREP_LIN_STAGE_MAX))) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.eval, WARPS_PER_CTA * 32, v.smem)) != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        if (dev < 64) per_sm_cache[dev][v.id] = per_sm;
        if (std::getenv("EML_B200_DEBUG") != nullptr)
            fprintf(stderr, "eml: warp kernel variant %d: %d CTAs of %d warps per SM, %zu B shared memory per CTA\n", v.id, per_sm,
                    WARPS_PER_CTA, v.smem);
    }
    // two d = 4 (d = 2) models per warp: launches with more models than warp slots (EML_PACK2=0 switches it off)
    static const int pack2_mode = [] { const char* e = std::getenv("EML_PACK2"); return e ? std::atoi(e) : 1; }();      // 2: whenever n > slots
    const bool pack2_on = pack2_mode != 0;
    using PackKern = void (*)(const DevProblem, const long long, const int*, double*, double*, int*, unsigned int*, const int*, const double*, const int);
    const PackKern pack_kern = real_h ? (PackKern)eval_pack2_kernel<true> : (PackKern)eval_pack2_kernel<false>;
    const size_t pack_smem = WARPS_PER_CTA * sizeof(Pack2Scratch);
    const bool pack2_possible = pack2_on && pack2_ok && chebyshev && !tl2;
    static int per_sm_pack_cache[64][2] = {{0}};
    int per_sm_pack = (dev < 64) ? per_sm_pack_cache[dev][real_h ? 1 : 0] : 0;
    if (pack2_possible && per_sm_pack == 0) {
        if ((e = cudaFuncSetAttribute(pack_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pack_smem)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_pack, pack_kern, WARPS_PER_CTA * 32, pack_smem)) != cudaSuccess) return e;
        if (per_sm_pack < 1) per_sm_pack = 1;
        if (dev < 64) per_sm_pack_cache[dev][real_h ? 1 : 0] = per_sm_pack;
    }
    const long long max_ctas = (long long)sms * per_sm;      // persistent grid: every resident warp fetches models
    if (chebyshev && (prep_ws == nullptr || prep_cap < 1)) return cudaErrorInvalidValue;
    const long long chunk_cap = chebyshev ? prep_cap : n_models;
    for (long long off = 0; off < n_models; off += chunk_cap) {
        const long long n = (n_models - off < chunk_cap) ? n_models - off : chunk_cap;
        // (a pair occupies one warp for one recurrence and TWO output phases: with fewer than ~3 models per warp slot the launch
        //  is bound by that latency, not by the FP64 pipe, and one model per warp is faster: 5 001 models 26.0 M against 27.9 M
        //  evaluations/s, 7 200 models 34.9 M against 30.0 M, 16 384 models 46.3 M against 33.3 M, 65 536 models 52.0 M against 35.5 M)
        const long long pack2_min = (pack2_mode == 2 ? 2LL : 5LL) * WARPS_PER_CTA * max_ctas / 2;      // 2.5 models per warp slot
        const bool pack2 = pack2_possible && n > pack2_min && n <= ORDER_MAX_MODELS && order != nullptr;
        const int kcap = pack2 ? Pack2Scratch::KCAP : v.kcap;
        const double zcap = chebyshev ? cheb_zcap_for(kcap) : 0.0;
        const double* p_ = params + off * P.n_params;
        const int* ts_ = target_set ? target_set + off : nullptr;
        double* ex_ = expect ? expect + (size_t)off * P.n_obs * P.n_times * 2 : nullptr;
        double* lo_ = loss ? loss + off : nullptr;
        int* st_ = status ? status + off : nullptr;
        // Model order: longest first.  Chebyshev variants: by the generator applications the prepare kernel computes (exact);
        // Taylor variants: by the caller's measured cost or a proxy computed from the parameters.
        const bool sort = order != nullptr && n > (long long)WARPS_PER_CTA * max_ctas && n <= ORDER_MAX_MODELS;
        float* cost = (order != nullptr && n <= ORDER_MAX_MODELS) ? reinterpret_cast<float*>(order + ORDER_MAX_MODELS) : nullptr;
        // up to ORDER_LIST_CAP models: cost-bin lists filled by the prepare kernel instead of an ordering kernel
        static const bool lists_on = [] { const char* e = std::getenv("EML_ORDER_LISTS"); return !(e && e[0] == '0'); }();
        const int binned = (chebyshev && sort && lists_on && n <= ORDER_LIST_CAP) ? 1 : 0;
        // (the fetch counter and the bin counts behind it are zero here: the evaluator's last warp leaves them so)
        if (chebyshev) {
            // persistent CTAs (the staged linear program is loaded once per CTA); as many as fit the device
            const size_t lin_bytes = (size_t)(P.lin_rows + P.lin_nnz) * 16;
            const int stage = (P.lin_rows > 0 && lin_bytes <= (size_t)PREP_LIN_STAGE_MAX) ? 1 : 0;
            const size_t psmem = v.prep_smem + (stage ? lin_bytes : 0);
            int prep_per_sm = (int)((size_t)(227 * 1024) / (psmem + 1024));
            if (prep_per_sm > 16 / (PREP_WARPS / 4)) prep_per_sm = 16 / (PREP_WARPS / 4);      // 64 warps per SM at most
            if (prep_per_sm < 1) prep_per_sm = 1;
            long long pctas = (n + PREP_WARPS - 1) / PREP_WARPS;
            if (pctas > (long long)sms * prep_per_sm) pctas = (long long)sms * prep_per_sm;
            v.prep<<<(unsigned)pctas, PREP_WARPS * 32, psmem, stream>>>(P, n, p_, prep_ws, (sort && !binned) ? cost : nullptr,
                                                                         binned ? order : nullptr, counter, zcap, kcap, stage);
            if (sort && !binned) order_models_kernel<<<1, 1024, 0, stream>>>((int)n, cost, order);
        } else if (sort) {
            launch_model_order(P, n, p_, order, cost_hint ? cost_hint + off : nullptr, stream);
        }
        // The evaluator behind a prepare kernel is a programmatic dependent launch: its CTAs become resident while the
        // prepare grid drains and wait (griddepcontrol.wait) for the records.  Not inside a stream capture (the host graphs),
        // and not for the chain engine, whose steps it made slower (4096 d = 16 chains: 10.35 M against 10.54 M steps/s;
        // the evaluation batch alone: 12.47 M against 12.38 M evaluations/s).
        static const bool pdl_env = [] { const char* e = std::getenv("EML_PDL"); return !(e && e[0] == '0'); }();
        const bool pdl_on = pdl_env && pdl_ok;
        cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
        if (chebyshev && pdl_on && (e = cudaStreamIsCapturing(stream, &capturing)) != cudaSuccess) return e;
        cudaLaunchAttribute pdl_attr[1];
        pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(WARPS_PER_CTA * 32);
        cfg.stream = stream;
        cfg.attrs = pdl_attr;
        cfg.numAttrs = (chebyshev && pdl_on && capturing == cudaStreamCaptureStatusNone) ? 1 : 0;
        const int* const order_arg = sort ? order : nullptr;
        if (pack2) {
            long long pc = ((n + 1) / 2 + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
            if (pc > (long long)sms * per_sm_pack) pc = (long long)sms * per_sm_pack;
            cfg.gridDim = dim3((unsigned)pc);
            cfg.dynamicSmemBytes = pack_smem;
            if ((e = cudaLaunchKernelEx(&cfg, pack_kern, P, n, ts_, ex_, lo_, st_, counter, order_arg, (const double*)prep_ws, binned)) != cudaSuccess) {
                cudaMemsetAsync(counter, 0, sizeof(unsigned int) * (ORDER_HIST_OFF + ORDER_BINS), stream);      // no evaluator ran: it would have left the workspace zero
                return e;
            }
            continue;
        }
        long long ctas = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        // every model gets a warp at once and the SMs are at most half full: the latency variant
        static const bool lat_on = [] { const char* e = std::getenv("EML_LATENCY_VARIANT"); return !(e && e[0] == '0'); }();
        const bool lat = lat_on && v.eval_lat != nullptr && ctas <= (long long)sms * per_sm_lat;
        if (ctas > max_ctas) ctas = max_ctas;
        cfg.gridDim = dim3((unsigned)ctas);
        cfg.dynamicSmemBytes = v.smem;
        if ((e = cudaLaunchKernelEx(&cfg, lat ? v.eval_lat : v.eval, P, n, p_, ts_, ex_, lo_, st_, counter, order_arg, (const double*)prep_ws,
                                    binned)) != cudaSuccess) {
            cudaMemsetAsync(counter, 0, sizeof(unsigned int) * (ORDER_HIST_OFF + ORDER_BINS), stream);      // as above
            return e;
        }
    }
    return cudaSuccess;
}

}  // namespace eml
