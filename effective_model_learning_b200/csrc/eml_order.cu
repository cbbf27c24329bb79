// Pre-pass of the d = 16 Chebyshev evaluator: EXACT cost of every model of the launch and a packed model list.
//
//   cheb16_cost_kernel   16 threads per model repeat, in FP32, what the evaluator's set-up does to choose its series:
//                        H from the parameter row, Gershgorin and 2 ||(H - mu)^8||_1^(1/8) spread bounds, per-site
//                        dissipator ranges, cheb_make_plan, the macro-step scan -- and so obtain the number of generator
//                        applications the evaluator will spend (the FP32 bound differs from the FP64 one by ~1e-6, i.e.
//                        by at most a term).  cost = fixed + (1 + output share) * applications, in applications.
//   pack_models_kernel   one CTA: counting sort by cost, best-fit-decreasing packing of the cost classes onto the
//                        evaluator's warp slots (eml_pack.cuh), models listed by their start time inside their slot.
//
// Nothing here touches the numerics of an evaluation: only the ORDER in which the persistent warps fetch the models
// changes (tests/test_gpu_fullsize.py checks that a model's result is independent of its position in the batch).
#include <cmath>
#include "eml_common.cuh"
#include "eml_cheb.cuh"
#include "eml_pack.cuh"

namespace eml {

constexpr int COST_MODELS_PER_CTA = 8;      // 16 threads per model
constexpr int COST_MAXP = 128;              // = MAXP of the warp kernel (plans with longer rows never reach it)
constexpr int COST_MAX_TERMS = 192;         // term lists up to this length are staged in shared memory
constexpr int COST_LD = 20;                 // row stride (floats) of the 16 x 16 matrices: rows are 16-byte aligned

// FP32 copies of cheb_site_ranges / cheb_make_plan (eml_cheb.cuh): the same formulas, the plan is only used to COUNT terms
__device__ __forceinline__ void cheb_site_ranges_f(const float* jw, const float* kd, const int n_sites, float& dlo, float& dhi,
                                                   float& dasym) {
    dlo = 0.f; dhi = 0.f; dasym = 0.f;
    for (int bp = 0; bp < n_sites; ++bp) {
        float slo = 3e38f, shi = -3e38f, sas = 0.f;
#pragma unroll
        for (int rc = 0; rc < 4; ++rc) {
            const int ra = rc >> 1, ca = rc & 1;
            const float dg = jw[bp * 8 + rc] - 0.5f * (kd[bp * 2 + ra] + kd[bp * 2 + ca]);
            const float fa = jw[bp * 8 + 4 + rc], fb = jw[bp * 8 + 4 + (3 - rc)];
            const float rad = 0.5f * fabsf(fa + fb);
            slo = fminf(slo, dg - rad); shi = fmaxf(shi, dg + rad); sas = fmaxf(sas, 0.5f * fabsf(fa - fb));
        }
        dlo += slo; dhi += shi; dasym += sas;
    }
}

struct ChebPlanF { float R, S, tau_max; };
__device__ __forceinline__ ChebPlanF cheb_make_plan_f(const float spread, const float dlo, const float dhi, const float dasym,
                                                      const float span, const float zcap, const int kcap) {
    const float ah = 0.5f * (dhi - dlo), yext = spread + dasym;
    float Rb = 1.f, Sb = 3e38f, Bb = 0.f;
    bool have_ok = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float delta = (i == 0) ? 0.03f : (i == 1) ? 0.06f : (i == 2) ? 0.12f : (i == 3) ? 0.25f : (i == 4) ? 0.5f : 1.0f;
        const float Rc = fmaxf(fmaxf(yext * (1.f + delta), ah), 1e-3f / fmaxf(span, 1e-30f));
        const float cc = Rc * Rc - ah * ah - yext * yext, xr2 = ah * ah * Rc * Rc;
        const float disc = sqrtf(cc * cc + 4.f * xr2);
        const float u = (cc > 0.f) ? 2.f * xr2 / (cc + disc) : 0.5f * (disc - cc);
        const float Bc = sqrtf(u), Ac = sqrtf(u + Rc * Rc);
        const bool ok = span * Bc <= (float)CHEB_HUMP;
        if ((ok && !have_ok) || (ok == have_ok && Ac + Bc < Sb)) { Rb = Rc; Sb = Ac + Bc; Bb = Bc; have_ok = ok; }
    }
    ChebPlanF pl;
    pl.R = Rb; pl.S = Sb;
    float tau_max = zcap / Sb;
    if (Bb > 0.f) tau_max = fminf(tau_max, (float)CHEB_HUMP / Bb);
    const float lnr = logf(Sb / Rb);
    if (lnr * (float)kcap > 560.f) {
        const float kall = 560.f / lnr;
        tau_max = fminf(tau_max, fmaxf(kall - (float)CHEB_C1 * cbrtf(kall) - (float)CHEB_C0, 1.f) / Sb);
    }
    pl.tau_max = tau_max;
    return pl;
}
__device__ __forceinline__ int cheb_terms_f(const float z) { return (int)ceilf(z + (float)CHEB_C1 * cbrtf(z) + (float)CHEB_C0); }

__global__ void __launch_bounds__(16 * COST_MODELS_PER_CTA)
cheb16_cost_kernel(const DevProblem P, const int n_models, const double* __restrict__ params, float* __restrict__ cost,
                   const double zcap, const int kcap, const float fixed, const float per_application) {
    constexpr int N = 4, D = 16;
    __shared__ __align__(16) float Ma[COST_MODELS_PER_CTA][D][COST_LD];
    __shared__ __align__(16) float Mb[COST_MODELS_PER_CTA][D][COST_LD];
    __shared__ double spar[COST_MODELS_PER_CTA * COST_MAXP];
    __shared__ EmlTerm sterms[COST_MAX_TERMS];
    __shared__ float jw[COST_MODELS_PER_CTA][32];
    __shared__ float kd[COST_MODELS_PER_CTA][8];
    const int ml = threadIdx.x >> 4, r = threadIdx.x & 15;
    const long long b0 = (long long)blockIdx.x * COST_MODELS_PER_CTA;
    const bool valid = b0 + ml < n_models;
    const int np = P.n_params;

    // ---- stage the parameter rows of the CTA's models (contiguous in HBM) and the term list
    {
        const long long rows = (n_models - b0 < COST_MODELS_PER_CTA) ? n_models - b0 : COST_MODELS_PER_CTA;
        for (int i = threadIdx.x; i < (int)rows * np; i += blockDim.x) spar[i] = params[b0 * np + i];
        if (P.n_terms <= COST_MAX_TERMS)
            for (int i = threadIdx.x; i < P.n_terms; i += blockDim.x) sterms[i] = P.terms[i];
    }
    __syncthreads();
    const EmlTerm* const terms = (P.n_terms <= COST_MAX_TERMS) ? sterms : P.terms;
    const double* const par = spar + (valid ? ml : 0) * np;      // rows past the end of the batch reuse row 0 (never written back)

    // ---- one pass over the term list: row r of H (real: the plan guarantees real Hamiltonian terms; natural bit order,
    //      site s on bit N - 1 - s) and this thread's two real jump weights jw[site][type][ra][0..1] (the evaluator's
    //      layout; the type-0 threads also collect kd[site][ra], the diagonal of sum rate * A^dag A)
#pragma unroll
    for (int c = 0; c < D; ++c) Ma[ml][r][c] = 0.f;
    const int jbp = r >> 2, jtype = (r >> 1) & 1, jra = r & 1;
    const int ja = jtype ? 1 - jra : jra;
    double w0 = 0.0, w1 = 0.0, kdv = 0.0;
    for (int ti = 0; ti < P.n_terms; ++ti) {
        const EmlTerm tm = terms[ti];
        const double v = par[tm.param];
        if (v == 0.0) continue;
        if (tm.kind == EML_TERM_LINDBLAD) {
            if (tm.site_a != jbp) continue;
            const cplx x = ld_op(P.op_table, tm.op_a, jra, ja);
            w0 += v * cmul(x, cconj(ld_op(P.op_table, tm.op_a, 0, jtype ? 1 : 0))).re;
            w1 += v * cmul(x, cconj(ld_op(P.op_table, tm.op_a, 1, jtype ? 0 : 1))).re;
            if (jtype == 0) {
                const cplx m0 = ld_op(P.op_table, tm.op_a, 0, jra), m1 = ld_op(P.op_table, tm.op_a, 1, jra);
                kdv += v * ((m0.re * m0.re + m0.im * m0.im) + (m1.re * m1.re + m1.im * m1.im));
            }
            continue;
        }
        const int sa = N - 1 - tm.site_a;
        const int ra = (r >> sa) & 1;
        if (tm.kind == EML_TERM_ENERGY) {
            for (int a = 0; a < 2; ++a) {
                const int c = (r & ~(1 << sa)) | (a << sa);
                Ma[ml][r][c] += (float)(v * ld_op(P.op_table, tm.op_a, ra, a).re);
            }
        } else {
            const int sb = N - 1 - tm.site_b;
            const int rb = (r >> sb) & 1;
            const double v2 = v * v;
            for (int a = 0; a < 2; ++a)
                for (int bb = 0; bb < 2; ++bb) {
                    const int c = (r & ~((1 << sa) | (1 << sb))) | (a << sa) | (bb << sb);
                    Ma[ml][r][c] += (float)(v2 * cmul(ld_op(P.op_table, tm.op_a, ra, a), ld_op(P.op_table, tm.op_b, rb, bb)).re);
                }
        }
    }
    jw[ml][2 * r] = (float)w0;
    jw[ml][2 * r + 1] = (float)w1;
    if (jtype == 0) kd[ml][jbp * 2 + jra] = (float)kdv;
    __syncwarp();

    // ---- Gershgorin spread and the mean of the diagonal (reductions over the 16 lanes of the model)
    float rad = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c)
        if (c != r) rad += fabsf(Ma[ml][r][c]);
    const float ctr = Ma[ml][r][r];
    float hi = ctr + rad, lo = ctr - rad, mu = ctr;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        mu += __shfl_xor_sync(0xffffffffu, mu, o);
    }
    mu *= 1.f / D;
    Ma[ml][r][r] = ctr - mu;
    __syncwarp();
    // ---- (H - mu)^8 by three squarings: thread r owns row r (the other rows are read as float4, broadcast to the 16 lanes)
    float (*src)[COST_LD] = Ma[ml];
    float (*dst)[COST_LD] = Mb[ml];
    for (int sq = 0; sq < 3; ++sq) {
        float acc[D];
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float a = src[r][k];
#pragma unroll
            for (int c4 = 0; c4 < D; c4 += 4) {
                const float4 s4 = *reinterpret_cast<const float4*>(&src[k][c4]);
                acc[c4] = fmaf(a, s4.x, acc[c4]); acc[c4 + 1] = fmaf(a, s4.y, acc[c4 + 1]);
                acc[c4 + 2] = fmaf(a, s4.z, acc[c4 + 2]); acc[c4 + 3] = fmaf(a, s4.w, acc[c4 + 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < D; ++c) dst[r][c] = acc[c];
        __syncwarp();
        float (*tmp)[COST_LD] = src; src = dst; dst = tmp;
    }
    float cs = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) cs += fabsf(src[k][r]);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) cs = fmaxf(cs, __shfl_xor_sync(0xffffffffu, cs, o));

    if (r == 0 && valid) {
        float spread = hi - lo;
        spread = fminf(spread, 2.f * sqrtf(sqrtf(sqrtf(cs))) * (1.f + 1e-6f));
        float dlo, dhi, dasym;
        cheb_site_ranges_f(jw[ml], kd[ml], N, dlo, dhi, dasym);
        const int T = P.n_times;
        const double* times = P.times;
        const ChebPlanF plan = cheb_make_plan_f(spread, dlo, dhi, dasym, (float)(times[T - 1] - times[0]), (float)zcap, kcap);
        // macro-step scan of the evaluator (the step ends are found by bisection; after 32 steps the rest is extrapolated)
        float napp = 0.f;
        int kidx = 0, steps = 0;
        while (kidx < T - 1 && steps < 32) {
            const double t0 = times[kidx];
            int lo_i = kidx, hi_i = T - 1;      // largest index with times[i] - t0 <= tau_max
            if (!((float)(times[hi_i] - t0) <= plan.tau_max))
                while (lo_i < hi_i) {
                    const int mid = (lo_i + hi_i + 1) >> 1;
                    if ((float)(times[mid] - t0) <= plan.tau_max) lo_i = mid; else hi_i = mid - 1;
                }
            else
                lo_i = hi_i;
            int q = lo_i - kidx, nsub = 1;
            if (q == 0) {
                const float ns = ceilf((float)(times[kidx + 1] - t0) / plan.tau_max);
                nsub = (ns >= 1.f && ns < 1e6f) ? (int)ns : 1;
                q = 1;
            }
            const float tau_end = (float)(times[kidx + q] - t0) / (float)nsub;
            if (tau_end * plan.R >= 1e-30f) {
                int Kc = cheb_terms_f(tau_end * plan.S);
                if (!(Kc <= kcap)) Kc = kcap;
                if (Kc < 0) Kc = 0;
                napp += (float)nsub * (float)Kc;
            }
            kidx += q;
            ++steps;
        }
        if (kidx < T - 1 && kidx > 0) napp *= (float)(T - 1) / (float)kidx;
        float c = fixed + per_application * napp;
        if (!(c >= 0.f && c < 1e30f)) c = 0.f;      // non-finite models are rejected by the evaluator at once
        cost[b0 + ml] = c;
    }
}

// One CTA.  order_ws: [0, n) models sorted by cost (longest first), [ORDER_KEYS_AT, +n) start keys; cost (dead after the
// sort) is overwritten by the packed list `out`.
constexpr int ORDER_KEYS_AT = 1 << 17;

__global__ void __launch_bounds__(1024)
pack_models_kernel(const int n_models, const float* cost, const int slots, const float eps, const float tail_fraction,
                   int* order_ws, int* out) {
    __shared__ unsigned int hist[PACK_BINS];
    __shared__ unsigned int start[PACK_BINS];
    __shared__ int n_class[PACK_CLASSES];
    __shared__ int s_class[PACK_CLASSES + 1];
    __shared__ PackTables tab;
    __shared__ unsigned int hist2[PACK_LEVELS + 1];
    __shared__ float smax;
    const int tid = threadIdx.x;
    int* const sorted = order_ws;
    int* const keys = order_ws + ORDER_KEYS_AT;
    if (tid < PACK_BINS) hist[tid] = 0;
    for (int i = tid; i < PACK_LEVELS + 1; i += blockDim.x) hist2[i] = 0;
    for (int i = tid; i < PACK_LEVELS; i += blockDim.x) tab.bins_at[i] = 0;
    if (tid < PACK_MASK_WORDS) tab.mask[tid] = 0u;
    if (tid == 0) smax = 0.f;
    __syncthreads();
    float mymax = 0.f;
    for (int b = tid; b < n_models; b += blockDim.x) mymax = fmaxf(mymax, cost[b]);
    atomicMax((int*)&smax, __float_as_int(mymax));      // non-negative floats order like ints
    __syncthreads();
    const float scale = (smax > 0.f) ? (PACK_BINS - 1) / smax : 0.f;
    for (int b = tid; b < n_models; b += blockDim.x) atomicAdd(&hist[pack_bin_of(cost[b], scale)], 1u);
    __syncthreads();
    if (tid < 32) {     // exclusive scan of the 256 bins: 8 bins per lane, then a shuffle scan of the lane totals
        unsigned int loc[PACK_BINS / 32], tot = 0;
#pragma unroll
        for (int i = 0; i < PACK_BINS / 32; ++i) { loc[i] = tot; tot += hist[tid * (PACK_BINS / 32) + i]; }
        unsigned int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += up;
        }
        const unsigned int base = inc - tot;
#pragma unroll
        for (int i = 0; i < PACK_BINS / 32; ++i) start[tid * (PACK_BINS / 32) + i] = base + loc[i];
    }
    __syncthreads();
    if (tid < PACK_CLASSES) {
        constexpr int PER = PACK_BINS / PACK_CLASSES;
        int c = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) c += (int)hist[tid * PER + i];
        n_class[tid] = c;
        s_class[tid] = (int)start[tid * PER];
    }
    if (tid == 0) s_class[PACK_CLASSES] = n_models;
    __syncthreads();
    // the packing (one thread) runs while the others scatter the models into the sorted list
    if (tid == 0) pack_classes(n_class, slots, eps, tab, true, true, tail_fraction);
    if (tid >= 32)
        for (int b = tid - 32; b < n_models; b += blockDim.x - 32) {
            const unsigned int slot = atomicAdd(&start[pack_bin_of(cost[b], scale)], 1u);
            sorted[slot] = b;
        }
    __syncthreads();
    // start key of every position of the sorted list; histogram of the keys.  Neighbouring positions mostly share a key
    // (1776 of the 4096 models of the bench batch have key 0), so the counts are aggregated per warp before the atomic.
    const int lane = tid & 31;
    for (int p0 = 0; p0 < n_models; p0 += blockDim.x) {
        const int p = p0 + tid;
        int key = -1;
        if (p < n_models) {
            int j = 0;      // class of position p: largest j with s_class[j] <= p
#pragma unroll
            for (int step = PACK_CLASSES / 2; step > 0; step >>= 1)
                if (s_class[j + step] <= p) j += step;
            key = pack_key_of(tab, j, p - s_class[j]);
            keys[p] = key;
        }
        const unsigned int grp = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs((int)grp) - 1) atomicAdd(&hist2[key], (unsigned int)__popc(grp));
    }
    __syncthreads();
    if (tid < 32) {     // exclusive scan of the cap + 2 key counts (in place): 33 entries per lane
        constexpr int PER = (PACK_LEVELS + 1 + 31) / 32;
        unsigned int tot = 0;
        for (int i = 0; i < PER; ++i) {
            const int k = tid * PER + i;
            if (k < PACK_LEVELS + 1) tot += hist2[k];
        }
        unsigned int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += up;
        }
        unsigned int run = inc - tot;
        for (int i = 0; i < PER; ++i) {
            const int k = tid * PER + i;
            if (k < PACK_LEVELS + 1) { const unsigned int h = hist2[k]; hist2[k] = run; run += h; }
        }
    }
    __syncthreads();
    for (int p0 = 0; p0 < n_models; p0 += blockDim.x) {
        const int p = p0 + tid;
        const int key = (p < n_models) ? keys[p] : -1;
        const unsigned int grp = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs((int)grp) - 1;
        unsigned int base = 0;
        if (key >= 0 && lane == leader) base = atomicAdd(&hist2[key], (unsigned int)__popc(grp));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (key >= 0) out[base + __popc(grp & ((1u << lane) - 1u))] = sorted[p];
    }
}

// Exact costs (or the caller's measured ones) -> packed list.  Returns the device pointer of the list inside order_ws.
// order_ws: 2 * 2^18 ints (the plan's ordering buffer).  slots: warp slots of the evaluator launch.
const int* launch_packed_order(const DevProblem& P, long long n_models, const double* params, int* order_ws,
                               const int* cost_hint, int slots, double zcap, int kcap, float fixed, float per_application,
                               cudaStream_t stream);

__global__ void __launch_bounds__(256) hint_cost_affine_kernel(const int n_models, const int* __restrict__ hint,
                                                               float* __restrict__ cost, const float fixed, const float per_application) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_models) cost[b] = fixed + per_application * (float)max(hint[b], 0);
}

const int* launch_packed_order(const DevProblem& P, long long n_models, const double* params, int* order_ws,
                               const int* cost_hint, int slots, double zcap, int kcap, float fixed, float per_application,
                               cudaStream_t stream) {
    float* cost = reinterpret_cast<float*>(order_ws + (1 << 18));
    const int n = (int)n_models;
    if (cost_hint != nullptr)
        hint_cost_affine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, cost_hint, cost, fixed, per_application);
    else
        cheb16_cost_kernel<<<(unsigned)((n + COST_MODELS_PER_CTA - 1) / COST_MODELS_PER_CTA), 16 * COST_MODELS_PER_CTA, 0, stream>>>(
            P, n, params, cost, zcap, kcap, fixed, per_application);
    int* out = order_ws + (1 << 18);      // the cost scratch is dead once the models are sorted
    // capacity 5 % above the mean load of the packed models; the cheapest 10 % of the models form the greedy tail
    pack_models_kernel<<<1, 1024, 0, stream>>>(n, cost, slots, 0.05f, 0.10f, order_ws, out);
    return out;
}

// Diagnostic: the generator applications cheb16_cost_kernel predicts for every model (what the packing is built on), to be
// compared with the counts the evaluator reports in `status`.  d = 16 plans with a real H only.
cudaError_t launch_predicted_applications(const DevProblem& P, long long n_models, const double* params, float* out,
                                          double zcap, int kcap, cudaStream_t stream) {
    const int n = (int)n_models;
    cheb16_cost_kernel<<<(unsigned)((n + COST_MODELS_PER_CTA - 1) / COST_MODELS_PER_CTA), 16 * COST_MODELS_PER_CTA, 0, stream>>>(
        P, n, params, out, zcap, kcap, 0.f, 1.f);
    return cudaGetLastError();
}

bool packed_order_applicable(const DevProblem& P, long long n_models, int slots) {
    return P.n_tls == 4 && P.n_times >= 1 && n_models > slots && n_models <= 8LL * slots && n_models <= PACK_MAX_MODELS;
}

}  // namespace eml
