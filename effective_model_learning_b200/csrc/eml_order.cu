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

__global__ void __launch_bounds__(16 * COST_MODELS_PER_CTA)
cheb16_cost_kernel(const DevProblem P, const int n_models, const double* __restrict__ params, float* __restrict__ cost,
                   const double zcap, const int kcap, const float fixed, const float per_application) {
    constexpr int N = 4, D = 16;
    __shared__ float Ma[COST_MODELS_PER_CTA][D][D + 1];
    __shared__ float Mb[COST_MODELS_PER_CTA][D][D + 1];
    __shared__ double jw[COST_MODELS_PER_CTA][32];
    __shared__ double kd[COST_MODELS_PER_CTA][8];
    const int ml = threadIdx.x >> 4, r = threadIdx.x & 15;
    const int b_raw = blockIdx.x * COST_MODELS_PER_CTA + ml;
    const bool valid = b_raw < n_models;
    const long long b = valid ? b_raw : n_models - 1;
    const double* par = params + b * P.n_params;

    // ---- one pass over the term list: row r of H (real: the plan guarantees real Hamiltonian terms; natural bit order,
    //      site s on bit N - 1 - s) and this thread's two real jump weights jw[site][type][ra][0..1] (the evaluator's
    //      layout; the type-0 threads also collect kd[site][ra], the diagonal of sum rate * A^dag A)
#pragma unroll
    for (int c = 0; c < D; ++c) Ma[ml][r][c] = 0.f;
    const int jbp = r >> 2, jtype = (r >> 1) & 1, jra = r & 1;
    const int ja = jtype ? 1 - jra : jra;
    double w0 = 0.0, w1 = 0.0, kdv = 0.0;
    for (int ti = 0; ti < P.n_terms; ++ti) {
        const EmlTerm tm = P.terms[ti];
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
    jw[ml][2 * r] = w0;
    jw[ml][2 * r + 1] = w1;
    if (jtype == 0) kd[ml][jbp * 2 + jra] = kdv;
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
    // ---- (H - mu)^8 by three squarings: thread r owns row r
    float (*src)[D + 1] = Ma[ml];
    float (*dst)[D + 1] = Mb[ml];
    for (int sq = 0; sq < 3; ++sq) {
        float a[D], acc[D];
#pragma unroll
        for (int k = 0; k < D; ++k) { a[k] = src[r][k]; acc[k] = 0.f; }
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = fmaf(a[k], src[k][c], acc[c]);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < D; ++c) dst[r][c] = acc[c];
        __syncwarp();
        float (*tmp)[D + 1] = src; src = dst; dst = tmp;
    }
    float cs = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) cs += fabsf(src[k][r]);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) cs = fmaxf(cs, __shfl_xor_sync(0xffffffffu, cs, o));

    if (r == 0 && valid) {
        double spread = (double)hi - (double)lo;
        spread = fmin(spread, 2.0 * sqrt(sqrt(sqrt((double)cs))) * (1.0 + 1e-6));
        double dlo, dhi, dasym;
        cheb_site_ranges(jw[ml], kd[ml], N, dlo, dhi, dasym);
        const int T = P.n_times;
        const double* times = P.times;
        const ChebPlan plan = cheb_make_plan(spread, dlo, dhi, dasym, times[T - 1] - times[0], zcap, kcap);
        // macro-step scan of the evaluator (the step ends are found by bisection; after 32 steps the rest is extrapolated)
        double napp = 0.0;
        int kidx = 0, steps = 0;
        while (kidx < T - 1 && steps < 32) {
            const double t0 = times[kidx];
            int lo_i = kidx, hi_i = T - 1;      // largest index with times[i] - t0 <= tau_max
            while (lo_i < hi_i) {
                const int mid = (lo_i + hi_i + 1) >> 1;
                if (times[mid] - t0 <= plan.tau_max) lo_i = mid; else hi_i = mid - 1;
            }
            int q = lo_i - kidx, nsub = 1;
            if (q == 0) {
                const double ns = ceil((times[kidx + 1] - t0) / plan.tau_max);
                nsub = (ns >= 1.0 && ns < 1e6) ? (int)ns : 1;
                q = 1;
            }
            const double tau_end = (times[kidx + q] - t0) / nsub;
            if (tau_end * plan.R >= 1e-40) {
                int Kc = cheb_terms(tau_end * plan.S);
                if (!(Kc <= kcap)) Kc = kcap;
                if (Kc < 0) Kc = 0;
                napp += (double)nsub * (double)Kc;
            }
            kidx += q;
            ++steps;
        }
        if (kidx < T - 1 && kidx > 0) napp *= (double)(T - 1) / (double)kidx;
        float c = fixed + per_application * (float)napp;
        if (!(c >= 0.f && c < 1e30f)) c = 0.f;      // non-finite models are rejected by the evaluator at once
        cost[b] = c;
    }
}

// One CTA.  order_ws: [0, n) models sorted by cost (longest first), [ORDER_KEYS_AT, +n) start keys; cost (dead after the
// sort) is overwritten by the packed list `out`.
constexpr int ORDER_KEYS_AT = 1 << 17;

__global__ void __launch_bounds__(1024)
pack_models_kernel(const int n_models, const float* cost, const int slots, const float eps, int* order_ws, int* out) {
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
    if (tid == 0) pack_classes(n_class, slots, eps, tab, true);
    if (tid >= 32)
        for (int b = tid - 32; b < n_models; b += blockDim.x - 32) {
            const unsigned int slot = atomicAdd(&start[pack_bin_of(cost[b], scale)], 1u);
            sorted[slot] = b;
        }
    __syncthreads();
    // start key of every position of the sorted list; histogram of the keys
    for (int j = 0; j < PACK_CLASSES; ++j)
        for (int p = s_class[j] + tid; p < s_class[j + 1]; p += blockDim.x) {
            const int key = pack_key_of(tab, j, p - s_class[j]);
            keys[p] = key;
            atomicAdd(&hist2[key], 1u);
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
    for (int p = tid; p < n_models; p += blockDim.x) {
        const unsigned int slot = atomicAdd(&hist2[keys[p]], 1u);
        out[slot] = sorted[p];
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
    pack_models_kernel<<<1, 1024, 0, stream>>>(n, cost, slots, 0.05f, order_ws, out);
    return out;
}

bool packed_order_applicable(const DevProblem& P, long long n_models, int slots) {
    return P.n_tls == 4 && P.n_times >= 1 && n_models > slots && n_models <= 8LL * slots && n_models <= PACK_MAX_MODELS;
}

}  // namespace eml
