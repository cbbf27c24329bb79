// Generic evaluator: one CTA per model, one thread per density-matrix element,
// rho / Taylor terms double-buffered in shared memory, FP64 FMA pipe.
// Works for every n_tls in 1..5 and every model the boundary can describe (non-Hermitian
// H or rho0 included).  It is the correctness baseline for the specialised kernels.
//
// Replaces, per model: build_Ls/build_H (reference basic_model.py:152-255), the qutip
// Liouvillian + ZVODE loop behind mesolve (basic_model.py:381-388), the e_ops traces and
// the loss of learning_chain.py:738-742.
#include "eml_common.cuh"

namespace eml {

struct GenericSmem {
    // offsets in doubles
    int G_re, G_im, Gt_re, Gt_im, S, v, red, rr, par, misc, total;
};

__host__ __device__ inline GenericSmem generic_layout(int n, int n_params) {
    const int D = 1 << n, DD = D * D;
    GenericSmem L;
    int o = 0;
    L.G_re = o; o += DD;
    L.G_im = o; o += DD;
    L.Gt_re = o; o += DD;
    L.Gt_im = o; o += DD;
    L.S = o; o += n * 32;
    L.v = o; o += 4 * DD;           // 2 buffers x (re plane, im plane)
    L.red = o; o += QMAX * 4 * D;   // [p][ab][rest][2], rest < D/2
    L.rr = o; o += QMAX * 8;        // [p][ab][2]
    L.par = o; o += n_params;
    L.misc = o; o += 2 * D + 8 + QMAX * MAX_OBS + 3 * QMAX;   // + xs[QMAX], xp[2][QMAX] (dense-output weights)
    L.total = o;
    return L;
}

size_t generic_smem_bytes(int n, int n_params) { return sizeof(double) * (size_t)generic_layout(n, n_params).total; }

template <int N>
__global__ void __launch_bounds__(((1 << (2 * N)) < 128) ? 128 : (1 << (2 * N)))
eval_generic_kernel(const DevProblem P, const long long n_models, const double* __restrict__ params,
                    const int* __restrict__ target_set, double* __restrict__ expect, double* __restrict__ loss,
                    int* __restrict__ status) {
    constexpr int D = 1 << N, DD = D * D;
    extern __shared__ double sm[];
    const GenericSmem L = generic_layout(N, P.n_params);
    double* G_re = sm + L.G_re;
    double* G_im = sm + L.G_im;
    double* Gt_re = sm + L.Gt_re;
    double* Gt_im = sm + L.Gt_im;
    double* S = sm + L.S;
    double* vbuf = sm + L.v;
    double* red = sm + L.red;
    double* rr = sm + L.rr;
    double* par = sm + L.par;
    double* misc = sm + L.misc;

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const long long b = blockIdx.x;
    if (b >= n_models) return;
    const int T = P.n_times, K = P.n_obs;

    // ---- load parameters, clear operator storage
    for (int i = tid; i < P.n_params; i += nthr) par[i] = params[b * P.n_params + i];
    for (int i = tid; i < 4 * DD; i += nthr) sm[L.G_re + i] = 0.0;   // G, Gt planes are contiguous
    for (int i = tid; i < N * 32; i += nthr) S[i] = 0.0;
    __syncthreads();

    // ---- K1: assemble G = -iH - K/2, Gt = +iH - K/2 (thread r owns row r) and the per-site
    //          jump superoperators S_k = sum rate * A (x) conj(A) (thread owns one entry)
    if (tid < D) {
        const int r = tid;
        for (int t = 0; t < P.n_terms; ++t) {
            const EmlTerm tm = P.terms[t];
            const double v = par[tm.param];
            if (v == 0.0) continue;
            const int sa = N - 1 - tm.site_a;
            const int ra = (r >> sa) & 1;
            if (tm.kind == EML_TERM_ENERGY) {
                for (int a = 0; a < 2; ++a) {
                    const cplx A = ld_op(P.op_table, tm.op_a, ra, a);
                    const int c = (r & ~(1 << sa)) | (a << sa);
                    const double hre = v * A.re, him = v * A.im;
                    G_re[r * D + c] += him;  G_im[r * D + c] -= hre;     // -i h
                    Gt_re[r * D + c] -= him; Gt_im[r * D + c] += hre;    // +i h
                }
            } else if (tm.kind == EML_TERM_COUPLING) {
                const int sb = N - 1 - tm.site_b;
                const int rb = (r >> sb) & 1;
                const double v2 = v * v;
                for (int a = 0; a < 2; ++a)
                    for (int bb = 0; bb < 2; ++bb) {
                        const cplx A = ld_op(P.op_table, tm.op_a, ra, a);
                        const cplx B = ld_op(P.op_table, tm.op_b, rb, bb);
                        const cplx AB = cmul(A, B);
                        const int c = (r & ~((1 << sa) | (1 << sb))) | (a << sa) | (bb << sb);
                        const double hre = v2 * AB.re, him = v2 * AB.im;
                        G_re[r * D + c] += him;  G_im[r * D + c] -= hre;
                        Gt_re[r * D + c] -= him; Gt_im[r * D + c] += hre;
                    }
            } else {  // Lindblad: K = rate * A^dag A on site_a
                for (int a = 0; a < 2; ++a) {
                    cplx M = {0.0, 0.0};
                    for (int z = 0; z < 2; ++z) {
                        const cplx m = cmul(cconj(ld_op(P.op_table, tm.op_a, z, ra)), ld_op(P.op_table, tm.op_a, z, a));
                        M.re += m.re; M.im += m.im;
                    }
                    const int c = (r & ~(1 << sa)) | (a << sa);
                    G_re[r * D + c] -= 0.5 * v * M.re;  G_im[r * D + c] -= 0.5 * v * M.im;
                    Gt_re[r * D + c] -= 0.5 * v * M.re; Gt_im[r * D + c] -= 0.5 * v * M.im;
                }
            }
        }
    }
    if (tid < N * 16) {
        const int site = tid >> 4, e = tid & 15;
        const int ra = (e >> 3) & 1, ca = (e >> 2) & 1, a = (e >> 1) & 1, bb = e & 1;
        double sre = 0.0, sim = 0.0;
        for (int t = 0; t < P.n_terms; ++t) {
            const EmlTerm tm = P.terms[t];
            if (tm.kind != EML_TERM_LINDBLAD || tm.site_a != site) continue;
            const double v = par[tm.param];
            if (v == 0.0) continue;
            const cplx x = cmul(ld_op(P.op_table, tm.op_a, ra, a), cconj(ld_op(P.op_table, tm.op_a, ca, bb)));
            sre += v * x.re; sim += v * x.im;
        }
        S[site * 32 + e * 2] = sre;
        S[site * 32 + e * 2 + 1] = sim;
    }
    __syncthreads();

    // ---- norm bound nu >= ||L||_1 (column-stacked): max col sum |G| + max row sum |Gt| + sum_k ||S_k||_1
    if (tid < D) {
        double cs = 0.0, rs = 0.0;
        for (int i = 0; i < D; ++i) {
            cs += hypot(G_re[i * D + tid], G_im[i * D + tid]);
            rs += hypot(Gt_re[tid * D + i], Gt_im[tid * D + i]);
        }
        misc[tid] = cs;
        misc[D + tid] = rs;
    }
    __syncthreads();
    if (tid == 0) {
        double mc = 0.0, mr = 0.0;
        double poison = 0.0;     // fmax drops NaNs: carry them into nu explicitly
        for (int i = 0; i < D; ++i) { mc = fmax(mc, misc[i]); mr = fmax(mr, misc[D + i]); poison += misc[i] + misc[D + i]; }
        double nu = mc + mr + 0.0 * poison;
        for (int s = 0; s < N; ++s) {
            double best = 0.0;
            for (int col = 0; col < 4; ++col) {
                double sum = 0.0;
                for (int row = 0; row < 4; ++row) sum += hypot(S[s * 32 + (row * 4 + col) * 2], S[s * 32 + (row * 4 + col) * 2 + 1]);
                best = fmax(best, sum);
            }
            nu += best;
        }
        misc[2 * D] = nu;
    }
    __syncthreads();
    const double nu = misc[2 * D];
    if (!(nu * (P.times[T - 1] - P.times[0]) < 1e5)) {         // non-finite generator or absurd step count: do not hang
        if (tid == 0) {
            if (loss != nullptr) loss[b] = nan("");
            if (status != nullptr) status[b] = -1;
        }
        return;
    }
    double* lossbuf = misc + 2 * D + 8;   // [QMAX*MAX_OBS] per-thread loss partials

    // ---- per-thread element
    const bool active = tid < DD;
    const int r = active ? tid / D : 0, c = active ? tid % D : 0;
    const int ob = N - 1 - P.obs_site;
    const int obsbit = 1 << ob;
    const bool diag = active && (((r ^ c) & ~obsbit) == 0);
    const int ab = (((r >> ob) & 1) << 1) | ((c >> ob) & 1);
    const int rest = (r & (obsbit - 1)) | ((r >> 1) & ~(obsbit - 1));

    double acc_re = 0.0, acc_im = 0.0;
    if (active) { acc_re = P.rho0[2 * tid]; acc_im = P.rho0[2 * tid + 1]; }
    // dense-output accumulators live in shared memory (`red`, owned by the diag threads) and the weights x^j in
    // xs_s / xp_s (double-buffered by term parity), which keeps the kernel at ~100 registers
    double* xs_s = misc + 2 * D + 8 + QMAX * MAX_OBS;
    double* xp_s = xs_s + QMAX;
    const int ybase = (ab * (D / 2) + rest) * 2;       // this diag thread's slot within red[p][...]
    const int ystride = 4 * (D / 2) * 2;
    double my_loss = 0.0;
    int n_apply = 0;
    bool capped = false;
    const int tset = (target_set != nullptr) ? target_set[b] : 0;
    const double* tgt = (P.targets != nullptr) ? P.targets + (size_t)tset * K * T * 2 : nullptr;

    // emits q outputs starting at time index kbase from the accumulators the diag threads keep in red[p][ab][rest]
    auto emit = [&](int q, int kbase) {
        __syncthreads();
        if (tid < q * 4) {
            double sre = 0.0, sim = 0.0;
            for (int i = 0; i < D / 2; ++i) { sre += red[(tid * (D / 2) + i) * 2]; sim += red[(tid * (D / 2) + i) * 2 + 1]; }
            rr[tid * 2] = sre; rr[tid * 2 + 1] = sim;
        }
        __syncthreads();
        if (tid < q * K) {
            const int p = tid / K, m = tid % K;
            // Tr(O rho_red) = sum_{a,b} O[a][b] rho_red[b][a];  rr index = a*2+b holds rho_red[a][b]
            double ere = 0.0, eim = 0.0;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int bb = 0; bb < 2; ++bb) {
                    const double ore = P.obs[m * 8 + (a * 2 + bb) * 2], oim = P.obs[m * 8 + (a * 2 + bb) * 2 + 1];
                    const double rre = rr[(p * 4 + bb * 2 + a) * 2], rim = rr[(p * 4 + bb * 2 + a) * 2 + 1];
                    ere += ore * rre - oim * rim;
                    eim += ore * rim + oim * rre;
                }
            const int k = kbase + p;
            if (expect != nullptr) {
                double* e = expect + (((size_t)b * K + m) * T + k) * 2;
                e[0] = ere; e[1] = eim;
            }
            if (tgt != nullptr) {
                const double dre = ere - tgt[((size_t)m * T + k) * 2], dim = eim - tgt[((size_t)m * T + k) * 2 + 1];
                my_loss += dre * dre + dim * dim;
            }
        }
        __syncthreads();
    };

    if (diag) { red[ybase] = acc_re; red[ybase + 1] = acc_im; }
    emit(1, 0);

    int cur = 0;
    int k = 0;
    while (k < T - 1) {
        int q, nsub;
        double htot;
        choose_step(P.times, T, k, nu, P.theta_target, q, nsub, htot);
        if (htot == 0.0) {               // repeated time points: the state does not move
            if (diag)
                for (int p = 0; p < q; ++p) { red[p * ystride + ybase] = acc_re; red[p * ystride + ybase + 1] = acc_im; }
            emit(q, k + 1);
            k += q;
            continue;
        }
        const double h = htot / nsub;
        const double t0 = P.times[k];
        __syncthreads();      // the previous emit is done reading red / rr
        if (tid < QMAX) xs_s[tid] = (tid < q) ? ((nsub == 1) ? (P.times[k + 1 + tid] - t0) / htot : 1.0) : 0.0;
        for (int sub = 0; sub < nsub; ++sub) {
            double* vr = vbuf + cur * 2 * DD;
            if (active) { vr[tid] = acc_re; vr[DD + tid] = acc_im; }
            if (diag && sub == nsub - 1)
                for (int p = 0; p < q; ++p) { red[p * ystride + ybase] = acc_re; red[p * ystride + ybase + 1] = acc_im; }
            __syncthreads();      // xs_s visible
            if (tid < QMAX) xp_s[tid] = xs_s[tid];          // buffer 0 = x^1, the weight of term 1
            __syncthreads();
            int small = 0;
            int j = 0;
            for (; j < MCAP; ++j) {
                const double* v_re = vbuf + cur * 2 * DD;
                const double* v_im = v_re + DD;
                double* w_re = vbuf + (cur ^ 1) * 2 * DD;
                double* w_im = w_re + DD;
                double ore = 0.0, oim = 0.0;
                if (active) {
                    // G v + v Gt
#pragma unroll 4
                    for (int i = 0; i < D; ++i) {
                        const double gre = G_re[r * D + i], gim = G_im[r * D + i];
                        const double are = v_re[i * D + c], aim = v_im[i * D + c];
                        ore += gre * are - gim * aim;
                        oim += gre * aim + gim * are;
                        const double bre = v_re[r * D + i], bim = v_im[r * D + i];
                        const double tre = Gt_re[i * D + c], tim = Gt_im[i * D + c];
                        ore += bre * tre - bim * tim;
                        oim += bre * tim + bim * tre;
                    }
                    // sum_c c v c^dag through the per-site 4x4 superoperators
#pragma unroll
                    for (int s = 0; s < N; ++s) {
                        const int sb = N - 1 - s;
                        const int row = (((r >> sb) & 1) << 1) | ((c >> sb) & 1);
                        const int r0 = r & ~(1 << sb), c0 = c & ~(1 << sb);
#pragma unroll
                        for (int col = 0; col < 4; ++col) {
                            const double sre = S[s * 32 + (row * 4 + col) * 2], sim = S[s * 32 + (row * 4 + col) * 2 + 1];
                            const int rr_ = r0 | ((col >> 1) << sb), cc_ = c0 | ((col & 1) << sb);
                            const double are = v_re[rr_ * D + cc_], aim = v_im[rr_ * D + cc_];
                            ore += sre * are - sim * aim;
                            oim += sre * aim + sim * are;
                        }
                    }
                    const double f = h / (double)(j + 1);
                    ore *= f; oim *= f;
                    w_re[tid] = ore; w_im[tid] = oim;
                    acc_re += ore; acc_im += oim;
                    if (diag && sub == nsub - 1) {
                        const double* xw = xp_s + (j & 1) * QMAX;      // x^(j+1), prepared before the last barrier
                        for (int p = 0; p < q; ++p) {
                            red[p * ystride + ybase] += xw[p] * ore;
                            red[p * ystride + ybase + 1] += xw[p] * oim;
                        }
                    }
                }
                if (tid < QMAX) xp_s[((j + 1) & 1) * QMAX + tid] = xp_s[(j & 1) * QMAX + tid] * xs_s[tid];   // x^(j+2) for the next term
                const int big = (fabs(ore) > P.tol) || (fabs(oim) > P.tol);
                const int anybig = __syncthreads_or(big);
                cur ^= 1;
                if (!anybig) {
                    if (++small >= 2) break;
                } else {
                    small = 0;
                }
            }
            n_apply += (j < MCAP) ? j + 1 : MCAP;
            if (j >= MCAP) capped = true;
        }
        emit(q, k + 1);
        k += q;
    }

    // ---- K5: loss = sum |expect - target|^2 / (T K), fixed summation order
    if (tid < QMAX * MAX_OBS) lossbuf[tid] = my_loss;
    __syncthreads();
    if (tid == 0) {
        if (loss != nullptr) {
            double s = 0.0;
            const int nlb = (QMAX * K < nthr) ? QMAX * K : nthr;
            for (int i = 0; i < nlb; ++i) s += lossbuf[i];
            loss[b] = (tgt != nullptr) ? s / ((double)T * (double)K) : 0.0;
        }
        if (status != nullptr) status[b] = capped ? -1 : n_apply;
    }
}

#define EML_INST(N) template __global__ void eval_generic_kernel<N>(const DevProblem, const long long, const double*, const int*, double*, double*, int*);
EML_INST(1) EML_INST(2) EML_INST(3) EML_INST(4) EML_INST(5)

cudaError_t launch_generic(const DevProblem& P, long long n_models, const double* params, const int* target_set,
                           double* expect, double* loss, int* status, cudaStream_t stream) {
    const int n = P.n_tls;
    const int dd = 1 << (2 * n);
    const int threads = dd < 128 ? 128 : dd;   // >= QMAX*MAX_OBS so every output slot has a thread
    const size_t smem = generic_smem_bytes(n, P.n_params);
    cudaError_t err = cudaSuccess;
#define EML_LAUNCH(N)                                                                                              \
    case N:                                                                                                        \
        err = cudaFuncSetAttribute(eval_generic_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (err != cudaSuccess) return err;                                                                        \
        eval_generic_kernel<N><<<(unsigned)n_models, threads, smem, stream>>>(P, n_models, params, target_set,     \
                                                                              expect, loss, status);               \
        break;
    switch (n) {
        EML_LAUNCH(1) EML_LAUNCH(2) EML_LAUNCH(3) EML_LAUNCH(4) EML_LAUNCH(5)
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace eml
