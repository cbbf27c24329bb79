// extern "C" boundary of libeml_b200.so (see include/eml_b200.h).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "eml_common.cuh"

namespace eml {
cudaError_t launch_generic(const DevProblem& P, long long n_models, const double* params, const int* target_set,
                           double* expect, double* loss, int* status, cudaStream_t stream);
size_t generic_smem_bytes(int n, int n_params);
cudaError_t launch_warp_mma(const DevProblem& P, long long n_models, const double* params, const int* target_set,
                            double* expect, double* loss, int* status, unsigned int* counter, int* order, bool real_h, bool chebyshev,
                            bool parity_h, bool sector_b, const int* cost_hint, double* prep_ws, long long prep_cap, bool pack2_ok,
                            bool pdl_ok, cudaStream_t stream);
size_t warp_prep_bytes_per_model(const DevProblem& P, bool real_h, bool chebyshev, bool parity_h, bool sector_b);
bool jump_op_is_diag_flip_real(const double* op);
cudaError_t launch_cta_mma(const DevProblem& P, long long n_models, const double* params, const int* target_set,
                           double* expect, double* loss, int* status, unsigned int* counter, int* order, bool real_h,
                           bool chebyshev, bool parity_h, const int* cost_hint, double* prep_ws, long long prep_cap, cudaStream_t stream);
bool cta_mma_supported(const DevProblem& P, bool hermitian);
size_t cta_orbit_prep_bytes_per_model(bool real_h);
bool cta_orbit_selected();
bool warp_mma_supported(const DevProblem& P, bool hermitian);
cudaError_t measure_fp64(int use_mma, double seconds, double* flops_per_s);
cudaError_t read_phase_clocks(unsigned long long* out, bool reset);
}  // namespace eml

#include "eml_plan.hpp"

static std::atomic<long long> g_launches{0};
std::string& eml_last_error_ref() { static thread_local std::string e; return e; }


template <typename T>
static int upload(EmlPlan* plan, const T* host, size_t count, const T** out) {
    *out = nullptr;
    if (count == 0) return EML_OK;
    void* d = nullptr;
    EML_CUDA(cudaMalloc(&d, count * sizeof(T)));
    plan->owned.push_back(d);
    EML_CUDA(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)d;
    return EML_OK;
}

static bool op_is_hermitian(const double* op) {
    // [[a,b],[c,d]] row-major complex
    const double tol = 1e-15;
    return std::fabs(op[1]) <= tol && std::fabs(op[7]) <= tol && std::fabs(op[2] - op[4]) <= tol &&
           std::fabs(op[3] + op[5]) <= tol;
}

// The generator assembly of the warp kernels as a linear program (see eml_common.cuh): column p of the map
// c -> (G_re, G_im, jw, kd) is what the term-by-term assembly (eml_warp_mma.cu: assemble_generator, reference
// basic_model.py:152-255) produces for the unit vector e_p; rows are listed longest first so that the lanes of a warp,
// which take 32 consecutive rows at a time, run gathers of similar length.
static int compile_linear_program(EmlPlan* plan, const EmlProblemDesc* d) {
    const int N = plan->Pw.n_tls, D = 1 << N, obs_site = plan->Pw.obs_site, P = d->n_params;
    if (P > eml::WARP_MAXP) return EML_OK;
    std::vector<int> sq(P, -1);
    for (int t = 0; t < d->n_terms; ++t) {
        const int want = d->terms[t].kind == EML_TERM_COUPLING ? 1 : 0;
        if (sq[d->terms[t].param] >= 0 && sq[d->terms[t].param] != want) return EML_OK;   // a column shared by a coupling and a linear term
        sq[d->terms[t].param] = want;
    }
    for (int& v : sq) if (v < 0) v = 0;
    auto pos = [&](int s) { return (s == obs_site) ? N - 1 : (s < obs_site ? N - 2 - s : N - 1 - s); };
    struct C2 { double re, im; };
    auto op = [&](int o, int r, int c) { const double* p = d->op_table + o * 8 + (r * 2 + c) * 2; return C2{p[0], p[1]}; };
    auto mul = [](C2 a, C2 b) { return C2{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; };
    const int n_out = 2 * D * (D + 1) + eml::WARP_MAXP + 40;
    std::vector<std::vector<eml::LinEntry>> rows(n_out);
    std::vector<double> y(n_out);
    for (int p = 0; p < P; ++p) {
        std::fill(y.begin(), y.end(), 0.0);
        for (int t = 0; t < d->n_terms; ++t) {
            const EmlTerm& tm = d->terms[t];
            if (tm.param != p) continue;
            const int sa = pos(tm.site_a);
            for (int r = 0; r < D; ++r) {
                const int ra = (r >> sa) & 1;
                if (tm.kind == EML_TERM_ENERGY) {
                    for (int a = 0; a < 2; ++a) {
                        const C2 A = op(tm.op_a, ra, a);
                        const int c = (r & ~(1 << sa)) | (a << sa);
                        y[eml::prep_off_gre(D, r, c)] += A.im;      // -i h
                        y[eml::prep_off_gim(D, r, c)] -= A.re;
                    }
                } else if (tm.kind == EML_TERM_COUPLING) {
                    const int sb = pos(tm.site_b), rb = (r >> sb) & 1;
                    for (int a = 0; a < 2; ++a)
                        for (int bb = 0; bb < 2; ++bb) {
                            const C2 AB = mul(op(tm.op_a, ra, a), op(tm.op_b, rb, bb));
                            const int c = (r & ~((1 << sa) | (1 << sb))) | (a << sa) | (bb << sb);
                            y[eml::prep_off_gre(D, r, c)] += AB.im;
                            y[eml::prep_off_gim(D, r, c)] -= AB.re;
                        }
                } else {
                    for (int a = 0; a < 2; ++a) {
                        C2 M{0.0, 0.0};
                        for (int z = 0; z < 2; ++z) {
                            const C2 az = op(tm.op_a, z, ra);
                            const C2 m = mul(C2{az.re, -az.im}, op(tm.op_a, z, a));
                            M.re += m.re; M.im += m.im;
                        }
                        const int c = (r & ~(1 << sa)) | (a << sa);
                        y[eml::prep_off_gre(D, r, c)] -= 0.5 * M.re;
                        y[eml::prep_off_gim(D, r, c)] -= 0.5 * M.im;
                    }
                }
            }
            if (tm.kind == EML_TERM_LINDBLAD) {
                for (int lane = 0; lane < 32; ++lane) {
                    const int bp = lane >> 3, type = (lane >> 2) & 1, ra = (lane >> 1) & 1, ca = lane & 1;
                    if (bp >= N || sa != bp) continue;
                    const int a = type ? 1 - ra : ra, bb = type ? 1 - ca : ca;
                    const C2 cb = op(tm.op_a, ca, bb);
                    y[eml::prep_off_jw(D, lane)] += mul(op(tm.op_a, ra, a), C2{cb.re, -cb.im}).re;
                    if (type == 0 && ca == 0) {
                        const C2 m0 = op(tm.op_a, 0, ra), m1 = op(tm.op_a, 1, ra);
                        y[eml::prep_off_kd(D, bp * 2 + ra)] += (m0.re * m0.re + m0.im * m0.im) + (m1.re * m1.re + m1.im * m1.im);
                    }
                }
            }
        }
        for (int o = 0; o < n_out; ++o)
            if (y[o] != 0.0) rows[o].push_back(eml::LinEntry{y[o], p, 0});
    }
    std::vector<int> idx;
    for (int o = 0; o < n_out; ++o) if (!rows[o].empty()) idx.push_back(o);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return rows[a].size() > rows[b].size(); });
    std::vector<eml::LinRow> lrow;
    std::vector<eml::LinEntry> lent;
    for (int o : idx) {
        lrow.push_back(eml::LinRow{o, (int)lent.size(), (int)(lent.size() + rows[o].size()), 0});
        lent.insert(lent.end(), rows[o].begin(), rows[o].end());
    }
    if (lrow.empty()) { lrow.push_back(eml::LinRow{0, 0, 0, 0}); lent.push_back(eml::LinEntry{0.0, 0, 0}); }
    int rc;
    const eml::LinRow* d_row = nullptr; const eml::LinEntry* d_ent = nullptr; const int* d_sq = nullptr;
    if ((rc = upload(plan, lrow.data(), lrow.size(), &d_row)) != EML_OK) return rc;
    if ((rc = upload(plan, lent.data(), lent.size(), &d_ent)) != EML_OK) return rc;
    if ((rc = upload(plan, sq.data(), sq.size() ? sq.size() : 1, &d_sq)) != EML_OK) return rc;
    for (eml::DevProblem* Q : {&plan->P, &plan->Pw}) {
        Q->lin_rows = (int)idx.size(); Q->lin_nnz = (int)lent.size(); Q->lin_row = d_row; Q->lin_ent = d_ent; Q->lin_sq = d_sq;
    }
    return EML_OK;
}

extern "C" {

int eml_version(void) { return EML_VERSION_MAJOR * 1000 + EML_VERSION_MINOR; }
const char* eml_last_error_string(void) { return eml_last_error_ref().c_str(); }
int64_t eml_launch_count(void) { return (int64_t)g_launches.load(); }
int eml_debug_phase_clocks(uint64_t* out8, int reset) {
    if (!out8) return fail(EML_ERR_INVALID_ARGUMENT, "out8 is NULL");
    unsigned long long tmp[8];
    EML_CUDA(eml::read_phase_clocks(tmp, reset != 0));
    for (int i = 0; i < 8; ++i) out8[i] = tmp[i];
    return EML_OK;
}

int eml_device_count(int* count) {
    if (!count) return fail(EML_ERR_INVALID_ARGUMENT, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(EML_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    return EML_OK;
}

int eml_plan_destroy(EmlPlan* plan) {
    if (!plan) return EML_OK;
    cudaSetDevice(plan->device);
    for (void* p : plan->owned) cudaFree(p);
    if (plan->d_targets) cudaFree(plan->d_targets);
    plan->drop_host_graphs();
    if (plan->d_prep) cudaFree(plan->d_prep);
    cudaFree(plan->d_params); cudaFree(plan->d_expect); cudaFree(plan->d_loss); cudaFree(plan->d_tset); cudaFree(plan->d_status);
    cudaFreeHost(plan->h_params); cudaFreeHost(plan->h_expect); cudaFreeHost(plan->h_loss); cudaFreeHost(plan->h_tset); cudaFreeHost(plan->h_status);
    if (plan->stream) cudaStreamDestroy(plan->stream);
    delete plan;
    return EML_OK;
}

int eml_plan_set_targets(EmlPlan* plan, int32_t n_target_sets, const double* targets) {
    if (!plan) return fail(EML_ERR_INVALID_ARGUMENT, "plan is NULL");
    if (n_target_sets < 0 || (n_target_sets > 0 && !targets)) return fail(EML_ERR_INVALID_ARGUMENT, "bad targets");
    EML_CUDA(cudaSetDevice(plan->device));
    const size_t bytes = (size_t)n_target_sets * plan->P.n_obs * plan->P.n_times * 2 * sizeof(double);
    if (bytes > plan->targets_bytes) {
        EML_CUDA(cudaDeviceSynchronize());
        if (plan->d_targets) cudaFree(plan->d_targets);
        plan->d_targets = nullptr;
        plan->targets_bytes = 0;
        EML_CUDA(cudaMalloc((void**)&plan->d_targets, bytes));
        plan->targets_bytes = bytes;
    }
    if (bytes) EML_CUDA(cudaMemcpy(plan->d_targets, targets, bytes, cudaMemcpyHostToDevice));
    const double* new_targets = n_target_sets > 0 ? plan->d_targets : nullptr;
    if (new_targets != plan->P.targets || n_target_sets != plan->P.n_target_sets) plan->drop_host_graphs();   // baked into the graphs' kernel arguments
    plan->P.targets = plan->Pw.targets = n_target_sets > 0 ? plan->d_targets : nullptr;
    plan->P.n_target_sets = plan->Pw.n_target_sets = n_target_sets;
    return EML_OK;
}

int eml_plan_create(const EmlProblemDesc* d, int device, const EmlOptions* opts, EmlPlan** out) {
    if (!out) return fail(EML_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!d) return fail(EML_ERR_INVALID_ARGUMENT, "desc is NULL");
    if (d->n_tls < 1 || d->n_tls > 5) return fail(EML_ERR_UNSUPPORTED, "n_tls must be in 1..5");
    if (d->n_params < 0 || d->n_terms < 0 || (d->n_terms > 0 && !d->terms))
        return fail(EML_ERR_INVALID_ARGUMENT, "bad term list");
    if (d->n_obs < 1 || d->n_obs > eml::MAX_OBS || !d->obs) return fail(EML_ERR_INVALID_ARGUMENT, "n_obs must be in 1..8");
    if (d->n_times < 1 || !d->times) return fail(EML_ERR_INVALID_ARGUMENT, "need at least one time");
    if (!d->rho0 || !d->op_table || d->n_ops < 1) return fail(EML_ERR_INVALID_ARGUMENT, "rho0/op_table missing");
    if (d->obs_site < 0 || d->obs_site >= d->n_tls) return fail(EML_ERR_INVALID_ARGUMENT, "obs_site out of range");
    for (int k = 1; k < d->n_times; ++k)
        if (!(d->times[k] >= d->times[k - 1])) return fail(EML_ERR_INVALID_ARGUMENT, "times must be non-decreasing");
    for (int t = 0; t < d->n_terms; ++t) {
        const EmlTerm& tm = d->terms[t];
        const bool coup = tm.kind == EML_TERM_COUPLING;
        if (tm.kind < 0 || tm.kind > 2 || tm.param < 0 || tm.param >= d->n_params || tm.site_a < 0 ||
            tm.site_a >= d->n_tls || tm.op_a < 0 || tm.op_a >= d->n_ops ||
            (coup && (tm.site_b < 0 || tm.site_b >= d->n_tls || tm.site_b == tm.site_a || tm.op_b < 0 || tm.op_b >= d->n_ops)))
            return fail(EML_ERR_INVALID_ARGUMENT, "term " + std::to_string(t) + " is malformed");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(EML_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= ndev) return fail(EML_ERR_INVALID_ARGUMENT, "device out of range");
    EML_CUDA(cudaSetDevice(device));

    EmlPlan* plan = new (std::nothrow) EmlPlan();
    if (!plan) return fail(EML_ERR_CUDA, "out of host memory");
    plan->device = device;
    eml::DevProblem& P = plan->P;
    P.n_tls = d->n_tls; P.n_params = d->n_params; P.n_terms = d->n_terms; P.n_obs = d->n_obs;
    P.n_times = d->n_times; P.n_target_sets = 0; P.obs_site = d->obs_site;
    P.theta_target = (opts && opts->theta_target > 0) ? opts->theta_target : 16.0;
    P.tol = (opts && opts->tol > 0) ? opts->tol : 1e-11;
    if (P.theta_target > 16.0) P.theta_target = 16.0;  // MCAP = 80 terms converge below 1e-15 for theta <= 16
    P.h_span = d->times[d->n_times - 1] - d->times[0];
    const int dd = 1 << (2 * d->n_tls);
    int rc;
#define EML_TRY(x) if ((rc = (x)) != EML_OK) { eml_plan_destroy(plan); return rc; }
    EML_TRY(upload(plan, d->terms, (size_t)d->n_terms, &P.terms));
    EML_TRY(upload(plan, d->op_table, (size_t)d->n_ops * 8, &P.op_table));
    EML_TRY(upload(plan, d->rho0, (size_t)dd * 2, &P.rho0));
    EML_TRY(upload(plan, d->obs, (size_t)d->n_obs * 8, &P.obs));
    EML_TRY(upload(plan, d->times, (size_t)d->n_times, &P.times));
    EML_TRY(eml_plan_set_targets(plan, d->n_target_sets, d->targets));

    // Hermiticity of the generator inputs decides whether the X + X^dag kernels may be used:
    // energies use Hermitian ops, every coupling term's op pair is Hermitian (x) Hermitian, rho0 = rho0^dag.
    bool herm = true;
    for (int t = 0; t < d->n_terms && herm; ++t) {
        const EmlTerm& tm = d->terms[t];
        if (tm.kind == EML_TERM_LINDBLAD) continue;
        herm = op_is_hermitian(d->op_table + tm.op_a * 8) &&
               (tm.kind != EML_TERM_COUPLING || op_is_hermitian(d->op_table + tm.op_b * 8));
    }
    const int D = 1 << d->n_tls;
    for (int r = 0; r < D && herm; ++r)
        for (int c = 0; c < D; ++c)
            if (std::fabs(d->rho0[2 * (r * D + c)] - d->rho0[2 * (c * D + r)]) > 1e-15 ||
                std::fabs(d->rho0[2 * (r * D + c) + 1] + d->rho0[2 * (c * D + r) + 1]) > 1e-15) { herm = false; break; }
    // the warp kernel additionally needs every jump operator's A (x) conj(A) to be real, diagonal + both-flipped
    for (int t = 0; t < d->n_terms && herm; ++t)
        if (d->terms[t].kind == EML_TERM_LINDBLAD && !eml::jump_op_is_diag_flip_real(d->op_table + d->terms[t].op_a * 8))
            herm = false;
    plan->hermitian = herm;
    // real H: every energy op and every coupling op-pair product has no imaginary entries
    bool realh = herm;
    for (int t = 0; t < d->n_terms && realh; ++t) {
        const EmlTerm& tm = d->terms[t];
        if (tm.kind == EML_TERM_LINDBLAD) continue;
        const double* A = d->op_table + tm.op_a * 8;
        for (int i = 0; i < 4 && realh; ++i) {
            if (tm.kind == EML_TERM_ENERGY) { if (std::fabs(A[2 * i + 1]) > 1e-15) realh = false; continue; }
            const double* B = d->op_table + tm.op_b * 8;
            for (int j = 0; j < 4; ++j)
                if (std::fabs(A[2 * i] * B[2 * j + 1] + A[2 * i + 1] * B[2 * j]) > 1e-15) realh = false;
        }
    }
    plan->real_h = realh;
    // parity-conserving H: energy operators diagonal, coupling pairs diag (x) diag or flip (x) flip -- then H is block
    // diagonal by the parity of the basis index and the d = 16 kernel needs half the DMMAs
    bool parity = herm;
    {
        auto is_diag = [&](const double* A) { return std::fabs(A[2]) + std::fabs(A[3]) + std::fabs(A[4]) + std::fabs(A[5]) < 1e-15; };
        auto is_flip = [&](const double* A) { return std::fabs(A[0]) + std::fabs(A[1]) + std::fabs(A[6]) + std::fabs(A[7]) < 1e-15; };
        for (int t = 0; t < d->n_terms && parity; ++t) {
            const EmlTerm& tm = d->terms[t];
            if (tm.kind == EML_TERM_LINDBLAD) continue;
            const double* A = d->op_table + tm.op_a * 8;
            if (tm.kind == EML_TERM_ENERGY) { parity = is_diag(A); continue; }
            const double* B = d->op_table + tm.op_b * 8;
            parity = (is_diag(A) && is_diag(B)) || (is_flip(A) && is_flip(B));
        }
    }
    plan->parity_h = parity;
    // Steady parity-even sector (d = 16 kernel): in the parity coordinates the even tiles of rho evolve independently of
    // the odd ones; they are a steady state when they are a multiple of the identity and every jump operator of the
    // library is unital (A A^dag = A^dag A).  Then only the odd tiles are evolved (EmlOptions.reserved bit 0 disables it).
    bool sector = parity && d->n_tls == 4 && !(opts && (opts->reserved & 1));
    for (int t = 0; t < d->n_terms && sector; ++t) {
        if (d->terms[t].kind != EML_TERM_LINDBLAD) continue;
        const double* A = d->op_table + d->terms[t].op_a * 8;      // [[a, b], [c, d]] complex
        // (A A^dag - A^dag A)_00 = |b|^2 - |c|^2;  (..)_01 = a conj(c) + b conj(d) - conj(a) b - conj(c) d
        const double bb = A[2] * A[2] + A[3] * A[3], cc = A[4] * A[4] + A[5] * A[5];
        const double ore = (A[0] * A[4] + A[1] * A[5]) + (A[2] * A[6] + A[3] * A[7]) - (A[0] * A[2] + A[1] * A[3]) - (A[4] * A[6] + A[5] * A[7]);
        const double oim = (A[1] * A[4] - A[0] * A[5]) + (A[3] * A[6] - A[2] * A[7]) - (A[0] * A[3] - A[1] * A[2]) - (A[4] * A[7] - A[5] * A[6]);
        if (std::fabs(bb - cc) > 1e-15 || std::fabs(ore) > 1e-15 || std::fabs(oim) > 1e-15) sector = false;
    }
    if (sector) {
        const int Dn = 1 << d->n_tls;
        double tr = 0.0;
        for (int r = 0; r < Dn; ++r) tr += d->rho0[2 * (r * Dn + r)];
        auto par = [](int x) { return __builtin_popcount((unsigned)x) & 1; };
        for (int r = 0; r < Dn && sector; ++r)
            for (int c = 0; c < Dn; ++c) {
                if (par(r) != par(c)) continue;
                const double want = (r == c) ? tr / Dn : 0.0;
                if (std::fabs(d->rho0[2 * (r * Dn + c)] - want) > 1e-15 || std::fabs(d->rho0[2 * (r * Dn + c) + 1]) > 1e-15) { sector = false; break; }
            }
    }
    plan->sector_b = sector;
    P.lin_rows = 0; P.lin_nnz = 0; P.lin_row = nullptr; P.lin_ent = nullptr; P.lin_sq = nullptr;
    // Models with fewer than 3 sites are embedded for the warp kernel: spectator sites (identity dynamics,
    // state |0><0|) are appended as the least significant factors, which changes neither the reduced dynamics of
    // the real sites nor the observables.
    plan->Pw = P;
    if (d->n_tls < 3) {
        const int pad = 3 - d->n_tls, D0 = 1 << d->n_tls, D1 = 8;
        std::vector<double> big((size_t)D1 * D1 * 2, 0.0);
        for (int r = 0; r < D0; ++r)
            for (int c = 0; c < D0; ++c) {
                big[2 * ((size_t)(r << pad) * D1 + (c << pad))] = d->rho0[2 * (r * D0 + c)];
                big[2 * ((size_t)(r << pad) * D1 + (c << pad)) + 1] = d->rho0[2 * (r * D0 + c) + 1];
            }
        plan->Pw.n_tls = 3;
        EML_TRY(upload(plan, big.data(), big.size(), &plan->Pw.rho0));
    }
    if (eml::warp_mma_supported(plan->Pw, herm)) EML_TRY(compile_linear_program(plan, d));
    {
        void* c = nullptr;
        cudaError_t ce = cudaMalloc(&c, sizeof(unsigned int) * EML_FETCH_WS_WORDS);
        if (ce != cudaSuccess) { eml_plan_destroy(plan); return fail(EML_ERR_CUDA, "cudaMalloc(counter)"); }
        plan->owned.push_back(c);
        // zero once: the evaluators leave the workspace zero behind them (eml_warp_mma.cu: fetch_ws_release)
        if (cudaMemset(c, 0, sizeof(unsigned int) * EML_FETCH_WS_WORDS) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { eml_plan_destroy(plan); return fail(EML_ERR_CUDA, "cudaMemset(counter)"); }
        plan->d_counter = (unsigned int*)c;
        if (eml::warp_mma_supported(plan->Pw, herm) || eml::cta_mma_supported(P, herm)) {
            void* o = nullptr;
            ce = cudaMalloc(&o, sizeof(int) * EML_ORDER_WS_INTS);   // order + cost scratch, or the cost-bin lists
            if (ce != cudaSuccess) { eml_plan_destroy(plan); return fail(EML_ERR_CUDA, "cudaMalloc(order)"); }
            plan->owned.push_back(o);
            plan->d_order = (int*)o;
        }
    }

    int want = opts ? opts->kernel : EML_KERNEL_AUTO;
    if (want == EML_KERNEL_AUTO)
        want = eml::warp_mma_supported(plan->Pw, herm) ? EML_KERNEL_WARP_CHEB
               : eml::cta_mma_supported(P, herm) ? EML_KERNEL_CTA_CHEB : EML_KERNEL_GENERIC;
    if ((want == EML_KERNEL_CTA_MMA || want == EML_KERNEL_CTA_CHEB) && !eml::cta_mma_supported(P, herm)) {
        eml_plan_destroy(plan);
        return fail(EML_ERR_UNSUPPORTED, "cta kernels need n_tls == 5, Hermitian H and rho0, generalized-permutation jump operators");
    }
    if ((want == EML_KERNEL_WARP_MMA || want == EML_KERNEL_WARP_CHEB) && !eml::warp_mma_supported(plan->Pw, herm)) {
        eml_plan_destroy(plan);
        return fail(EML_ERR_UNSUPPORTED, "warp kernels need n_tls <= 4, Hermitian H and rho0, generalized-permutation jump operators");
    }
    if (want != EML_KERNEL_GENERIC && want != EML_KERNEL_WARP_MMA && want != EML_KERNEL_CTA_MMA && want != EML_KERNEL_WARP_CHEB && want != EML_KERNEL_CTA_CHEB) {
        eml_plan_destroy(plan);
        return fail(EML_ERR_INVALID_ARGUMENT, "unknown kernel id");
    }
    plan->kernel = want;
    if (eml::generic_smem_bytes(d->n_tls, d->n_params) > 220 * 1024) {
        eml_plan_destroy(plan);
        return fail(EML_ERR_UNSUPPORTED, "parameter vector too long for shared memory");
    }
    *out = plan;
    return EML_OK;
}

const char* eml_plan_kernel_name(const EmlPlan* plan) {
    if (!plan) return "";
    if (plan->kernel == EML_KERNEL_WARP_MMA) return plan->real_h ? "warp_mma_realH" : "warp_mma";
    if (plan->kernel == EML_KERNEL_WARP_CHEB) {
        if (plan->sector_b && plan->Pw.n_tls == 4) return plan->real_h ? "warp_cheb_parity_sector_realH" : "warp_cheb_parity_sector";
        if (plan->parity_h && plan->Pw.n_tls == 4) return plan->real_h ? "warp_cheb_parity_realH" : "warp_cheb_parity";
        return plan->real_h ? "warp_cheb_realH" : "warp_cheb";
    }
    if (plan->kernel == EML_KERNEL_CTA_MMA) return plan->real_h ? "cta_mma_realH" : "cta_mma";
    if (plan->kernel == EML_KERNEL_CTA_CHEB) {
        if (plan->parity_h) return plan->real_h ? "cta_cheb_parity_realH" : "cta_cheb_parity";
        return plan->real_h ? "cta_cheb_realH" : "cta_cheb";
    }
    return "generic";
}

}  // extern "C"

// bytes of prepare-kernel workspace one model needs under this plan (0: the plan's kernel has no prepare pass)
size_t eml_plan_prep_bytes_per_model(const EmlPlan* plan) {
    if (plan && plan->kernel == EML_KERNEL_CTA_CHEB && plan->parity_h && eml::cta_orbit_selected())
        return eml::cta_orbit_prep_bytes_per_model(plan->real_h);      // the d = 32 orbit kernel's records
    if (!plan || plan->kernel != EML_KERNEL_WARP_CHEB) return 0;
    return eml::warp_prep_bytes_per_model(plan->Pw, plan->real_h, true, plan->parity_h, plan->sector_b);
}

// The plan's own prepare workspace grows to the largest batch seen (capped; larger batches run in chunks): the first call
// with a larger batch than any before pays one cudaMalloc, steady-state calls allocate nothing.
static const long long kPrepMaxModels = 1 << 16;
static int ensure_prep(EmlPlan* plan, long long n_models) {
    const size_t per = eml_plan_prep_bytes_per_model(plan);
    if (per == 0) return EML_OK;
    const long long cap = (per > 2048) ? (1 << 14) : kPrepMaxModels;       // d = 32 records are 4.7 KB: 77 MB at most
    long long want = n_models < 1024 ? (per > 2048 ? (n_models < 256 ? 256 : n_models) : 1024) : n_models;
    if (want > cap) want = cap;
    if (want <= plan->prep_cap) return EML_OK;
    plan->drop_host_graphs();
    if (plan->d_prep) { EML_CUDA(cudaFree(plan->d_prep)); plan->d_prep = nullptr; plan->prep_cap = 0; }   // cudaFree waits for work in flight
    void* p = nullptr;
    EML_CUDA(cudaMalloc(&p, per * (size_t)want));
    plan->d_prep = (double*)p;
    plan->prep_cap = want;
    return EML_OK;
}

// internal entry point: eml_eval_batch plus an optional measured cost per model for the longest-first ordering
int eml_eval_batch_hinted(EmlPlan* plan, int64_t n_models, const double* params, const int32_t* target_set, double* expect,
                          double* loss, int32_t* status, void* stream, const int32_t* cost_hint, unsigned int* counter_ws,
                          int* order_ws, double* prep_ws, long long prep_cap) {
    // counter_ws / order_ws / prep_ws: caller-owned work-fetch counter, ordering buffer (EML_ORDER_WS_INTS ints) and prepare
    // workspace (prep_cap models x eml_plan_prep_bytes_per_model), so that several evaluations of one plan can be in
    // flight on different streams; NULL = the plan's own
    if (!plan) return fail(EML_ERR_INVALID_ARGUMENT, "plan is NULL");
    unsigned int* const counter = counter_ws ? counter_ws : plan->d_counter;
    int* const order = counter_ws ? order_ws : plan->d_order;
    if (n_models < 0 || n_models > 0x7fffffffLL) return fail(EML_ERR_INVALID_ARGUMENT, "n_models out of range");
    if (n_models == 0) return EML_OK;
    if (!params && plan->P.n_params > 0) return fail(EML_ERR_INVALID_ARGUMENT, "params is NULL");
    if (target_set && plan->P.n_target_sets == 0) return fail(EML_ERR_INVALID_ARGUMENT, "target_set given but plan has no targets");
    int cur = -1;
    EML_CUDA(cudaGetDevice(&cur));
    if (cur != plan->device) EML_CUDA(cudaSetDevice(plan->device));
    if (eml_plan_prep_bytes_per_model(plan) > 0 && !counter_ws) {
        const int rc = ensure_prep(plan, n_models);
        if (rc != EML_OK) { if (cur != plan->device && cur >= 0) cudaSetDevice(cur); return rc; }
        prep_ws = plan->d_prep; prep_cap = plan->prep_cap;
    }
    cudaError_t e;
    if (plan->kernel == EML_KERNEL_WARP_MMA || plan->kernel == EML_KERNEL_WARP_CHEB)
        e = eml::launch_warp_mma(plan->Pw, n_models, params, target_set, expect, loss, status, counter, order, plan->real_h,
                                 plan->kernel == EML_KERNEL_WARP_CHEB, plan->parity_h, plan->sector_b, cost_hint, prep_ws, prep_cap,
                                 plan->P.n_tls <= 2, /* pdl_ok: not for the chain engine's launches (measured slower there) */ counter_ws == nullptr,
                                 (cudaStream_t)stream);
    else if (plan->kernel == EML_KERNEL_CTA_MMA || plan->kernel == EML_KERNEL_CTA_CHEB)
        e = eml::launch_cta_mma(plan->P, n_models, params, target_set, expect, loss, status, counter, order,
                                plan->real_h, plan->kernel == EML_KERNEL_CTA_CHEB, plan->parity_h, cost_hint, prep_ws, prep_cap,
                                (cudaStream_t)stream);
    else
        e = eml::launch_generic(plan->P, n_models, params, target_set, expect, loss, status, (cudaStream_t)stream);
    g_launches.fetch_add(1);   // the evaluator kernel (the optional ordering pre-pass is not counted)
    if (cur != plan->device && cur >= 0) cudaSetDevice(cur);
    if (e != cudaSuccess) return fail(EML_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    return EML_OK;
}

extern "C" {

int eml_eval_batch(EmlPlan* plan, int64_t n_models, const double* params, const int32_t* target_set, double* expect,
                   double* loss, int32_t* status, void* stream) {
    return eml_eval_batch_hinted(plan, n_models, params, target_set, expect, loss, status, stream, nullptr, nullptr, nullptr, nullptr, 0);
}

}  // extern "C"

static int ensure_staging(EmlPlan* plan, long long n) {
    if (!plan->stream) EML_CUDA(cudaStreamCreateWithFlags(&plan->stream, cudaStreamNonBlocking));
    if (n <= plan->cap_models) return EML_OK;
    long long cap = plan->cap_models ? plan->cap_models : 1;
    while (cap < n) cap *= 2;
    EML_CUDA(cudaStreamSynchronize(plan->stream));
    plan->drop_host_graphs();
    cudaFree(plan->d_params); cudaFree(plan->d_expect); cudaFree(plan->d_loss); cudaFree(plan->d_tset); cudaFree(plan->d_status);
    cudaFreeHost(plan->h_params); cudaFreeHost(plan->h_expect); cudaFreeHost(plan->h_loss); cudaFreeHost(plan->h_tset); cudaFreeHost(plan->h_status);
    plan->d_params = plan->d_expect = plan->d_loss = nullptr; plan->d_tset = plan->d_status = nullptr;
    plan->h_params = plan->h_expect = plan->h_loss = nullptr; plan->h_tset = plan->h_status = nullptr;
    plan->cap_models = 0;
    const size_t np = (size_t)cap * (plan->P.n_params > 0 ? plan->P.n_params : 1);
    const size_t ne = (size_t)cap * plan->P.n_obs * plan->P.n_times * 2;
    EML_CUDA(cudaMalloc((void**)&plan->d_params, np * sizeof(double)));
    EML_CUDA(cudaMalloc((void**)&plan->d_expect, ne * sizeof(double)));
    EML_CUDA(cudaMalloc((void**)&plan->d_loss, cap * sizeof(double)));
    EML_CUDA(cudaMalloc((void**)&plan->d_tset, cap * sizeof(int)));
    EML_CUDA(cudaMalloc((void**)&plan->d_status, cap * sizeof(int)));
    EML_CUDA(cudaMallocHost((void**)&plan->h_params, np * sizeof(double)));
    EML_CUDA(cudaMallocHost((void**)&plan->h_expect, ne * sizeof(double)));
    EML_CUDA(cudaMallocHost((void**)&plan->h_loss, cap * sizeof(double)));
    EML_CUDA(cudaMallocHost((void**)&plan->h_tset, cap * sizeof(int)));
    EML_CUDA(cudaMallocHost((void**)&plan->h_status, cap * sizeof(int)));
    plan->cap_models = cap;
    return EML_OK;
}

extern "C" {

// EML_HOST_ZEROCOPY=0: every host call copies through device buffers (A/B measurements)
static bool zero_copy_on() {
    static const bool on = [] { const char* e = std::getenv("EML_HOST_ZEROCOPY"); return !(e && e[0] == '0'); }();
    return on;
}
// EML_HOST_ZEROCOPY=2: loss / status written to host memory by the kernels, but the parameters copied by the DMA engine
static bool zero_copy_params_on() {
    static const bool on = [] { const char* e = std::getenv("EML_HOST_ZEROCOPY"); return !(e && (e[0] == '0' || e[0] == '2')); }();
    return on;
}

int eml_eval_batch_host(EmlPlan* plan, int64_t n_models, const double* params, const int32_t* target_set,
                        double* expect, double* loss) {
    if (!plan) return fail(EML_ERR_INVALID_ARGUMENT, "plan is NULL");
    if (n_models < 0 || n_models > 0x7fffffffLL) return fail(EML_ERR_INVALID_ARGUMENT, "n_models out of range");
    if (n_models == 0) return EML_OK;
    if (!params && plan->P.n_params > 0) return fail(EML_ERR_INVALID_ARGUMENT, "params is NULL");
    int cur = -1;
    EML_CUDA(cudaGetDevice(&cur));
    EML_CUDA(cudaSetDevice(plan->device));
    int rc = ensure_staging(plan, n_models);
    if (rc != EML_OK) return rc;
    const size_t np = (size_t)n_models * plan->P.n_params;
    const size_t ne = (size_t)n_models * plan->P.n_obs * plan->P.n_times * 2;
    cudaStream_t s = plan->stream;
    // ---- small batches: replay a CUDA graph of the whole call (through the plan's page-locked staging buffers)
    static const bool graphs_on = [] { const char* e = std::getenv("EML_HOST_GRAPH"); return !(e && e[0] == '0'); }();
    if (graphs_on && n_models <= 8 && ++plan->host_calls_small > 2) {
        EmlPlan::HostGraph* hg = nullptr;
        for (EmlPlan::HostGraph& g : plan->host_graphs)
            if (g.n == n_models && g.has_tset == (target_set != nullptr) && g.has_expect == (expect != nullptr) && g.has_loss == (loss != nullptr)) hg = &g;
        if (!hg && plan->host_graphs.size() < 16) {
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                // the kernels read the parameters from, and write loss / status to, the plan's page-locked buffers directly
                // (unified addressing: three copy nodes fewer on the critical path of a one-model call); observables, which a
                // kernel writes in many small pieces, keep their device buffer and one bulk copy
                const bool zc = zero_copy_on();
                if (np && !zc) ok = cudaMemcpyAsync(plan->d_params, plan->h_params, np * sizeof(double), cudaMemcpyHostToDevice, s) == cudaSuccess && ok;
                if (target_set) ok = cudaMemcpyAsync(plan->d_tset, plan->h_tset, n_models * sizeof(int), cudaMemcpyHostToDevice, s) == cudaSuccess && ok;
                const long long before = g_launches.load();
                ok = eml_eval_batch(plan, n_models, zc ? plan->h_params : plan->d_params, target_set ? plan->d_tset : nullptr,
                                    expect ? plan->d_expect : nullptr, loss ? (zc ? plan->h_loss : plan->d_loss) : nullptr,
                                    zc ? plan->h_status : plan->d_status, s) == EML_OK && ok;
                g_launches.store(before);      // nothing ran yet; replays are counted when they are launched
                if (expect) ok = cudaMemcpyAsync(plan->h_expect, plan->d_expect, ne * sizeof(double), cudaMemcpyDeviceToHost, s) == cudaSuccess && ok;
                if (loss && !zc) ok = cudaMemcpyAsync(plan->h_loss, plan->d_loss, n_models * sizeof(double), cudaMemcpyDeviceToHost, s) == cudaSuccess && ok;
                if (!zc) ok = cudaMemcpyAsync(plan->h_status, plan->d_status, n_models * sizeof(int), cudaMemcpyDeviceToHost, s) == cudaSuccess && ok;
                ok = cudaStreamEndCapture(s, &graph) == cudaSuccess && ok && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (ok) {
                plan->host_graphs.push_back(EmlPlan::HostGraph{(long long)n_models, target_set != nullptr, expect != nullptr, loss != nullptr, exec});
                hg = &plan->host_graphs.back();
            } else {
                cudaGetLastError();
                plan->host_calls_small = -(1 << 30);      // capture is not possible here: stay on the eager path
            }
        }
        if (hg) {
            if (np) std::memcpy(plan->h_params, params, np * sizeof(double));
            if (target_set) std::memcpy(plan->h_tset, target_set, n_models * sizeof(int));
            EML_CUDA(cudaGraphLaunch(hg->exec, s));
            g_launches.fetch_add(1);
            EML_CUDA(cudaStreamSynchronize(s));
            if (expect) std::memcpy(expect, plan->h_expect, ne * sizeof(double));
            if (loss) std::memcpy(loss, plan->h_loss, n_models * sizeof(double));
            if (cur >= 0 && cur != plan->device) cudaSetDevice(cur);
            for (long long i = 0; i < n_models; ++i)
                if (plan->h_status[i] < 0)
                    return fail(EML_ERR_NOT_CONVERGED, "model " + std::to_string(i) + " did not converge (series cap or non-finite generator)");
            return EML_OK;
        }
    }
    // Page-locked caller buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) need no staging.  The parameter matrix
    // is read exactly once, by the prepare kernel (or by the d = 32 / generic evaluator), and loss / status are written once
    // per model: with unified addressing the kernels use the page-locked host memory itself, so the H2D copy (1.1 MB per
    // 4096 d = 16 models) overlaps the prepare kernel instead of preceding it and the two small D2H copies disappear.
    // Observables (many small stores per model) keep the device buffer and one bulk copy.
    auto pinned_device_ptr = [](const void* ptr) -> void* {
        if (!ptr) return nullptr;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
    };
    const bool zc = zero_copy_on();
    const double* params_dev = np ? (const double*)pinned_device_ptr(params) : nullptr;
    double* loss_dev = loss ? (double*)pinned_device_ptr(loss) : nullptr;
    const bool expect_pinned = expect && pinned_device_ptr(expect) != nullptr;
    const bool params_pinned = params_dev != nullptr, loss_pinned = loss_dev != nullptr;
    const double* params_arg = plan->d_params;
    if (params_pinned && zc && zero_copy_params_on()) {
        params_arg = params_dev;
    } else if (params_pinned) {
        EML_CUDA(cudaMemcpyAsync(plan->d_params, params, np * sizeof(double), cudaMemcpyHostToDevice, s));
    } else if (np && zc) {
        std::memcpy(plan->h_params, params, np * sizeof(double));      // pageable caller array: one host copy, then as above
        params_arg = plan->h_params;
    } else if (np) {
        // staged in chunks so that the DMA of one chunk overlaps the host copy of the next
        const size_t chunk = (size_t)1 << 15;      // 256 KB of doubles
        for (size_t off = 0; off < np; off += chunk) {
            const size_t cnt = (np - off < chunk) ? np - off : chunk;
            std::memcpy(plan->h_params + off, params + off, cnt * sizeof(double));
            EML_CUDA(cudaMemcpyAsync(plan->d_params + off, plan->h_params + off, cnt * sizeof(double), cudaMemcpyHostToDevice, s));
        }
    }
    if (target_set) {
        std::memcpy(plan->h_tset, target_set, n_models * sizeof(int));
        EML_CUDA(cudaMemcpyAsync(plan->d_tset, plan->h_tset, n_models * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    double* const loss_arg = !loss ? nullptr : (zc ? (loss_pinned ? loss_dev : plan->h_loss) : plan->d_loss);
    rc = eml_eval_batch(plan, n_models, params_arg, target_set ? plan->d_tset : nullptr,
                        expect ? plan->d_expect : nullptr, loss_arg, zc ? plan->h_status : plan->d_status, s);
    if (rc != EML_OK) return rc;
    if (expect) EML_CUDA(cudaMemcpyAsync(expect_pinned ? expect : plan->h_expect, plan->d_expect, ne * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (loss && !zc) EML_CUDA(cudaMemcpyAsync(loss_pinned ? loss : plan->h_loss, plan->d_loss, n_models * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (!zc) EML_CUDA(cudaMemcpyAsync(plan->h_status, plan->d_status, n_models * sizeof(int), cudaMemcpyDeviceToHost, s));
    EML_CUDA(cudaStreamSynchronize(s));
    if (expect && !expect_pinned) std::memcpy(expect, plan->h_expect, ne * sizeof(double));
    if (loss && !loss_pinned) std::memcpy(loss, plan->h_loss, n_models * sizeof(double));
    if (cur >= 0 && cur != plan->device) cudaSetDevice(cur);
    for (long long i = 0; i < n_models; ++i)
        if (plan->h_status[i] < 0)
            return fail(EML_ERR_NOT_CONVERGED, "model " + std::to_string(i) + " did not converge (series cap or non-finite generator)");
    return EML_OK;
}

int eml_measure_fp64_peak(int device, int use_mma, double seconds, double* flops_per_s) {
    if (!flops_per_s) return fail(EML_ERR_INVALID_ARGUMENT, "flops_per_s is NULL");
    EML_CUDA(cudaSetDevice(device));
    EML_CUDA(eml::measure_fp64(use_mma, seconds, flops_per_s));
    return EML_OK;
}

}  // extern "C"
