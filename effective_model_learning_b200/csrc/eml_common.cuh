// Shared device-side definitions for the Lindblad evaluator kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include "eml_b200.h"

// Index checks of the shared-memory tables (trace table, coefficient table, transposition tiles, prepare records): compiled
// in with -DEML_DEBUG_BOUNDS (tools/build_variant.sh dbg -DEML_DEBUG_BOUNDS), a violation traps the kernel.  The pool's
// compute-sanitizer is closed, so the GPU suite is run once per round against this build instead.
#ifdef EML_DEBUG_BOUNDS
#define EML_CHECK(cond) do { if (!(cond)) { printf("eml bounds check failed: %s (%s line %d)\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define EML_CHECK(cond) do { } while (0)
#endif

namespace eml {

constexpr int QMAX = 16;   // output times covered by one Taylor expansion (dense output)
constexpr int MCAP = 80;   // hard cap on Taylor terms per expansion (theta <= 16 with a tight norm bound needs ~66)
constexpr int MAX_OBS = 8;

// The generator of a model is LINEAR in the transformed parameters c_p (energy, coupling strength SQUARED, rate): every
// entry of G = -iH - 1/2 sum c^dag c, of the jump weights jw and of the site diagonals kd is a fixed linear form in c.
// The plan compiles the term list into those forms once (host, eml_capi.cu: compile_linear_program); the prepare kernel
// of the warp evaluators then assembles a model with independent gathers instead of walking the term list row by row.
constexpr int WARP_MAXP = 128;      // parameter-vector cap of the warp kernels
struct LinEntry { double val; int col; int pad; };      // one contribution: val * c[col]
struct LinRow { int offset, begin, end, pad; };          // output = scratch double at `offset`, contributions [begin, end)
// offsets (in doubles, from G_re) of the prepare kernel's scratch: G_re[D][D+1], G_im[D][D+1], par[WARP_MAXP], jw[32], kd[8]
__host__ __device__ constexpr int prep_off_gre(int D, int r, int c) { return r * (D + 1) + c; }
__host__ __device__ constexpr int prep_off_gim(int D, int r, int c) { return D * (D + 1) + r * (D + 1) + c; }
__host__ __device__ constexpr int prep_off_jw(int D, int j) { return 2 * D * (D + 1) + WARP_MAXP + j; }
__host__ __device__ constexpr int prep_off_kd(int D, int j) { return 2 * D * (D + 1) + WARP_MAXP + 32 + j; }

// Device-resident constants of a plan; passed to kernels by value.
struct DevProblem {
    int n_tls, n_params, n_terms, n_obs, n_times, n_target_sets, obs_site;
    int lin_rows, lin_nnz;      // lin_rows = 0: no linear program (the prepare kernel walks the term list)
    const LinRow* lin_row;      // [lin_rows], longest first
    const LinEntry* lin_ent;
    const int* lin_sq;          // [n_params] 1: the parameter enters squared (coupling strength)
    const EmlTerm* terms;   // [n_terms]
    const double* op_table; // [n_ops][8]
    const double* rho0;     // [d*d][2]
    const double* obs;      // [K][8]
    const double* times;    // [T]
    const double* targets;  // [S][K][T][2] or nullptr
    double theta_target;
    double tol;
    double h_span;              // times[T-1] - times[0] (host copy, for launch-time decisions)
};

struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cconj(cplx a) { return {a.re, -a.im}; }
__device__ __forceinline__ cplx ld_op(const double* __restrict__ op_table, int op, int row, int col) {
    const double* p = op_table + op * 8 + (row * 2 + col) * 2;
    return {p[0], p[1]};
}

// Macro-step schedule shared by all kernels so that they produce the same expansion points:
// starting at output index k (state known at times[k]) choose q >= 1 output intervals to cover
// with one expansion such that nu * span <= theta (at least one), and nsub >= 1 sub-steps when
// a single interval alone exceeds theta.
__device__ __forceinline__ void choose_step(const double* __restrict__ times, int T, int k, double nu, double theta,
                                            int& q, int& nsub, double& htot) {
    const double t0 = times[k];
    q = 1;
    while (q < QMAX && k + q < T - 1 && nu * (times[k + q + 1] - t0) <= theta) ++q;
    htot = times[k + q] - t0;
    nsub = 1;
    if (q == 1) {
        const double th = nu * htot;
        if (th > theta) nsub = (int)ceil(th / theta);
    }
}

}  // namespace eml
