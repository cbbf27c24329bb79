// Batched reversible-jump MCMC on device-resident parameter vectors (scope row N1).
// One thread per chain for the proposal and the accept/reject kernels; the evaluation in between is the
// ordinary eml_eval_batch on the proposal matrix.  Per-chain semantics: LearningChain.run
// (reference learning_chain.py:346-596), LearningModel.change_params (learning_model.py:112-194),
// ProcessHandler add/remove moves (process_handling.py:60-602).
#include <cmath>
#include <new>
#include <string>
#include <vector>

#include <cstdlib>
#include "eml_plan.hpp"

namespace eml {

constexpr int N_MOVES = 9;
// move indices (reference order of chain_step_options)
enum { MV_TWEAK = 0, MV_ADD_QL, MV_REM_QL, MV_ADD_DL, MV_REM_DL, MV_ADD_Q2D, MV_REM_Q2D, MV_ADD_D2D, MV_REM_D2D };

struct ChainHyperDev {
    int n_params, window;
    const int* col_class;
    double priority[N_MOVES];   // normalised
    double temperature, temperature_scale, complexity_factor;
    double lo[3], hi[3], jump_initial[3], jump_annealed[3];
    double jump_rescale, acc_target, acc_band;
    long long shock_anneal_at;
};

// per-chain scalar state (structure of arrays in one allocation)
struct ChainState {
    double* cur;        // [C][P]
    double* prop;       // [C][P]
    double* best;       // [C][P]
    double* cur_loss;   // [C]
    double* prop_loss;  // [C]
    double* best_loss;  // [C]
    double* jump;       // [C][3]  current jump lengths: couplings, energies, Ls
    double* log_ratio;  // [C]  log(p_back / p_there) + log prior ratio of the pending proposal
    double* temp;       // [C]  MH temperature of the pending proposal
    double* stats;      // [C][8]
    unsigned long long* rng_ctr;   // [C]
    unsigned long long* window_bits;  // [C] last accept flags, newest in bit 0
    long long* step;    // [C] iteration counter i of the reference loop
    int* kwin;          // [C] counter k of the acceptance window
    int* move;          // [C] pending move
    int* counters;      // [C][4] tot_RJ, acc_RJ, tot_tweak, acc_tweak
    int* annealed;      // [C]
    double* last_ratio; // [C] last window acceptance ratio
    int* bad_evals;     // [C] proposals rejected because their evaluation did not converge / gave a non-finite loss
};

// ---- counter-based RNG: splitmix64 of (key ^ counter * golden); mirrored in tests/chain_reference.py
__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Rng {
    unsigned long long key, ctr;
    __device__ unsigned long long next() { return splitmix64(key ^ (ctr++ * 0x9E3779B97F4A7C15ull)); }
    __device__ double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)
    __device__ double normal() {
        const double u1 = 1.0 - uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    __device__ double gamma(double shape, double scale) {   // Marsaglia-Tsang
        double boost = 1.0;
        if (shape < 1.0) { boost = pow(uniform(), 1.0 / shape); shape += 1.0; }
        const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        for (int it = 0; it < 64; ++it) {
            const double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = uniform();
            if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v * scale;
        }
        return boost * d * scale;
    }
};
__host__ __device__ inline unsigned long long chain_key(unsigned long long seed, long long chain) {
    return splitmix64(seed ^ splitmix64((unsigned long long)chain));
}

__device__ inline double Phi(double y) { return 0.5 * (1.0 + erf(y * 0.70710678118654752440)); }
__device__ inline double xi(double m, double s, double a, double b) { return Phi((b - m) / s) - Phi((a - m) / s); }

// parameter class (0 couplings, 1 energies, 2 Ls) of a column class
__device__ inline int pclass(int cc) { return (cc == EML_COL_DEFECT_ENERGY) ? 1 : (cc == EML_COL_Q2D_COUPLING || cc == EML_COL_D2D_COUPLING) ? 0 : 2; }
__device__ inline bool tweakable(int cc, double v) { return cc == EML_COL_DEFECT_ENERGY || (cc != EML_COL_FIXED && v != 0.0); }

// number of ways of each move on a parameter row, and the number of processes (model complexity)
__device__ inline void count_ways(const ChainHyperDev& H, const double* row, double* ways, int& n_proc) {
    int present[5] = {0, 0, 0, 0, 0}, total[5] = {0, 0, 0, 0, 0};
    for (int p = 0; p < H.n_params; ++p) {
        const int cc = H.col_class[p];
        if (cc >= 5) continue;
        total[cc]++;
        if (row[p] != 0.0) present[cc]++;
    }
    ways[MV_TWEAK] = 1;
    ways[MV_ADD_QL] = total[EML_COL_QUBIT_L] - present[EML_COL_QUBIT_L];
    ways[MV_REM_QL] = present[EML_COL_QUBIT_L];
    ways[MV_ADD_DL] = total[EML_COL_DEFECT_L] - present[EML_COL_DEFECT_L];
    ways[MV_REM_DL] = present[EML_COL_DEFECT_L];
    ways[MV_ADD_Q2D] = total[EML_COL_Q2D_COUPLING] - present[EML_COL_Q2D_COUPLING];
    ways[MV_REM_Q2D] = present[EML_COL_Q2D_COUPLING];
    ways[MV_ADD_D2D] = total[EML_COL_D2D_COUPLING] - present[EML_COL_D2D_COUPLING];
    ways[MV_REM_D2D] = present[EML_COL_D2D_COUPLING];
    n_proc = present[1] + present[2] + present[3] + present[4];
}
__device__ inline int move_class(int mv) {
    return (mv == MV_ADD_QL || mv == MV_REM_QL) ? EML_COL_QUBIT_L : (mv == MV_ADD_DL || mv == MV_REM_DL) ? EML_COL_DEFECT_L
         : (mv == MV_ADD_Q2D || mv == MV_REM_Q2D) ? EML_COL_Q2D_COUPLING : EML_COL_D2D_COUPLING;
}
__device__ inline bool is_add(int mv) { return mv == MV_ADD_QL || mv == MV_ADD_DL || mv == MV_ADD_Q2D || mv == MV_ADD_D2D; }
__device__ inline int complement(int mv) { return mv == MV_TWEAK ? MV_TWEAK : (is_add(mv) ? mv + 1 : mv - 1); }

__device__ inline double xi_product(const ChainHyperDev& H, const double* row, const double* jump) {
    double out = 1.0;
    for (int p = 0; p < H.n_params; ++p) {
        const int cc = H.col_class[p];
        if (!tweakable(cc, row[p])) continue;
        const int pc = pclass(cc);
        out *= xi(row[p], jump[pc], H.lo[pc], H.hi[pc]);
    }
    return out;
}

__global__ void chains_propose_kernel(const ChainHyperDev H, const ChainState S, const long long c_lo, const long long c_hi,
                                      const unsigned long long seed, const long long first_chain) {
    const long long c = c_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    const int P = H.n_params;
    double* cur = S.cur + c * P;
    double* prop = S.prop + c * P;
    double* jump = S.jump + c * 3;
    Rng rng{chain_key(seed, first_chain + c), S.rng_ctr[c]};

    // shock annealing: switch to the annealed jump lengths and restart from the best model (learning_chain.py:349-367)
    if (H.shock_anneal_at > 0 && S.step[c] == H.shock_anneal_at) {
        for (int k = 0; k < 3; ++k) jump[k] = H.jump_annealed[k];
        S.annealed[c] = 1;
        S.step[c] += 1;
        for (int p = 0; p < P; ++p) cur[p] = S.best[c * P + p];
        S.cur_loss[c] = S.best_loss[c];
        S.window_bits[c] = (S.window_bits[c] << 1) | 1ull;
    }
    // MH temperature (learning_chain.py:371, 945-958)
    const double T = (H.temperature_scale > 0.0) ? rng.gamma(H.temperature, H.temperature_scale) : H.temperature;
    S.temp[c] = T;
    // acceptance window -> jump length adaptation (learning_chain.py:374-388)
    if (S.kwin[c] >= H.window) {
        S.kwin[c] = 0;
        const unsigned long long mask = (H.window >= 64) ? ~0ull : ((1ull << H.window) - 1ull);
        const double ratio = (double)__popcll(S.window_bits[c] & mask) / (double)H.window;
        S.last_ratio[c] = ratio;
        if (ratio - H.acc_target > H.acc_band) { for (int k = 0; k < 3; ++k) jump[k] *= 1.0 / H.jump_rescale; }
        else if (ratio - H.acc_target < -H.acc_band) { for (int k = 0; k < 3; ++k) jump[k] *= H.jump_rescale; }
    }
    S.kwin[c] += 1;

    // move choice: probability ~ priority x number of ways (learning_chain.py:409-414)
    double ways[N_MOVES];
    int n_cur;
    count_ways(H, cur, ways, n_cur);
    double wsum = 0.0;
    for (int m = 0; m < N_MOVES; ++m) wsum += H.priority[m] * ways[m];
    const double u = rng.uniform();
    int mv = N_MOVES - 1;
    if (!(wsum > 0.0)) {
        // no move with a non-zero priority is available (e.g. only jump moves enabled and none applicable): re-propose the
        // current model with acceptance probability 0 -- the step is counted, the chain does not move
        for (int p = 0; p < P; ++p) prop[p] = cur[p];
        S.log_ratio[c] = 0.0;
        S.move[c] = MV_TWEAK;
        S.counters[c * 4 + 2] += 1;
        S.rng_ctr[c] = rng.ctr;
        return;
    }
    {
        double cdf = 0.0;
        for (int m = 0; m < N_MOVES; ++m) {
            cdf += H.priority[m] * ways[m] / wsum;
            if (cdf > u) { mv = m; break; }
        }
        while (mv > 0 && ways[mv] == 0.0) --mv;    // guard against round-off at the end of the cdf
    }
    for (int p = 0; p < P; ++p) prop[p] = cur[p];

    double p_there, p_back;
    if (mv == MV_TWEAK) {
        // Gaussian jitter, up to 10 redraws into the class bounds (learning_model.py:144-191)
        for (int p = 0; p < P; ++p) {
            const int cc = H.col_class[p];
            if (!tweakable(cc, cur[p])) continue;
            const int pc = pclass(cc);
            for (int a = 0; a < 10; ++a) {
                const double cand = cur[p] + jump[pc] * rng.normal();
                if (cand >= H.lo[pc] && cand <= H.hi[pc]) { prop[p] = cand; break; }
            }
        }
        p_there = xi_product(H, prop, jump);    // learning_chain.py:454-487
        p_back = xi_product(H, cur, jump);
    } else {
        const int cc = move_class(mv);
        const bool add = is_add(mv);
        const int n = (int)ways[mv];
        int pick = (int)(rng.uniform() * n);
        if (pick >= n) pick = n - 1;
        for (int p = 0; p < P; ++p) {
            if (H.col_class[p] != cc || ((cur[p] == 0.0) != add)) continue;
            if (pick-- == 0) {
                const int pc = pclass(cc);
                prop[p] = add ? H.lo[pc] + (H.hi[pc] - H.lo[pc]) * rng.uniform() : 0.0;   // process_handling.py:98-103
                break;
            }
        }
        double ways_p[N_MOVES];
        int n_dummy;
        count_ways(H, prop, ways_p, n_dummy);
        double wsum_p = 0.0;
        for (int m = 0; m < N_MOVES; ++m) wsum_p += H.priority[m] * ways_p[m];
        p_there = H.priority[mv] / wsum;                           // learning_chain.py:493-498
        p_back = H.priority[complement(mv)] / wsum_p;
    }
    int n_prop;
    {
        double wtmp[N_MOVES];
        count_ways(H, prop, wtmp, n_prop);
    }
    // prior ratio exp(-n/cf) with the reference's sub-precision guards (learning_chain.py:792-800)
    const double PP = exp(-(double)n_prop / H.complexity_factor), PC = exp(-(double)n_cur / H.complexity_factor);
    double prior_ratio;
    if (fabs(PP) < 1e-40 && fabs(PC) < 1e-40) prior_ratio = 1.0;
    else if (fabs(PC) < 1e-40) prior_ratio = INFINITY;
    else prior_ratio = PP / PC;
    S.log_ratio[c] = prior_ratio * p_back / p_there;     // stored as a plain ratio
    S.move[c] = mv;
    S.counters[c * 4 + (mv == MV_TWEAK ? 2 : 0)] += 1;
    S.rng_ctr[c] = rng.ctr;
}

__global__ void chains_accept_kernel(const ChainHyperDev H, const ChainState S, const int* __restrict__ eval_status,
                                     const long long c_lo, const long long c_hi,
                                     const unsigned long long seed, const long long first_chain, double* hist_loss,
                                     double* hist_aprob, int* hist_code, double* hist_params) {
    const long long c = c_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    const int P = H.n_params;
    Rng rng{chain_key(seed, first_chain + c), S.rng_ctr[c]};
    const double Lp = S.prop_loss[c], Lc = S.cur_loss[c];
    const double a = exp(-1.0 / S.temp[c] * (Lp - Lc)) * S.log_ratio[c];     // learning_chain.py:802-805
    // a proposal whose evaluation did not converge (status < 0: series cap, overflow guard, non-finite generator) still
    // carries a finite-looking loss: it must never be accepted or become the best model.  The uniform draw is consumed
    // either way so that the random stream does not depend on it.
    const bool usable = eval_status[c] >= 0 && isfinite(Lp);
    const bool accept = (rng.uniform() < a) && usable;
    if (!usable) S.bad_evals[c] += 1;
    const int mv = S.move[c];
    if (accept) {
        for (int p = 0; p < P; ++p) S.cur[c * P + p] = S.prop[c * P + p];
        S.cur_loss[c] = Lp;
        if (Lp < S.best_loss[c]) {
            S.best_loss[c] = Lp;
            for (int p = 0; p < P; ++p) S.best[c * P + p] = S.prop[c * P + p];
        }
        S.counters[c * 4 + (mv == MV_TWEAK ? 3 : 1)] += 1;
    }
    S.window_bits[c] = (S.window_bits[c] << 1) | (accept ? 1ull : 0ull);
    S.step[c] += 1;
    S.rng_ctr[c] = rng.ctr;
    if (hist_loss) hist_loss[c] = Lp;
    if (hist_aprob) hist_aprob[c] = a;
    if (hist_code) hist_code[c] = (mv << 1) | (accept ? 1 : 0);
    if (hist_params)
        for (int p = 0; p < P; ++p) hist_params[c * P + p] = S.cur[c * P + p];
    // statistics record for the checkpoint all_gather
    int n_proc = 0;
    for (int p = 0; p < P; ++p) {
        const int cc = H.col_class[p];
        if (cc >= 1 && cc <= 4 && S.cur[c * P + p] != 0.0) ++n_proc;
    }
    double* st = S.stats + c * EML_CHAIN_STATS;
    st[0] = S.cur_loss[c]; st[1] = S.best_loss[c];
    st[2] = S.counters[c * 4 + 3]; st[3] = S.counters[c * 4 + 2];
    st[4] = S.counters[c * 4 + 1]; st[5] = S.counters[c * 4 + 0];
    st[6] = S.last_ratio[c]; st[7] = n_proc;
}

// ---- warp-per-chain versions of the two kernels (the default; EML_CHAIN_THREAD_KERNELS=1 selects the thread-per-chain ones
//      above, which they reproduce bit for bit: same random-stream positions, same order of every floating-point operation that
//      feeds a result).  A thread-per-chain proposal is ~100 us of serial work (loops over the parameter row, Box-Muller draws,
//      two error functions per tweaked parameter) on 2048 threads = 16 CTAs; here lane l owns the parameters l, l + 32, ...
//      The Gaussian tweaks consume the chain's random stream sequentially (up to 10 redraws per parameter, the next parameter
//      continues where the previous one stopped): the lanes start from the stream positions a scan of "one draw each" gives and
//      repeat for the lanes behind a redraw until the positions are consistent (redraws are rare: usually one pass).
constexpr int CH_MAXQ = 4;      // parameters per lane: n_params <= 128

struct WaysW { double ways[N_MOVES]; int n_proc; };
// number of ways of each move from the lanes' parameters: counts per class packed in bytes (n_params <= 128 < 256)
__device__ inline WaysW count_ways_warp(const int (&cc)[CH_MAXQ], const double (&row)[CH_MAXQ], const bool (&have)[CH_MAXQ]) {
    unsigned tot = 0, pres = 0;
#pragma unroll
    for (int i = 0; i < CH_MAXQ; ++i)
        if (have[i] && cc[i] >= 1 && cc[i] <= 4) {
            tot += 1u << (8 * (cc[i] - 1));
            if (row[i] != 0.0) pres += 1u << (8 * (cc[i] - 1));
        }
    tot = __reduce_add_sync(0xffffffffu, tot);
    pres = __reduce_add_sync(0xffffffffu, pres);
    auto T = [&](int c) { return (int)((tot >> (8 * (c - 1))) & 255u); };
    auto Pn = [&](int c) { return (int)((pres >> (8 * (c - 1))) & 255u); };
    WaysW w;
    w.ways[MV_TWEAK] = 1;
    w.ways[MV_ADD_QL] = T(EML_COL_QUBIT_L) - Pn(EML_COL_QUBIT_L);       w.ways[MV_REM_QL] = Pn(EML_COL_QUBIT_L);
    w.ways[MV_ADD_DL] = T(EML_COL_DEFECT_L) - Pn(EML_COL_DEFECT_L);     w.ways[MV_REM_DL] = Pn(EML_COL_DEFECT_L);
    w.ways[MV_ADD_Q2D] = T(EML_COL_Q2D_COUPLING) - Pn(EML_COL_Q2D_COUPLING); w.ways[MV_REM_Q2D] = Pn(EML_COL_Q2D_COUPLING);
    w.ways[MV_ADD_D2D] = T(EML_COL_D2D_COUPLING) - Pn(EML_COL_D2D_COUPLING); w.ways[MV_REM_D2D] = Pn(EML_COL_D2D_COUPLING);
    w.n_proc = Pn(1) + Pn(2) + Pn(3) + Pn(4);
    return w;
}
// product over the tweakable parameters in parameter order (the order of the sequential loop in xi_product)
__device__ inline double xi_product_warp(const ChainHyperDev& H, const int (&cc)[CH_MAXQ], const double (&row)[CH_MAXQ],
                                         const bool (&have)[CH_MAXQ], const double (&jump)[3], const int nq) {
    double out = 1.0;
    for (int i = 0; i < nq; ++i) {
        const bool tw = have[i] && tweakable(cc[i], row[i]);
        const int pc = pclass(cc[i]);
        const double x = tw ? xi(row[i], jump[pc], H.lo[pc], H.hi[pc]) : 1.0;
        unsigned tm = __ballot_sync(0xffffffffu, tw);
#pragma unroll 1
        while (tm) {       // in parameter order
            const int l = __ffs(tm) - 1;
            tm &= tm - 1u;
            out *= __shfl_sync(0xffffffffu, x, l);
        }
    }
    return out;
}

__device__ __forceinline__ void propose_warp(const ChainHyperDev& H, const ChainState& S, const long long c, const int lane,
                                             const unsigned long long seed, const long long first_chain) {
    const int P = H.n_params, nq = (P + 31) >> 5;
    double* cur = S.cur + c * P;
    double* prop = S.prop + c * P;
    Rng rng{chain_key(seed, first_chain + c), S.rng_ctr[c]};      // every lane follows the same stream
    int cc[CH_MAXQ];
    bool have[CH_MAXQ];
    double curv[CH_MAXQ], propv[CH_MAXQ];
#pragma unroll
    for (int i = 0; i < CH_MAXQ; ++i) {
        const int p = lane + 32 * i;
        have[i] = i < nq && p < P;
        cc[i] = have[i] ? H.col_class[p] : EML_COL_FIXED;
        curv[i] = have[i] ? cur[p] : 0.0;
    }
    double jump[3] = {S.jump[c * 3 + 0], S.jump[c * 3 + 1], S.jump[c * 3 + 2]};
    long long step = S.step[c];
    int kwin = S.kwin[c];
    unsigned long long wbits = S.window_bits[c];
    bool anneal_now = false;
    double cur_loss_new = 0.0;
    // shock annealing: switch to the annealed jump lengths and restart from the best model (learning_chain.py:349-367)
    if (H.shock_anneal_at > 0 && step == H.shock_anneal_at) {
        for (int k = 0; k < 3; ++k) jump[k] = H.jump_annealed[k];
        anneal_now = true;
        step += 1;
#pragma unroll
        for (int i = 0; i < CH_MAXQ; ++i)
            if (have[i]) curv[i] = S.best[c * P + lane + 32 * i];
        cur_loss_new = S.best_loss[c];
        wbits = (wbits << 1) | 1ull;
    }
    // MH temperature (learning_chain.py:371, 945-958)
    const double T = (H.temperature_scale > 0.0) ? rng.gamma(H.temperature, H.temperature_scale) : H.temperature;
    // acceptance window -> jump length adaptation (learning_chain.py:374-388)
    double last_ratio = 0.0;
    bool new_ratio = false;
    if (kwin >= H.window) {
        kwin = 0;
        const unsigned long long mask = (H.window >= 64) ? ~0ull : ((1ull << H.window) - 1ull);
        const double ratio = (double)__popcll(wbits & mask) / (double)H.window;
        last_ratio = ratio; new_ratio = true;
        if (ratio - H.acc_target > H.acc_band) { for (int k = 0; k < 3; ++k) jump[k] *= 1.0 / H.jump_rescale; }
        else if (ratio - H.acc_target < -H.acc_band) { for (int k = 0; k < 3; ++k) jump[k] *= H.jump_rescale; }
    }
    kwin += 1;
    __syncwarp();      // every lane has read the chain's scalars: lane 0 may overwrite them
    if (lane == 0) {
        if (anneal_now) { S.annealed[c] = 1; S.cur_loss[c] = cur_loss_new; }
        S.step[c] = step; S.window_bits[c] = wbits; S.kwin[c] = kwin; S.temp[c] = T;
        if (new_ratio) S.last_ratio[c] = last_ratio;
        for (int k = 0; k < 3; ++k) S.jump[c * 3 + k] = jump[k];
    }
    if (anneal_now) {
#pragma unroll
        for (int i = 0; i < CH_MAXQ; ++i)
            if (have[i]) cur[lane + 32 * i] = curv[i];
    }

    // move choice: probability ~ priority x number of ways (learning_chain.py:409-414)
    const WaysW wc = count_ways_warp(cc, curv, have);
    double wsum = 0.0;
    for (int m = 0; m < N_MOVES; ++m) wsum += H.priority[m] * wc.ways[m];
    const double u = rng.uniform();
#pragma unroll
    for (int i = 0; i < CH_MAXQ; ++i) propv[i] = curv[i];
    if (!(wsum > 0.0)) {
        // no move with a non-zero priority is available: re-propose the current model with acceptance probability 0
#pragma unroll
        for (int i = 0; i < CH_MAXQ; ++i)
            if (have[i]) prop[lane + 32 * i] = curv[i];
        if (lane == 0) {
            S.log_ratio[c] = 0.0; S.move[c] = MV_TWEAK; S.counters[c * 4 + 2] += 1; S.rng_ctr[c] = rng.ctr;
        }
        return;
    }
    int mv = N_MOVES - 1;
    {
        double cdf = 0.0;
        for (int m = 0; m < N_MOVES; ++m) {
            cdf += H.priority[m] * wc.ways[m] / wsum;
            if (cdf > u) { mv = m; break; }
        }
        while (mv > 0 && wc.ways[mv] == 0.0) --mv;    // guard against round-off at the end of the cdf
    }
    double p_there, p_back;
    if (mv == MV_TWEAK) {
        // Gaussian jitter, up to 10 redraws into the class bounds (learning_model.py:144-191)
        for (int i = 0; i < nq; ++i) {
            const bool tw = have[i] && tweakable(cc[i], curv[i]);
            const int pc = pclass(cc[i]);
            int att = tw ? 1 : 0, my_off = -1;
            double val = curv[i];
            unsigned long long total = 0;
            for (;;) {
                int incl = att;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += n;
                }
                const int excl = incl - att;
                total = (unsigned long long)__shfl_sync(0xffffffffu, incl, 31);
                int natt = att;
                if (tw && excl != my_off) {
                    my_off = excl;
                    Rng r{rng.key, rng.ctr + 2ull * (unsigned long long)excl};
                    val = curv[i]; natt = 10;
                    for (int a = 0; a < 10; ++a) {
                        const double cand = curv[i] + jump[pc] * r.normal();
                        if (cand >= H.lo[pc] && cand <= H.hi[pc]) { val = cand; natt = a + 1; break; }
                    }
                }
                const bool changed = natt != att;
                att = natt;
                if (!__any_sync(0xffffffffu, changed)) break;
            }
            if (tw) propv[i] = val;
            rng.ctr += 2ull * total;
        }
        p_there = xi_product_warp(H, cc, propv, have, jump, nq);    // learning_chain.py:454-487
        p_back = xi_product_warp(H, cc, curv, have, jump, nq);
    } else {
        const int mcc = move_class(mv);
        const bool add = is_add(mv);
        const int n = (int)wc.ways[mv];
        int pick = (int)(rng.uniform() * n);
        if (pick >= n) pick = n - 1;
        bool done = false;
        for (int i = 0; i < nq && !done; ++i) {
            const bool cand = have[i] && cc[i] == mcc && ((curv[i] == 0.0) == add);
            const unsigned cm = __ballot_sync(0xffffffffu, cand);
            const int cnt = __popc(cm);
            if (pick < cnt) {
                int sel = 0;       // lane of the (pick + 1)-th candidate
                {
                    unsigned rest = cm;
                    for (int j = 0; j < pick; ++j) rest &= rest - 1u;
                    sel = __ffs(rest) - 1;
                }
                const int pc = pclass(mcc);
                const double v = add ? H.lo[pc] + (H.hi[pc] - H.lo[pc]) * rng.uniform() : 0.0;   // process_handling.py:98-103
                if (lane == sel) propv[i] = v;
                done = true;
            } else {
                pick -= cnt;
            }
        }
        const WaysW wp = count_ways_warp(cc, propv, have);
        double wsum_p = 0.0;
        for (int m = 0; m < N_MOVES; ++m) wsum_p += H.priority[m] * wp.ways[m];
        p_there = H.priority[mv] / wsum;                           // learning_chain.py:493-498
        p_back = H.priority[complement(mv)] / wsum_p;
    }
    const int n_prop = count_ways_warp(cc, propv, have).n_proc, n_cur = wc.n_proc;
#pragma unroll
    for (int i = 0; i < CH_MAXQ; ++i)
        if (have[i]) prop[lane + 32 * i] = propv[i];
    if (lane == 0) {
        // prior ratio exp(-n/cf) with the reference's sub-precision guards (learning_chain.py:792-800)
        const double PP = exp(-(double)n_prop / H.complexity_factor), PC = exp(-(double)n_cur / H.complexity_factor);
        double prior_ratio;
        if (fabs(PP) < 1e-40 && fabs(PC) < 1e-40) prior_ratio = 1.0;
        else if (fabs(PC) < 1e-40) prior_ratio = INFINITY;
        else prior_ratio = PP / PC;
        S.log_ratio[c] = prior_ratio * p_back / p_there;     // stored as a plain ratio
        S.move[c] = mv;
        S.counters[c * 4 + (mv == MV_TWEAK ? 2 : 0)] += 1;
        S.rng_ctr[c] = rng.ctr;
    }
}

__device__ __forceinline__ void accept_warp(const ChainHyperDev& H, const ChainState& S, const int* __restrict__ eval_status,
                                            const long long c, const int lane, const unsigned long long seed,
                                            const long long first_chain, double* hist_loss, double* hist_aprob,
                                            int* hist_code, double* hist_params) {
    const int P = H.n_params;
    Rng rng{chain_key(seed, first_chain + c), S.rng_ctr[c]};
    const double Lp = S.prop_loss[c], Lc = S.cur_loss[c];
    const double best_before = S.best_loss[c];
    const double a = exp(-1.0 / S.temp[c] * (Lp - Lc)) * S.log_ratio[c];     // learning_chain.py:802-805
    // a proposal whose evaluation did not converge must never be accepted or become the best model; the uniform draw is
    // consumed either way so that the random stream does not depend on it
    const bool usable = eval_status[c] >= 0 && isfinite(Lp);
    const bool accept = (rng.uniform() < a) && usable;
    const int mv = S.move[c];
    const bool new_best = accept && Lp < best_before;
    int n_proc = 0;
    for (int p = lane; p < P; p += 32) {
        const double pv = S.prop[c * P + p], cv = accept ? pv : S.cur[c * P + p];
        if (accept) {
            S.cur[c * P + p] = pv;
            if (new_best) S.best[c * P + p] = pv;
        }
        if (hist_params) hist_params[c * P + p] = cv;
        const int cls = H.col_class[p];
        if (cls >= 1 && cls <= 4 && cv != 0.0) ++n_proc;
    }
    n_proc = __reduce_add_sync(0xffffffffu, n_proc);
    __syncwarp();      // every lane has read the chain's scalars
    if (lane == 0) {
        if (!usable) S.bad_evals[c] += 1;
        if (accept) {
            S.cur_loss[c] = Lp;
            if (new_best) S.best_loss[c] = Lp;
            S.counters[c * 4 + (mv == MV_TWEAK ? 3 : 1)] += 1;
        }
        S.window_bits[c] = (S.window_bits[c] << 1) | (accept ? 1ull : 0ull);
        S.step[c] += 1;
        S.rng_ctr[c] = rng.ctr;
        if (hist_loss) hist_loss[c] = Lp;
        if (hist_aprob) hist_aprob[c] = a;
        if (hist_code) hist_code[c] = (mv << 1) | (accept ? 1 : 0);
        // statistics record for the checkpoint all_gather
        double* st = S.stats + c * EML_CHAIN_STATS;
        st[0] = accept ? Lp : Lc; st[1] = new_best ? Lp : best_before;
        st[2] = S.counters[c * 4 + 3]; st[3] = S.counters[c * 4 + 2];
        st[4] = S.counters[c * 4 + 1]; st[5] = S.counters[c * 4 + 0];
        st[6] = S.last_ratio[c]; st[7] = n_proc;
    }
}

// eight CTAs (32 chains) per SM: 4096 chains are a single wave on 148 SMs
__global__ void __launch_bounds__(128, 8) chains_propose_warp_kernel(const ChainHyperDev H, const ChainState S, const long long c_lo,
                                                                     const long long c_hi, const unsigned long long seed,
                                                                     const long long first_chain) {
    const long long c = c_lo + (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= c_hi) return;      // warp-uniform
    propose_warp(H, S, c, threadIdx.x & 31, seed, first_chain);
}
__global__ void __launch_bounds__(128, 8) chains_accept_warp_kernel(const ChainHyperDev H, const ChainState S, const int* __restrict__ eval_status,
                                                                    const long long c_lo, const long long c_hi, const unsigned long long seed,
                                                                    const long long first_chain, double* hist_loss, double* hist_aprob,
                                                                    int* hist_code, double* hist_params) {
    const long long c = c_lo + (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= c_hi) return;      // warp-uniform
    accept_warp(H, S, eval_status, c, threadIdx.x & 31, seed, first_chain, hist_loss, hist_aprob, hist_code, hist_params);
}
// accept / reject of step i and the proposal of step i + 1 in one launch (the same random-stream order as the two kernels)
__global__ void __launch_bounds__(128, 8) chains_accept_propose_warp_kernel(const ChainHyperDev H, const ChainState S,
                                                                            const int* __restrict__ eval_status, const long long c_lo,
                                                                            const long long c_hi, const unsigned long long seed,
                                                                            const long long first_chain, double* hist_loss,
                                                                            double* hist_aprob, int* hist_code, double* hist_params) {
    const long long c = c_lo + (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= c_hi) return;      // warp-uniform
    const int lane = threadIdx.x & 31;
    accept_warp(H, S, eval_status, c, lane, seed, first_chain, hist_loss, hist_aprob, hist_code, hist_params);
    __syncwarp();       // the chain's scalars and rows written by the accept part are visible to every lane
    propose_warp(H, S, c, lane, seed, first_chain);
}

__global__ void chains_init_kernel(const ChainHyperDev H, const ChainState S, const long long n_chains) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    const int P = H.n_params;
    for (int p = 0; p < P; ++p) { S.best[c * P + p] = S.cur[c * P + p]; S.prop[c * P + p] = S.cur[c * P + p]; }
    S.cur_loss[c] = S.prop_loss[c];
    S.best_loss[c] = S.prop_loss[c];
    for (int k = 0; k < 3; ++k) S.jump[c * 3 + k] = H.jump_initial[k];
    S.rng_ctr[c] = 0; S.window_bits[c] = 0; S.step[c] = 0; S.kwin[c] = 0; S.move[c] = 0; S.annealed[c] = 0;
    S.last_ratio[c] = 0.0; S.bad_evals[c] = 0;
    for (int k = 0; k < 4; ++k) S.counters[c * 4 + k] = 0;
    int n_proc = 0;
    for (int p = 0; p < P; ++p) {
        const int cc = H.col_class[p];
        if (cc >= 1 && cc <= 4 && S.cur[c * P + p] != 0.0) ++n_proc;
    }
    double* st = S.stats + c * EML_CHAIN_STATS;
    st[0] = st[1] = S.prop_loss[c];
    st[2] = st[3] = st[4] = st[5] = st[6] = 0.0;
    st[7] = n_proc;
}

}  // namespace eml

struct EmlChains {
    EmlPlan* plan = nullptr;
    long long n = 0;
    unsigned long long seed = 0;
    long long first_chain = 0;
    eml::ChainHyperDev H{};
    eml::ChainState S{};
    std::vector<void*> owned;
    int* d_target_set = nullptr;
    int* d_status = nullptr;
    std::vector<int> h_col_class;
    // Large batches advance as two independent groups of chains on two streams: while the last (shortest) models
    // of one group's evaluation finish, the other group's next evaluation already fills the idle SMs.
    static constexpr int MAX_GROUPS = 4;
    int n_groups = 1;
    long long group_lo[MAX_GROUPS] = {}, group_hi[MAX_GROUPS] = {};
    cudaStream_t gstream[MAX_GROUPS] = {};
    cudaEvent_t fork = nullptr, join[MAX_GROUPS] = {};
    unsigned int* g_counter[MAX_GROUPS] = {};
    int* g_order[MAX_GROUPS] = {};
    double* g_prep[MAX_GROUPS] = {};     // prepare-kernel workspace of each group (its own: the groups run concurrently)
};

template <typename T>
static int dalloc(EmlChains* ch, T** out, size_t count) {
    void* p = nullptr;
    EML_CUDA(cudaMalloc(&p, (count ? count : 1) * sizeof(T)));
    ch->owned.push_back(p);
    *out = (T*)p;
    return EML_OK;
}

extern "C" {

int eml_chains_destroy(EmlChains* ch) {
    if (!ch) return EML_OK;
    cudaSetDevice(ch->plan->device);
    cudaDeviceSynchronize();
    for (void* p : ch->owned) cudaFree(p);
    for (int g = 0; g < EmlChains::MAX_GROUPS; ++g) {
        if (ch->gstream[g]) cudaStreamDestroy(ch->gstream[g]);
        if (ch->join[g]) cudaEventDestroy(ch->join[g]);
    }
    if (ch->fork) cudaEventDestroy(ch->fork);
    delete ch;
    return EML_OK;
}

int eml_chains_create(EmlPlan* plan, const EmlChainHyper* h, int64_t n_chains, const double* initial_params,
                      const int32_t* target_set, uint64_t seed, int64_t first_chain, EmlChains** out) {
    if (!out) return fail(EML_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!plan || !h || !initial_params) return fail(EML_ERR_INVALID_ARGUMENT, "plan/hyper/initial_params is NULL");
    if (n_chains < 1 || n_chains > 0x7fffffffLL) return fail(EML_ERR_INVALID_ARGUMENT, "n_chains out of range");
    if (h->n_params != plan->P.n_params) return fail(EML_ERR_INVALID_ARGUMENT, "hyper.n_params != plan n_params");
    if (h->acceptance_window < 1 || h->acceptance_window > 64) return fail(EML_ERR_INVALID_ARGUMENT, "acceptance_window must be in 1..64");
    if (plan->P.n_target_sets < 1) return fail(EML_ERR_INVALID_ARGUMENT, "the plan has no targets");
    if (!(h->complexity_factor > 0) || !(h->jump_rescale > 0)) return fail(EML_ERR_INVALID_ARGUMENT, "bad hyperparameters");
    for (int k = 0; k < 3; ++k)
        if (!(h->bounds_hi[k] > h->bounds_lo[k]) || !(h->jump_initial[k] > 0)) return fail(EML_ERR_INVALID_ARGUMENT, "bounds/jump lengths required");
    double psum = 0;
    for (int m = 0; m < eml::N_MOVES; ++m) { if (h->priority[m] < 0) return fail(EML_ERR_INVALID_ARGUMENT, "negative priority"); psum += h->priority[m]; }
    if (!(psum > 0)) return fail(EML_ERR_INVALID_ARGUMENT, "all priorities are zero");
    EML_CUDA(cudaSetDevice(plan->device));
    EmlChains* ch = new (std::nothrow) EmlChains();
    if (!ch) return fail(EML_ERR_CUDA, "out of host memory");
    ch->plan = plan; ch->n = n_chains; ch->seed = seed; ch->first_chain = first_chain;
    const int P = h->n_params;
    eml::ChainHyperDev& H = ch->H;
    H.n_params = P; H.window = h->acceptance_window;
    for (int m = 0; m < eml::N_MOVES; ++m) H.priority[m] = h->priority[m] / psum;
    H.temperature = h->temperature; H.temperature_scale = h->temperature_scale; H.complexity_factor = h->complexity_factor;
    for (int k = 0; k < 3; ++k) { H.lo[k] = h->bounds_lo[k]; H.hi[k] = h->bounds_hi[k]; H.jump_initial[k] = h->jump_initial[k]; H.jump_annealed[k] = h->jump_annealed[k]; }
    H.jump_rescale = h->jump_rescale; H.acc_target = h->acceptance_target; H.acc_band = h->acceptance_band;
    H.shock_anneal_at = h->shock_anneal_at;
    int rc;
#define CH_TRY(x) if ((rc = (x)) != EML_OK) { eml_chains_destroy(ch); return rc; }
    int* d_cls = nullptr;
    CH_TRY(dalloc(ch, &d_cls, (size_t)P));
    if (cudaMemcpy(d_cls, h->col_class, sizeof(int) * P, cudaMemcpyHostToDevice) != cudaSuccess) { eml_chains_destroy(ch); return fail(EML_ERR_CUDA, "memcpy col_class"); }
    H.col_class = d_cls;
    ch->h_col_class.assign(h->col_class, h->col_class + P);
    eml::ChainState& S = ch->S;
    const size_t C = (size_t)n_chains;
    CH_TRY(dalloc(ch, &S.cur, C * P)); CH_TRY(dalloc(ch, &S.prop, C * P)); CH_TRY(dalloc(ch, &S.best, C * P));
    CH_TRY(dalloc(ch, &S.cur_loss, C)); CH_TRY(dalloc(ch, &S.prop_loss, C)); CH_TRY(dalloc(ch, &S.best_loss, C));
    CH_TRY(dalloc(ch, &S.jump, C * 3)); CH_TRY(dalloc(ch, &S.log_ratio, C)); CH_TRY(dalloc(ch, &S.temp, C));
    CH_TRY(dalloc(ch, &S.stats, C * EML_CHAIN_STATS)); CH_TRY(dalloc(ch, &S.rng_ctr, C)); CH_TRY(dalloc(ch, &S.window_bits, C));
    CH_TRY(dalloc(ch, &S.step, C)); CH_TRY(dalloc(ch, &S.kwin, C)); CH_TRY(dalloc(ch, &S.move, C));
    CH_TRY(dalloc(ch, &S.counters, C * 4)); CH_TRY(dalloc(ch, &S.annealed, C)); CH_TRY(dalloc(ch, &S.last_ratio, C));
    CH_TRY(dalloc(ch, &S.bad_evals, C));
    CH_TRY(dalloc(ch, &ch->d_status, C));
    if (target_set) {
        for (int64_t i = 0; i < n_chains; ++i)
            if (target_set[i] < 0 || target_set[i] >= plan->P.n_target_sets) { eml_chains_destroy(ch); return fail(EML_ERR_INVALID_ARGUMENT, "target_set entry out of range"); }
        CH_TRY(dalloc(ch, &ch->d_target_set, C));
        if (cudaMemcpy(ch->d_target_set, target_set, sizeof(int) * C, cudaMemcpyHostToDevice) != cudaSuccess) { eml_chains_destroy(ch); return fail(EML_ERR_CUDA, "memcpy target_set"); }
    }
    if (cudaMemcpy(S.cur, initial_params, sizeof(double) * C * P, cudaMemcpyHostToDevice) != cudaSuccess) { eml_chains_destroy(ch); return fail(EML_ERR_CUDA, "memcpy initial params"); }
    // loss of the initial models (LearningChain.__init__, learning_chain.py:298)
    CH_TRY(eml_eval_batch(plan, n_chains, S.cur, ch->d_target_set, nullptr, S.prop_loss, ch->d_status, nullptr));
    {   // a chain cannot start from a model whose evaluation failed: its acceptance probabilities would all be NaN
        std::vector<int> st(C);
        std::vector<double> l0(C);
        if (cudaMemcpy(st.data(), ch->d_status, sizeof(int) * C, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(l0.data(), S.prop_loss, sizeof(double) * C, cudaMemcpyDeviceToHost) != cudaSuccess) {
            eml_chains_destroy(ch);
            return fail(EML_ERR_CUDA, "memcpy initial status");
        }
        for (size_t i = 0; i < C; ++i)
            if (st[i] < 0 || !std::isfinite(l0[i])) {
                eml_chains_destroy(ch);
                return fail(EML_ERR_NOT_CONVERGED, "the initial model of chain " + std::to_string(i) + " could not be evaluated (series cap, overflow guard or non-finite parameters)");
            }
    }
    // groups: two halves from 2048 chains on, so that one half's proposal / prepare / accept kernels and the tail of its
    // evaluation overlap the other half's evaluation -- except when a half would fit the evaluator's warp slots (16 per SM)
    // while the whole batch does not: such a half is bound by the latency of its longest model (2048 of 4096 d = 16 chains:
    // 245 us against 333 us for all 4096 in one launch with 1.7 models per slot, measured 10.2 M against 9.5 M steps/s)
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device);
        const long long slots = 16LL * sms;
        ch->n_groups = (n_chains >= 2048 && !(n_chains / 2 <= slots && n_chains > slots)) ? 2 : 1;
    }
    if (const char* env = std::getenv("EML_CHAIN_GROUPS")) {       // A/B switch for measurements
        const int want = std::atoi(env);
        if (want >= 1 && want <= EmlChains::MAX_GROUPS && n_chains >= want) ch->n_groups = want;
    }
    for (int g = 0; g < ch->n_groups; ++g) {
        ch->group_lo[g] = g * n_chains / ch->n_groups;
        ch->group_hi[g] = (g + 1) * n_chains / ch->n_groups;
        CH_TRY(dalloc(ch, &ch->g_counter[g], EML_FETCH_WS_WORDS));
        EML_CUDA(cudaMemset(ch->g_counter[g], 0, sizeof(unsigned int) * EML_FETCH_WS_WORDS));
        EML_CUDA(cudaDeviceSynchronize());      // (the group streams do not synchronise with the stream of the memset)
        CH_TRY(dalloc(ch, &ch->g_order[g], EML_ORDER_WS_INTS));
        CH_TRY(dalloc(ch, &ch->g_prep[g], eml_plan_prep_bytes_per_model(plan) / sizeof(double) * (size_t)(ch->group_hi[g] - ch->group_lo[g])));
        if (cudaStreamCreateWithFlags(&ch->gstream[g], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ch->join[g], cudaEventDisableTiming) != cudaSuccess) {
            eml_chains_destroy(ch);
            return fail(EML_ERR_CUDA, "chains: stream/event creation failed");
        }
    }
    if (cudaEventCreateWithFlags(&ch->fork, cudaEventDisableTiming) != cudaSuccess) { eml_chains_destroy(ch); return fail(EML_ERR_CUDA, "chains: event creation failed"); }
    const unsigned blocks = (unsigned)((n_chains + 127) / 128);
    eml::chains_init_kernel<<<blocks, 128>>>(H, S, n_chains);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { eml_chains_destroy(ch); return fail(EML_ERR_CUDA, std::string("chains init: ") + cudaGetErrorString(e)); }
    *out = ch;
    return EML_OK;
}

int eml_chains_run(EmlChains* ch, int64_t n_steps, double* hist_loss, double* hist_aprob, int32_t* hist_code,
                   double* hist_params, void* stream) {
    if (!ch) return fail(EML_ERR_INVALID_ARGUMENT, "chains is NULL");
    if (n_steps < 0) return fail(EML_ERR_INVALID_ARGUMENT, "n_steps < 0");
    EML_CUDA(cudaSetDevice(ch->plan->device));
    cudaStream_t s = (cudaStream_t)stream;
    // fork: the group streams start after everything already queued on the caller's stream
    EML_CUDA(cudaEventRecord(ch->fork, s));
    for (int g = 0; g < ch->n_groups; ++g) EML_CUDA(cudaStreamWaitEvent(ch->gstream[g], ch->fork, 0));
    for (int64_t i = 0; i < n_steps; ++i) {
        for (int g = 0; g < ch->n_groups; ++g) {
            const long long lo = ch->group_lo[g], hi = ch->group_hi[g], cnt = hi - lo;
            if (cnt <= 0) continue;
            cudaStream_t gs = ch->gstream[g];
            const unsigned blocks = (unsigned)((cnt + 127) / 128);
            const int P = ch->H.n_params;
            static const bool thread_kernels = [] { const char* e = std::getenv("EML_CHAIN_THREAD_KERNELS"); return e && e[0] == '1'; }();
            const bool warp_kernels = !thread_kernels && P <= 32 * eml::CH_MAXQ;
            const unsigned wblocks = (unsigned)((cnt + 3) / 4);       // a warp per chain, four chains per CTA
            // warp kernels: the accept of step i and the proposal of step i + 1 are one launch (EML_CHAIN_FUSED=0: two)
            static const bool fused_on = [] { const char* e = std::getenv("EML_CHAIN_FUSED"); return !(e && e[0] == '0'); }();
            const bool fused = warp_kernels && fused_on;
            if (warp_kernels) {
                if (!fused || i == 0) eml::chains_propose_warp_kernel<<<wblocks, 128, 0, gs>>>(ch->H, ch->S, lo, hi, ch->seed, ch->first_chain);
            } else {
                eml::chains_propose_kernel<<<blocks, 128, 0, gs>>>(ch->H, ch->S, lo, hi, ch->seed, ch->first_chain);
            }
            // the generator applications each chain's previous proposal needed order this step's models (longest first)
            int rc = eml_eval_batch_hinted(ch->plan, cnt, ch->S.prop + lo * P, ch->d_target_set ? ch->d_target_set + lo : nullptr,
                                           nullptr, ch->S.prop_loss + lo, ch->d_status + lo, gs, ch->d_status + lo,
                                           ch->g_counter[g], ch->g_order[g], ch->g_prep[g], cnt);
            if (rc != EML_OK) return rc;
            double* const hl = hist_loss ? hist_loss + i * ch->n : nullptr;
            double* const ha = hist_aprob ? hist_aprob + i * ch->n : nullptr;
            int* const hc = hist_code ? hist_code + i * ch->n : nullptr;
            double* const hp = hist_params ? hist_params + i * ch->n * P : nullptr;
            if (fused && i + 1 < n_steps)
                eml::chains_accept_propose_warp_kernel<<<wblocks, 128, 0, gs>>>(ch->H, ch->S, ch->d_status, lo, hi, ch->seed, ch->first_chain, hl, ha, hc, hp);
            else if (warp_kernels)
                eml::chains_accept_warp_kernel<<<wblocks, 128, 0, gs>>>(ch->H, ch->S, ch->d_status, lo, hi, ch->seed, ch->first_chain, hl, ha, hc, hp);
            else
                eml::chains_accept_kernel<<<blocks, 128, 0, gs>>>(ch->H, ch->S, ch->d_status, lo, hi, ch->seed, ch->first_chain, hl, ha, hc, hp);
        }
    }
    // join: the caller's stream continues after both groups are done
    for (int g = 0; g < ch->n_groups; ++g) {
        EML_CUDA(cudaEventRecord(ch->join[g], ch->gstream[g]));
        EML_CUDA(cudaStreamWaitEvent(s, ch->join[g], 0));
    }
    EML_CUDA(cudaGetLastError());
    return EML_OK;
}

int eml_chains_read(EmlChains* ch, double* current_params, double* current_loss, double* best_params, double* best_loss, double* stats) {
    if (!ch) return fail(EML_ERR_INVALID_ARGUMENT, "chains is NULL");
    EML_CUDA(cudaSetDevice(ch->plan->device));
    EML_CUDA(cudaDeviceSynchronize());
    const size_t C = (size_t)ch->n, P = (size_t)ch->H.n_params;
    if (current_params) EML_CUDA(cudaMemcpy(current_params, ch->S.cur, sizeof(double) * C * P, cudaMemcpyDeviceToHost));
    if (best_params) EML_CUDA(cudaMemcpy(best_params, ch->S.best, sizeof(double) * C * P, cudaMemcpyDeviceToHost));
    if (current_loss) EML_CUDA(cudaMemcpy(current_loss, ch->S.cur_loss, sizeof(double) * C, cudaMemcpyDeviceToHost));
    if (best_loss) EML_CUDA(cudaMemcpy(best_loss, ch->S.best_loss, sizeof(double) * C, cudaMemcpyDeviceToHost));
    if (stats) EML_CUDA(cudaMemcpy(stats, ch->S.stats, sizeof(double) * C * EML_CHAIN_STATS, cudaMemcpyDeviceToHost));
    return EML_OK;
}

const double* eml_chains_stats_device(EmlChains* ch) { return ch ? ch->S.stats : nullptr; }

int eml_chains_bad_evaluations(EmlChains* ch, int32_t* per_chain, int64_t* total) {
    if (!ch) return fail(EML_ERR_INVALID_ARGUMENT, "chains is NULL");
    EML_CUDA(cudaSetDevice(ch->plan->device));
    EML_CUDA(cudaDeviceSynchronize());
    std::vector<int> host((size_t)ch->n);
    EML_CUDA(cudaMemcpy(host.data(), ch->S.bad_evals, sizeof(int) * (size_t)ch->n, cudaMemcpyDeviceToHost));
    long long sum = 0;
    for (size_t i = 0; i < host.size(); ++i) { sum += host[i]; if (per_chain) per_chain[i] = host[i]; }
    if (total) *total = sum;
    return EML_OK;
}

// ---- checkpoint / resume: every per-chain array, in a fixed order
struct Segment { void* ptr; size_t bytes; };
static std::vector<Segment> state_segments(const EmlChains* ch) {
    const size_t C = (size_t)ch->n, P = (size_t)ch->H.n_params;
    const eml::ChainState& S = ch->S;
    return {{S.cur, C * P * 8}, {S.best, C * P * 8}, {S.cur_loss, C * 8}, {S.best_loss, C * 8}, {S.jump, C * 3 * 8},
            {S.stats, C * EML_CHAIN_STATS * 8}, {S.rng_ctr, C * 8}, {S.window_bits, C * 8}, {S.step, C * 8},
            {S.last_ratio, C * 8}, {S.kwin, C * 4}, {S.counters, C * 4 * 4}, {S.annealed, C * 4}, {ch->d_status, C * 4},
            {S.bad_evals, C * 4}};
}
static const unsigned long long kBlobMagic = 0x454d4c4348413032ull;   // "EMLCHA02"
constexpr int kBlobHead = 6;      // magic, n_chains, n_params, seed, first_chain, layout hash
// FNV-1a over everything a checkpoint silently depends on: column classes, bounds, priorities, jump lengths, window ...
static unsigned long long layout_hash(const EmlChains* ch, const std::vector<int>& col_class) {
    unsigned long long h = 0xcbf29ce484222325ull;
    auto mix = [&h](const void* p, size_t n) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    };
    const eml::ChainHyperDev& H = ch->H;
    mix(col_class.data(), col_class.size() * sizeof(int));
    mix(&H.window, sizeof(H.window)); mix(H.priority, sizeof(H.priority));
    mix(&H.temperature, sizeof(double)); mix(&H.temperature_scale, sizeof(double)); mix(&H.complexity_factor, sizeof(double));
    mix(H.lo, sizeof(H.lo)); mix(H.hi, sizeof(H.hi)); mix(H.jump_initial, sizeof(H.jump_initial)); mix(H.jump_annealed, sizeof(H.jump_annealed));
    mix(&H.jump_rescale, sizeof(double)); mix(&H.acc_target, sizeof(double)); mix(&H.acc_band, sizeof(double));
    mix(&H.shock_anneal_at, sizeof(H.shock_anneal_at));
    return h;
}

size_t eml_chains_state_bytes(const EmlChains* ch) {
    if (!ch) return 0;
    size_t total = kBlobHead * sizeof(unsigned long long);
    for (const Segment& s : state_segments(ch)) total += s.bytes;
    return total;
}

int eml_chains_export(EmlChains* ch, void* blob) {
    if (!ch || !blob) return fail(EML_ERR_INVALID_ARGUMENT, "chains/blob is NULL");
    EML_CUDA(cudaSetDevice(ch->plan->device));
    EML_CUDA(cudaDeviceSynchronize());
    unsigned long long* head = (unsigned long long*)blob;
    head[0] = kBlobMagic; head[1] = (unsigned long long)ch->n; head[2] = (unsigned long long)ch->H.n_params; head[3] = ch->seed;
    head[4] = (unsigned long long)ch->first_chain; head[5] = layout_hash(ch, ch->h_col_class);
    char* out = (char*)(head + kBlobHead);
    for (const Segment& s : state_segments(ch)) {
        EML_CUDA(cudaMemcpy(out, s.ptr, s.bytes, cudaMemcpyDeviceToHost));
        out += s.bytes;
    }
    return EML_OK;
}

int eml_chains_import(EmlChains* ch, const void* blob) {
    if (!ch || !blob) return fail(EML_ERR_INVALID_ARGUMENT, "chains/blob is NULL");
    const unsigned long long* head = (const unsigned long long*)blob;
    if (head[0] != kBlobMagic || head[1] != (unsigned long long)ch->n || head[2] != (unsigned long long)ch->H.n_params)
        return fail(EML_ERR_INVALID_ARGUMENT, "checkpoint does not match this engine (magic / n_chains / n_params)");
    if (head[4] != (unsigned long long)ch->first_chain)
        return fail(EML_ERR_INVALID_ARGUMENT, "checkpoint was written for a different first_chain (chain partition)");
    if (head[5] != layout_hash(ch, ch->h_col_class))
        return fail(EML_ERR_INVALID_ARGUMENT, "checkpoint was written under a different parameter layout or different chain hyperparameters");
    EML_CUDA(cudaSetDevice(ch->plan->device));
    EML_CUDA(cudaDeviceSynchronize());
    ch->seed = head[3];     // the random streams continue from the checkpoint's seed and counters
    const char* in = (const char*)(head + kBlobHead);
    for (const Segment& s : state_segments(ch)) {
        EML_CUDA(cudaMemcpy(s.ptr, in, s.bytes, cudaMemcpyHostToDevice));
        in += s.bytes;
    }
    return EML_OK;
}

}  // extern "C"
