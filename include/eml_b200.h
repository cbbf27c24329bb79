/*
 * eml_b200.h -- C ABI of the B200-native Lindblad model evaluator.
 *
 * This is the drop-in boundary for ONE hot path of henry104413/effective_model_learning:
 *
 *     LearningChain.total_dev(model)            reference learning_chain.py:714-742
 *       -> BasicModel.calculate_dynamics(...)   reference basic_model.py:314-393
 *            -> qutip.mesolve(H, rho0, times, c_ops, e_ops)      basic_model.py:381-388
 *
 * i.e. "parameters of a TLS model -> <O_m>(t_k) for every observable and time -> mean
 * squared deviation from the target data".  The reference has no FFI of its own (it is
 * pure Python calling qutip); these entry points are what a ctypes binding inserted at
 * basic_model.py:381 would call instead of qutip.mesolve.
 *
 * Conventions
 *   - plain C, no C++/torch types; all matrices row-major, complex numbers as
 *     interleaved (re, im) doubles; site 0 is the MOST significant factor of the
 *     Kronecker product (qutip.tensor order, reference definitions.py:28-38).
 *   - every function returns EML_OK (0) or a negative EmlStatus; nothing throws across
 *     the ABI; eml_last_error_string() gives the thread-local message.
 *   - a plan owns device copies of everything shared by a batch (operator library,
 *     rho0, observables, times, targets).  eml_eval_batch() takes DEVICE pointers,
 *     allocates nothing, does not synchronise and launches on the caller's stream.
 *     eml_eval_batch_host() takes HOST pointers and stages through buffers owned by the
 *     plan (host<->device copies included), returning after the results are on the host;
 *     page-locked caller buffers (cudaHostAlloc / cudaHostRegister) are used for the DMA directly.
 *   - a plan is bound to one device; calls on one plan must not overlap in time
 *     (different plans / devices are independent).
 */
#ifndef EML_B200_H
#define EML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EML_VERSION_MAJOR 0
#define EML_VERSION_MINOR 1

typedef enum EmlStatus {
    EML_OK = 0,
    EML_ERR_INVALID_ARGUMENT = -1,
    EML_ERR_CUDA = -2,
    EML_ERR_UNSUPPORTED = -3,   /* e.g. n_tls outside 1..5 */
    EML_ERR_NOT_CONVERGED = -4, /* a model did not converge: series cap hit or non-finite generator (reported by the host entry point) */
    EML_ERR_NO_DEVICE = -5
} EmlStatus;

/* How one library process enters the generator.  `param` selects the column of the
 * per-model parameter vector that carries its value v (0.0 = process absent,
 * reference basic_model.py:575-765 `vectorise_under_library`). */
typedef enum EmlTermKind {
    EML_TERM_ENERGY = 0,   /* H += v   * op_a@site_a                     basic_model.py:206-221 */
    EML_TERM_COUPLING = 1, /* H += v^2 * op_a@site_a (x) op_b@site_b     basic_model.py:226-251 (strength on both factors) */
    EML_TERM_LINDBLAD = 2  /* c  = sqrt(v) * op_a@site_a                 basic_model.py:152-189 */
} EmlTermKind;

typedef struct EmlTerm {
    int32_t param;  /* column of params[b][*] */
    int32_t kind;   /* EmlTermKind */
    int32_t site_a; /* 0 .. n_tls-1 */
    int32_t site_b; /* coupling partner, else -1 */
    int32_t op_a;   /* row of op_table */
    int32_t op_b;   /* row of op_table (couplings), else -1 */
} EmlTerm;

/* Everything shared by the models of a batch.  All pointers are HOST pointers and are
 * copied by eml_plan_create. */
typedef struct EmlProblemDesc {
    int32_t n_tls;    /* number of two-level systems n; d = 2^n; 1 <= n <= 5 */
    int32_t n_params; /* P: length of one model's parameter vector */
    int32_t n_terms;
    const EmlTerm* terms; /* [n_terms] */
    int32_t n_ops;
    const double* op_table; /* [n_ops][8]: 2x2 complex, row-major (definitions.py:42-61) */
    const double* rho0;     /* [d*d][2]: initial density matrix (basic_model.py:259-281) */
    int32_t obs_site;       /* site the observables act on (first qubit, basic_model.py:365-375) */
    int32_t n_obs;          /* K, 1..8 */
    const double* obs;      /* [K][8]: 2x2 complex operators */
    int32_t n_times;        /* T >= 1; rho0 sits at times[0] (qutip semantics) */
    const double* times;    /* [T], non-decreasing */
    int32_t n_target_sets;  /* S >= 0 */
    const double* targets;  /* [S][K][T][2] complex targets, or NULL when S == 0 */
} EmlProblemDesc;

typedef struct EmlPlan EmlPlan; /* opaque */

/* Tunables of the evaluator (0 / negative = keep default). */
typedef struct EmlOptions {
    double theta_target; /* Taylor-series kernels: norm-bound * step targeted per expansion (default and maximum 16) */
    double tol;          /* Taylor-series kernels: absolute per-element stopping tolerance (default 1e-11); the
                            Chebyshev kernels truncate at |J_k| < 1e-14 regardless */
    int32_t kernel;      /* EML_KERNEL_AUTO or a specific kernel, for tests/benchmarks */
    int32_t reserved;    /* flags; bit 0 (EML_FLAG_NO_SECTOR_REDUCTION): always evolve the whole density matrix, even when
                            the parity-even half of rho0 is a steady state of the library (n_tls = 4 Chebyshev kernel) */
} EmlOptions;

#define EML_FLAG_NO_SECTOR_REDUCTION 1

#define EML_KERNEL_AUTO 0
#define EML_KERNEL_GENERIC 1 /* one CTA per model, rho in shared memory, FP64 FMA */
#define EML_KERNEL_WARP_MMA 2 /* one warp per model, rho in registers, FP64 DMMA (n_tls <= 4), Taylor series with dense output */
#define EML_KERNEL_CTA_MMA 3  /* four warps per model, FP64 DMMA, terms exchanged through shared memory (n_tls == 5) */
#define EML_KERNEL_WARP_CHEB 4 /* the warp kernel with the Chebyshev (Bessel-coefficient) series: one recurrence serves
                                  every output time of a macro step; the default for n_tls <= 4 */
#define EML_KERNEL_CTA_CHEB 5  /* the four-warp kernel with the Chebyshev series; the default for n_tls == 5 */

int eml_version(void);
const char* eml_last_error_string(void);
int eml_device_count(int* count);

int eml_plan_create(const EmlProblemDesc* desc, int device, const EmlOptions* opts, EmlPlan** out);
int eml_plan_destroy(EmlPlan* plan);
/* replace the targets of an existing plan (same K, T); host pointer */
int eml_plan_set_targets(EmlPlan* plan, int32_t n_target_sets, const double* targets);

/*
 * Evaluate n_models models.  DEVICE pointers:
 *   params      [n_models][P]            parameter vectors
 *   target_set  [n_models] or NULL       which target set each model is scored against (NULL = set 0)
 *   expect      [n_models][K][T][2] or NULL   <O_m>(t_k), complex
 *   loss        [n_models] or NULL       sum_m sum_k |expect - target|^2 / (T*K)   (learning_chain.py:738-742)
 *   status      [n_models] or NULL       >= 0: number of generator applications (series terms) the model
 *                                        needed; -1: not converged (series cap hit, or a non-finite / absurd generator)
 * `stream` is a cudaStream_t passed as void*.
 * No host synchronisation.  The Chebyshev warp kernels (n_tls <= 4) run a prepare kernel first (assembly of the generators,
 * spectral bounds, series plan, exact cost per model for the longest-first order) whose per-model records live in a
 * workspace owned by the plan: the first call with a larger batch than any before grows it with one cudaMalloc (capped at
 * 65536 models; larger batches run in chunks), every later call allocates nothing.
 */
int eml_eval_batch(EmlPlan* plan, int64_t n_models, const double* params, const int32_t* target_set,
                   double* expect, double* loss, int32_t* status, void* stream);

/* Same with HOST pointers; copies in, launches, copies out and synchronises. */
int eml_eval_batch_host(EmlPlan* plan, int64_t n_models, const double* params, const int32_t* target_set,
                        double* expect, double* loss);

/* Number of kernels launched through this library since load (for bench accounting). */
int64_t eml_launch_count(void);
/* Profiling builds only (-DEML_PROFILE_PHASES; zeros otherwise): warp-cycles the d <= 16 Chebyshev evaluator spent per
 * phase since the last reset, summed over warps -- [0] assembly + bounds, [1] plan + first output, [2] recurrence,
 * [3] output sums, [4] epilogue.  Current device. */
int eml_debug_phase_clocks(uint64_t* out8, int reset);
/* Name of the kernel the plan dispatches to ("generic", "warp_mma", ...). */
const char* eml_plan_kernel_name(const EmlPlan* plan);
/* FP64 microbenchmarks used to establish the roofline denominator (flops per second). */
int eml_measure_fp64_peak(int device, int use_mma, double seconds, double* flops_per_s);

/* ------------------------------------------------------------------------------------------------
 * Batched reversible-jump MCMC over device-resident parameter vectors ("next" row N1 of the scope table).
 * One chain = one row of the plan's parameter layout; a step is  propose -> eml_eval_batch -> accept/reject,
 * entirely on the device.  Per-chain semantics follow LearningChain.run (reference learning_chain.py:346-596):
 * move choice with probability  priority x number-of-ways, tweak = Gaussian jitter with <= 10 redraws into the
 * class bounds (learning_model.py:112-194), additions uniform in the bounds (process_handling.py:98-103),
 * acceptance exp(-dLoss/T) x exp(-dComplexity/complexity_factor) x p_back/p_there (learning_chain.py:792-805),
 * acceptance-window adaptation of the jump lengths (:374-388), shock annealing (:349-367).
 * Random numbers: counter-based splitmix64 stream per chain (reproducible; NOT numpy's legacy stream --
 * reference-trajectory parity is the job of the single-chain LearningChain class).
 */
typedef enum EmlColumnClass {
    EML_COL_DEFECT_ENERGY = 0, /* always present, tweaked with the 'energies' jump length */
    EML_COL_Q2D_COUPLING = 1,
    EML_COL_D2D_COUPLING = 2,
    EML_COL_QUBIT_L = 3,
    EML_COL_DEFECT_L = 4,
    EML_COL_FIXED = 5          /* e.g. qubit energies: never changed */
} EmlColumnClass;

typedef struct EmlChainHyper {
    int32_t n_params;
    int32_t acceptance_window;  /* <= 64 */
    const int32_t* col_class;   /* [n_params] EmlColumnClass (host pointer, copied) */
    double priority[9];         /* tweak, add qubit L, remove qubit L, add defect L, remove defect L,
                                   add q-d coupling, remove q-d coupling, add d-d coupling, remove d-d coupling */
    double temperature;         /* MH temperature, or gamma shape when temperature_scale > 0 */
    double temperature_scale;   /* 0: fixed temperature */
    double complexity_factor;
    double bounds_lo[3];        /* parameter classes: 0 couplings, 1 energies, 2 Ls */
    double bounds_hi[3];
    double jump_initial[3];
    double jump_annealed[3];
    double jump_rescale;        /* jump_length_rescaling_factor */
    double acceptance_target;
    double acceptance_band;
    int64_t shock_anneal_at;    /* <= 0: never */
} EmlChainHyper;

#define EML_CHAIN_STATS 8 /* current_loss, best_loss, acc_tweak, tot_tweak, acc_RJ, tot_RJ, last_window_acceptance, n_processes */

typedef struct EmlChains EmlChains; /* opaque */

/* initial_params: HOST [n_chains][P]; target_set: HOST [n_chains] or NULL.  Evaluates the initial models.
 * Chain c draws from the random stream of global chain number first_chain + c, so a sweep gives the same
 * results however its chains are partitioned over engines and GPUs. */
int eml_chains_create(EmlPlan* plan, const EmlChainHyper* hyper, int64_t n_chains, const double* initial_params,
                      const int32_t* target_set, uint64_t seed, int64_t first_chain, EmlChains** out);
int eml_chains_destroy(EmlChains* chains);
/* n_steps proposals per chain, asynchronous on `stream`.  Optional DEVICE history buffers [n_steps][n_chains]:
 * proposal loss, acceptance probability, code = move << 1 | accepted; and hist_params [n_steps][n_chains][P]:
 * the chain's current parameter vector after the step (the accepted models the reference keeps in
 * explored_proposals, learning_chain.py:585-586, are the rows whose code has bit 0 set). */
int eml_chains_run(EmlChains* chains, int64_t n_steps, double* hist_loss, double* hist_aprob, int32_t* hist_code,
                   double* hist_params, void* stream);
/* Synchronises and copies the chain state to HOST arrays (any may be NULL):
 * current/best params [n_chains][P], current/best loss [n_chains], stats [n_chains][EML_CHAIN_STATS]. */
int eml_chains_read(EmlChains* chains, double* current_params, double* current_loss, double* best_params,
                    double* best_loss, double* stats);
/* Device pointer of the per-chain statistics table [n_chains][EML_CHAIN_STATS] (refreshed by every run), for
 * the checkpoint all_gather. */
const double* eml_chains_stats_device(EmlChains* chains);
/* Proposals that were force-rejected because their evaluation did not converge (status < 0) or gave a non-finite loss
 * (synchronises): per_chain [n_chains] and / or the total; either may be NULL.  Such a proposal is never accepted and never
 * becomes a chain's best model; the reference (qutip.mesolve with nsteps = 1e9, basic_model.py:386) would have
 * integrated it -- a documented deviation. */
int eml_chains_bad_evaluations(EmlChains* chains, int32_t* per_chain, int64_t* total);
/* Checkpoint / resume: the complete chain state (models, losses, jump lengths, acceptance windows, counters,
 * random-stream positions) as one opaque host blob.  Importing it into an engine created with the same plan layout,
 * hyperparameters, n_chains, seed and first_chain continues the chains exactly where they were. */
size_t eml_chains_state_bytes(const EmlChains* chains);
int eml_chains_export(EmlChains* chains, void* host_blob);
int eml_chains_import(EmlChains* chains, const void* host_blob);

#ifdef __cplusplus
}
#endif
#endif /* EML_B200_H */
