// Internal definitions shared by the translation units behind the C ABI.
#pragma once
#include <string>
#include <vector>

#include "eml_common.cuh"

std::string& eml_last_error_ref();
inline int fail(int code, const std::string& msg) {
    eml_last_error_ref() = msg;
    return code;
}
#define EML_CUDA(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess)                                                                                  \
            return fail(EML_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
    } while (0)

struct EmlPlan {
    int device = 0;
    eml::DevProblem P{};
    eml::DevProblem Pw{};     // the same problem as the warp kernel sees it (n_tls < 3 embedded into 3 sites)
    int kernel = EML_KERNEL_GENERIC;
    bool hermitian = false;   // H (every coupling term set) and rho0 Hermitian -> specialised kernels allowed
    bool real_h = false;      // additionally every Hamiltonian term is a real matrix (G.re diagonal): half the DMMAs
    bool parity_h = false;    // additionally every Hamiltonian term conserves the parity of the basis index: G block diagonal
    bool sector_b = false;    // additionally the parity-even part of rho0 is a steady state (d = 16): only the odd tiles evolve
    std::vector<void*> owned; // device allocations
    // staging for the host entry point
    cudaStream_t stream = nullptr;
    long long cap_models = 0;
    double *d_params = nullptr, *d_expect = nullptr, *d_loss = nullptr;
    int *d_tset = nullptr, *d_status = nullptr;
    double *h_params = nullptr, *h_expect = nullptr, *h_loss = nullptr;  // pinned
    int *h_tset = nullptr, *h_status = nullptr;
    double* d_targets = nullptr;
    size_t targets_bytes = 0;
    unsigned int* d_counter = nullptr;  // work-fetch counter of the persistent warp kernel (EML_FETCH_WS_WORDS words)
    int* d_order = nullptr;             // ordering workspace (EML_ORDER_WS_INTS ints)
    double* d_prep = nullptr;           // records of the prepare kernel (Chebyshev warp kernels), prep_cap models
    long long prep_cap = 0;
    // Small host batches (a single chain's total_dev: ONE model per call) are bound by the launch calls of the host:
    // upload, memset, prepare, evaluator, downloads are captured once per call shape as a CUDA graph and replayed.
    struct HostGraph { long long n; int has_tset, has_expect, has_loss; cudaGraphExec_t exec; };
    std::vector<HostGraph> host_graphs;
    int host_calls_small = 0;           // small-batch calls seen (the first ones run eagerly: lazy initialisations)
    void drop_host_graphs() {
        for (HostGraph& g : host_graphs) cudaGraphExecDestroy(g.exec);
        host_graphs.clear();
    }
};

// words of a work-fetch workspace: the evaluator's counter ([0]) and, 16-byte aligned behind it, the counts of the 256 cost
// bins the prepare kernel sorts the models of a launch into (eml_warp_mma.cu: ORDER_HIST_OFF, ORDER_BINS)
constexpr size_t EML_FETCH_WS_WORDS = 4 + 256;
// ints of an ordering workspace: the sorted model order (2^18) + float cost scratch (2^18) of large launches, or the 256
// cost-bin lists of 4096 entries of launches up to 4096 models
constexpr size_t EML_ORDER_WS_INTS = (size_t)1 << 20;

int eml_eval_batch_hinted(EmlPlan* plan, int64_t n_models, const double* params, const int32_t* target_set, double* expect,
                          double* loss, int32_t* status, void* stream, const int32_t* cost_hint, unsigned int* counter_ws,
                          int* order_ws, double* prep_ws, long long prep_cap);
size_t eml_plan_prep_bytes_per_model(const EmlPlan* plan);
