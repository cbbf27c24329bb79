// Packing the models of one launch onto the persistent warps -- the sequential core, shared by the device kernel
// (eml_order.cu) and the host unit test (tests/pack_harness.cpp compiles this header with g++).
//
// The d = 16 evaluator hands models to S = 148 x 12 resident warps through one atomic counter.  With 2.3 models per warp
// slot (4096 chains) a longest-first list leaves 18 % of the warp-slot time idle even with exact costs (DESIGN.md
// section 5, tools/schedule_model.py): the slots that drew a long first model end far later than the others.  A
// best-fit-decreasing PACKING of the costs into S bins of capacity (1 + eps) * mean load fixes that; the models are
// then LISTED BY THEIR START TIME INSIDE THEIR BIN, so that the dynamic fetch reproduces the packing when the cost
// model is right and still balances greedily where it is not.
//
// Costs are quantised to PACK_CLASSES classes (class j costs PACK_CLASSES - j units, j = 0 the most expensive) and the
// packing moves whole groups: one step places min(#models left in the class, #bins at the best-fitting residual
// level) models at once -- about 90 sequential steps for 4096 models (tools/schedule_model.py: 0.94 busy with 64
// classes, 0.82 for the longest-first list).
#pragma once

#if defined(__CUDACC__)
#define EML_HD __host__ __device__ __forceinline__
#else
#define EML_HD inline
#endif

namespace eml {

constexpr int PACK_CLASSES = 64;
constexpr int PACK_LEVELS = 1024;        // residual levels 0 .. cap, cap < PACK_LEVELS
constexpr int PACK_MAX_CHUNKS = 1024;
constexpr int PACK_MAX_MODELS = 16384;   // beyond that (or beyond 8 models per slot) a longest-first list is good enough
constexpr int PACK_MASK_WORDS = PACK_LEVELS / 32;

struct PackTables {
    int bins_at[PACK_LEVELS];                   // number of bins whose residual capacity is exactly r units
    unsigned int mask[PACK_MASK_WORDS];         // bit r set <=> bins_at[r] > 0
    unsigned short chunk_key[PACK_MAX_CHUNKS];  // start time (units) of the models of a chunk inside their bins
    int chunk_end[PACK_MAX_CHUNKS];             // cumulative number of models of the class placed up to this chunk
    int cbeg[PACK_CLASSES + 1];                 // chunks of class j: cbeg[j] .. cbeg[j + 1] - 1
    int cap;                                    // bin capacity in units; key cap + 1 marks "did not fit"
};

// 256 fine bins for the longest-first order inside a class (bin 0 = most expensive); class = bin / 4
constexpr int PACK_BINS = 256;
EML_HD int pack_bin_of(const float cost, const float scale) {
    int q = (int)(cost * scale);               // NaN converts to 0
    q = q < 0 ? 0 : (q > PACK_BINS - 1 ? PACK_BINS - 1 : q);
    return PACK_BINS - 1 - q;
}

EML_HD int pack_find_level(const unsigned int* mask, const int from, const int cap) {
    int w = from >> 5;
    unsigned int word = mask[w] & (0xffffffffu << (from & 31));
    const int wlast = cap >> 5;
    while (word == 0u) {
        if (++w > wlast) return -1;
        word = mask[w];
    }
#if defined(__CUDA_ARCH__)
    const int bit = __ffs((int)word) - 1;
#else
    const int bit = __builtin_ctz(word);
#endif
    const int r = 32 * w + bit;
    return r <= cap ? r : -1;
}

// Best-fit decreasing over the class histogram n_class[PACK_CLASSES] into `slots` bins.  Sequential (one thread).
// cleared: bins_at and mask are already zero (the kernel clears them with all its threads).
EML_HD void pack_classes(const int* n_class, const int slots, const float eps, PackTables& t, const bool cleared = false) {
    long long total = 0;
    for (int j = 0; j < PACK_CLASSES; ++j) total += (long long)n_class[j] * (PACK_CLASSES - j);
    long long cap = (long long)((double)total * (1.0 + (double)eps) / (double)(slots > 0 ? slots : 1)) + 1;
    if (cap < PACK_CLASSES) cap = PACK_CLASSES;          // the most expensive model fits an empty bin
    if (cap > PACK_LEVELS - 1) cap = PACK_LEVELS - 1;
    t.cap = (int)cap;
    if (!cleared) {
        for (int r = 0; r < PACK_LEVELS; ++r) t.bins_at[r] = 0;
        for (int w = 0; w < PACK_MASK_WORDS; ++w) t.mask[w] = 0u;
    }
    t.bins_at[cap] = slots;
    t.mask[cap >> 5] |= 1u << (cap & 31);
    int nchunks = 0;
    for (int j = 0; j < PACK_CLASSES; ++j) {
        t.cbeg[j] = nchunks;
        const int c = PACK_CLASSES - j;
        int rem = n_class[j], placed = 0;
        while (rem > 0 && nchunks < PACK_MAX_CHUNKS) {
            const int r = pack_find_level(t.mask, c, t.cap);
            if (r < 0) break;
            const int have = t.bins_at[r];
            const int m = rem < have ? rem : have;
            placed += m;
            t.chunk_key[nchunks] = (unsigned short)(t.cap - r);
            t.chunk_end[nchunks] = placed;
            ++nchunks;
            if (have == m) t.mask[r >> 5] &= ~(1u << (r & 31));
            t.bins_at[r] = have - m;
            t.bins_at[r - c] += m;
            t.mask[(r - c) >> 5] |= 1u << ((r - c) & 31);
            rem -= m;
        }
    }
    t.cbeg[PACK_CLASSES] = nchunks;
}

// start key of the model at offset `off` inside class j (the models of a class are interchangeable)
EML_HD int pack_key_of(const PackTables& t, const int j, const int off) {
    for (int k = t.cbeg[j]; k < t.cbeg[j + 1]; ++k)
        if (off < t.chunk_end[k]) return (int)t.chunk_key[k];
    return t.cap + 1;
}

}  // namespace eml
