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

EML_HD int pack_ctz(const unsigned int word) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)word) - 1;
#else
    return __builtin_ctz(word);
#endif
}

// Best-fit decreasing over the class histogram n_class[PACK_CLASSES] into `slots` bins.  Sequential (one thread).
// cleared: bins_at and mask are already zero (the kernel clears them with all its threads).
// Capacities below 128 units (a few models per bin: the case the packing exists for) keep the level bitmask in four
// registers, so that a step costs one dependent shared-memory load (bins_at[r]) instead of three or four;
// `registers_path = false` forces the general path (the host test compares the two).
// tail_fraction: the cheapest classes holding at most this fraction of the models are NOT packed; they stay at the end of
// the list (longest first) and are fetched greedily, which absorbs the run-time noise of the packed part
// (tools/schedule_model.py: with 5 % noise 0.87 busy without a tail, 0.90 with 10 %).
EML_HD void pack_classes(const int* n_class, const int slots, const float eps, PackTables& t, const bool cleared = false,
                         const bool registers_path = true, const float tail_fraction = 0.f) {
    int n_all = 0;
    for (int j = 0; j < PACK_CLASSES; ++j) n_all += n_class[j];
    int j_cut = PACK_CLASSES, n_tail = 0;                // classes j_cut .. PACK_CLASSES - 1 form the greedy tail
    const int tail_max = (int)(tail_fraction * (float)n_all);
    while (j_cut > 1 && n_tail + n_class[j_cut - 1] <= tail_max) { --j_cut; n_tail += n_class[j_cut]; }
    long long total = 0;
    for (int j = 0; j < j_cut; ++j) total += (long long)n_class[j] * (PACK_CLASSES - j);
    long long cap = (long long)((double)total * (1.0 + (double)eps) / (double)(slots > 0 ? slots : 1)) + 1;
    if (cap < PACK_CLASSES) cap = PACK_CLASSES;          // the most expensive model fits an empty bin
    if (cap > PACK_LEVELS - 1) cap = PACK_LEVELS - 1;
    t.cap = (int)cap;
    if (!cleared) {
        for (int r = 0; r < PACK_LEVELS; ++r) t.bins_at[r] = 0;
        for (int w = 0; w < PACK_MASK_WORDS; ++w) t.mask[w] = 0u;
    }
    t.bins_at[cap] = slots;
    int nchunks = 0;
    if (registers_path && cap < 128) {
        unsigned int m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;      // bit r of (m3 m2 m1 m0) set <=> bins_at[r] > 0
        auto flip = [&](const int r, const bool on) {
            const unsigned int bit = 1u << (r & 31);
            const int w = r >> 5;
            if (on) { m0 |= (w == 0) ? bit : 0u; m1 |= (w == 1) ? bit : 0u; m2 |= (w == 2) ? bit : 0u; m3 |= (w == 3) ? bit : 0u; }
            else    { m0 &= (w == 0) ? ~bit : ~0u; m1 &= (w == 1) ? ~bit : ~0u; m2 &= (w == 2) ? ~bit : ~0u; m3 &= (w == 3) ? ~bit : ~0u; }
        };
        flip((int)cap, true);
        for (int j = 0; j < PACK_CLASSES; ++j) {
            t.cbeg[j] = nchunks;
            const int c = PACK_CLASSES - j;
            const int cw = c >> 5;
            const unsigned int high = 0xffffffffu << (c & 31);
            int rem = (j < j_cut) ? n_class[j] : 0, placed = 0;
            while (rem > 0 && nchunks < PACK_MAX_CHUNKS) {
                // lowest set bit at or above c
                const unsigned int w0 = (cw == 0) ? (m0 & high) : 0u;
                const unsigned int w1 = (cw == 1) ? (m1 & high) : (cw < 1 ? m1 : 0u);
                const unsigned int w2 = (cw == 2) ? (m2 & high) : (cw < 2 ? m2 : 0u);
                const unsigned int w3 = (cw == 3) ? (m3 & high) : (cw < 3 ? m3 : 0u);
                int r;
                if (w0) r = pack_ctz(w0);
                else if (w1) r = 32 + pack_ctz(w1);
                else if (w2) r = 64 + pack_ctz(w2);
                else if (w3) r = 96 + pack_ctz(w3);
                else break;
                const int have = t.bins_at[r];
                const int m = rem < have ? rem : have;
                placed += m;
                t.chunk_key[nchunks] = (unsigned short)(t.cap - r);
                t.chunk_end[nchunks] = placed;
                ++nchunks;
                if (have == m) flip(r, false);
                t.bins_at[r] = have - m;
                t.bins_at[r - c] += m;
                flip(r - c, true);
                rem -= m;
            }
        }
        t.mask[0] = m0; t.mask[1] = m1; t.mask[2] = m2; t.mask[3] = m3;
        t.cbeg[PACK_CLASSES] = nchunks;
        return;
    }
    t.mask[cap >> 5] |= 1u << (cap & 31);
    for (int j = 0; j < PACK_CLASSES; ++j) {
        t.cbeg[j] = nchunks;
        const int c = PACK_CLASSES - j;
        int rem = (j < j_cut) ? n_class[j] : 0, placed = 0;
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
