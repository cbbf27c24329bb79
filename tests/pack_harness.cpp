// Host emulation of pack_models_kernel (effective_model_learning_b200/csrc/eml_order.cu): the same sequential core
// (eml_pack.cuh) with the kernel's parallel phases written as plain loops.  Built by tests/test_pack.py with g++.
#include <algorithm>
#include <vector>

#include "eml_pack.cuh"

// the two paths of pack_classes (level bitmask in registers / in memory) on one class histogram: 1 if all tables agree
extern "C" int pack_paths_agree(const int* n_class, int slots, float eps, float tail_fraction) {
    using namespace eml;
    static PackTables a, b;
    pack_classes(n_class, slots, eps, a, false, true, tail_fraction);
    pack_classes(n_class, slots, eps, b, false, false, tail_fraction);
    if (a.cap != b.cap) return 0;
    for (int j = 0; j <= PACK_CLASSES; ++j) if (a.cbeg[j] != b.cbeg[j]) return 0;
    for (int k = 0; k < a.cbeg[PACK_CLASSES]; ++k) if (a.chunk_key[k] != b.chunk_key[k] || a.chunk_end[k] != b.chunk_end[k]) return 0;
    for (int r = 0; r < PACK_LEVELS; ++r) if (a.bins_at[r] != b.bins_at[r]) return 0;
    for (int w = 0; w < PACK_MASK_WORDS; ++w) if (a.mask[w] != b.mask[w]) return 0;
    return a.cap < 128 ? 2 : 1;      // 2: the register path was taken
}

extern "C" int pack_order_host(int n, const float* cost, int slots, float eps, float tail_fraction, int* order_out, int* keys_out) {
    using namespace eml;
    float mx = 0.f;
    for (int b = 0; b < n; ++b) mx = std::max(mx, cost[b]);
    const float scale = (mx > 0.f) ? (PACK_BINS - 1) / mx : 0.f;
    std::vector<int> hist(PACK_BINS, 0), start(PACK_BINS, 0), sorted(n);
    for (int b = 0; b < n; ++b) hist[pack_bin_of(cost[b], scale)]++;
    for (int i = 1; i < PACK_BINS; ++i) start[i] = start[i - 1] + hist[i - 1];
    std::vector<int> fill(start);
    for (int b = 0; b < n; ++b) sorted[fill[pack_bin_of(cost[b], scale)]++] = b;
    int n_class[PACK_CLASSES], s_class[PACK_CLASSES + 1];
    for (int j = 0; j < PACK_CLASSES; ++j) {
        n_class[j] = 0;
        for (int i = 0; i < PACK_BINS / PACK_CLASSES; ++i) n_class[j] += hist[j * (PACK_BINS / PACK_CLASSES) + i];
        s_class[j] = start[j * (PACK_BINS / PACK_CLASSES)];
    }
    s_class[PACK_CLASSES] = n;
    static PackTables t;
    pack_classes(n_class, slots, eps, t, false, true, tail_fraction);
    std::vector<int> key(n), hist2(t.cap + 2, 0), start2(t.cap + 2, 0);
    for (int j = 0; j < PACK_CLASSES; ++j)
        for (int p = s_class[j]; p < s_class[j + 1]; ++p) {
            key[p] = pack_key_of(t, j, p - s_class[j]);
            hist2[key[p]]++;
        }
    for (int k = 1; k < t.cap + 2; ++k) start2[k] = start2[k - 1] + hist2[k - 1];
    for (int p = 0; p < n; ++p) {
        order_out[start2[key[p]]] = sorted[p];
        if (keys_out) keys_out[start2[key[p]]] = key[p];
        start2[key[p]]++;
    }
    return t.cap;
}
