// ba_plan.h -- host-side batch planner shared by the CUDA library and the test emulator.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

#include "ba_common.cuh"
#include "ba_fill16.cuh"

namespace {

struct Plan {
    std::vector<WorkItem> items;        // all items, grouped by class then size
    std::vector<int64_t> aux_off;       // per item (Fitted kinds, 32-bit path)
    std::vector<int64_t> slot_off;      // per item (16-bit path): first segment slot
    std::vector<int32_t> slot_cap;      // per item (16-bit path)
    struct Launch {
        int cols;
        int64_t first, count;           // item range
    };
    struct Wave {
        int64_t first, count;           // item range
        int64_t code_bytes, aux_ints, slots;
        std::vector<Launch> launches;
        int max_n_multipass;            // largest n among pairs needing >1 pass (0 = none)
    };
    std::vector<Wave> waves;
};

inline int cols_for(int m)
{
    if (m <= 0) return 1;
    int c = (m + 31) / 32;
    return c > BA_MAX_COLS ? BA_MAX_COLS : c;
}

// capacity of the per-pair segment slot buffer of the 16-bit path; pairs that emit more are
// re-run on the exact two-pass path
inline int32_t slot_cap_for(int n, int m)
{
    int64_t c = ((int64_t)n + m) / 12 + 12;
    const int64_t hi = (int64_t)n + m + 2;
    return (int32_t)(c < hi ? c : hi);
}

// Length-bucketed work list over the pairs `ids`: grouped by the number of columns a lane
// owns (the template instance they need), largest-first inside a group so the dynamic work
// counter balances the tail, then cut into waves whose traceback data fits the arena.
// mode16: checkpoint bytes of ba_fill16 instead of the code bytes of ba_fill32.
inline void build_plan(const int32_t *ids, int64_t nids, const int32_t *ref_len, const int32_t *qry_len, int kind,
                       int affine, bool mode16, int64_t arena_cap, Plan &plan)
{
    std::vector<std::vector<int32_t>> by_class(BA_MAX_COLS + 1);
    std::vector<char> uniform(BA_MAX_COLS + 1, 1);
    std::vector<int64_t> first_cells(BA_MAX_COLS + 1, -1);
    for (int64_t x = 0; x < nids; x++) {
        const int32_t p = ids ? ids[x] : (int32_t)x;
        const int c = cols_for(qry_len[p]);
        const int64_t cells = (int64_t)ref_len[p] * (int64_t)qry_len[p];
        if (first_cells[c] < 0)
            first_cells[c] = cells;
        else if (first_cells[c] != cells)
            uniform[c] = 0;
        by_class[c].push_back(p);
    }
    plan.items.reserve((size_t)nids);
    const bool need_aux = (kind == KIND_FIT) && !mode16;
    if (need_aux) plan.aux_off.reserve((size_t)nids);
    if (mode16) {
        plan.slot_off.reserve((size_t)nids);
        plan.slot_cap.reserve((size_t)nids);
    }
    Plan::Wave wave;
    wave.first = 0;
    wave.count = 0;
    wave.code_bytes = 0;
    wave.aux_ints = 0;
    wave.slots = 0;
    wave.max_n_multipass = 0;
    auto flush = [&]() {
        if (wave.count > 0) plan.waves.push_back(wave);
        wave.first += wave.count;
        wave.count = 0;
        wave.code_bytes = 0;
        wave.aux_ints = 0;
        wave.slots = 0;
        wave.max_n_multipass = 0;
        wave.launches.clear();
    };
    for (int c = BA_MAX_COLS; c >= 1; c--) {
        auto &v = by_class[c];
        if (v.empty()) continue;
        if (!uniform[c]) {
            std::stable_sort(v.begin(), v.end(), [&](int32_t a, int32_t b) {
                return (int64_t)ref_len[a] * qry_len[a] > (int64_t)ref_len[b] * qry_len[b];
            });
        }
        Plan::Launch L;
        L.cols = c;
        L.first = wave.first + wave.count;
        L.count = 0;
        for (int32_t p : v) {
            int64_t cb = mode16 ? ba16_ck_bytes(c, ref_len[p], qry_len[p])
                                : ba_code_bytes(affine, c, ref_len[p], qry_len[p]);
            cb = (cb + 255) & ~(int64_t)255;
            if (wave.count > 0 && wave.code_bytes + cb > arena_cap) {
                if (L.count > 0) wave.launches.push_back(L);
                flush();
                L.first = wave.first;
                L.count = 0;
            }
            WorkItem it;
            it.pair = p;
            it.cols = c;
            it.code_off = wave.code_bytes;
            plan.items.push_back(it);
            if (need_aux) {
                plan.aux_off.push_back(wave.aux_ints);
                wave.aux_ints += (int64_t)ref_len[p] + 1;
            }
            if (mode16) {
                const int32_t cap = slot_cap_for(ref_len[p], qry_len[p]);
                plan.slot_off.push_back(wave.slots);
                plan.slot_cap.push_back(cap);
                wave.slots += cap;
            }
            wave.code_bytes += cb;
            wave.count++;
            L.count++;
            if (qry_len[p] > 32 * c && ref_len[p] > wave.max_n_multipass) wave.max_n_multipass = ref_len[p];
        }
        if (L.count > 0) wave.launches.push_back(L);
    }
    flush();
}

}  // namespace
