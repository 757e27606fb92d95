// ba_plan.h -- host-side batch planner shared by the CUDA library and the test emulator.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

#include "ba_common.cuh"
#include "ba_fill16.cuh"

namespace {

struct Plan {
    // per-item arrays live in caller-provided (pinned) buffers of `nids` entries each
    WorkItem *items = nullptr;          // all items, grouped by class then size
    int64_t *aux_off = nullptr;         // per item (Fitted kinds, 32-bit path)
    int64_t *slot_off = nullptr;        // per item (16-bit path): first segment slot
    int32_t *slot_cap = nullptr;        // per item (16-bit path)
    struct Launch {
        int cols;
        int64_t first, count;           // item range
    };
    struct Wave {
        int64_t first, count;           // item range
        int64_t code_bytes, aux_ints, slots;
        std::vector<Launch> launches;
        int max_n_multipass;            // largest n among pairs needing >1 pass (0 = none)
        int32_t max_pair;               // largest pair id of the wave (which upload groups it needs)
    };
    std::vector<Wave> waves;
};

// Columns per lane (32-bit path) / per strip (16-bit path): the narrowest class that covers the
// query in the fewest passes, e.g. m = 300 -> two passes of 32*5 = 160 columns instead of 2 x 256.
inline int cols_for(int m)
{
    if (m <= 0) return 1;
    const int npass = (m + 32 * BA_MAX_COLS - 1) / (32 * BA_MAX_COLS);
    const int c = (m + 32 * npass - 1) / (32 * npass);
    return c > BA_MAX_COLS ? BA_MAX_COLS : c;
}

// 16-bit path: queries of 257..320 letters get ONE pass of 10-column strips (two strips per lane, 20 columns)
// instead of two passes of five -- half the steps, no inter-pass boundary rows, 1.08 instead of 1.48 bytes of
// checkpoints per cell (BASELINE config 3, 300 aa x 300 aa: 1.05 -> 1.28 TCUPS).  Using 10-column strips for
// multi-pass plans as well (513..640 letters in two passes instead of three, 10 kb in 32 instead of 40) was
// measured and dropped: the fill of config 4 gains 20 % but its traceback over 10-column tiles loses as much, and
// config 5 loses 9 % overall.  wide10 = 0 keeps the 8-column ceiling.
inline int cols16_for(int m, bool wide10)
{
    if (wide10 && m > 32 * BA_MAX_COLS && m <= 32 * BA16_MAX_COLS) return BA16_MAX_COLS;
    return cols_for(m);
}

// capacity of the per-pair segment slot buffer of the 16-bit path; pairs that emit more are
// re-run on the exact two-pass path
inline int32_t slot_cap_for(int n, int m, int affine)
{
    // linear gaps scatter: an indel costs the same wherever it opens
    int64_t c = affine ? ((int64_t)n + m) / 12 + 12 : ((int64_t)n + m) / 5 + 16;
    const int64_t hi = (int64_t)n + m + 2;
    return (int32_t)(c < hi ? c : hi);
}

// Length-bucketed work list over the pairs `ids`: grouped by the number of columns a lane
// owns (the template instance they need), largest-first inside a group so the dynamic work
// counter balances the tail, then cut into waves whose traceback data fits the arena.
// mode16: checkpoint bytes of ba_fill16 instead of the code bytes of ba_fill32.
// target_waves > 1 (16-bit path): additionally cut the batch into about that many waves of equal
// traceback bytes, so that the traceback of wave w overlaps the fill of wave w+1 (ba_api.cu).
inline void build_plan(const int32_t *ids, int64_t nids, const int32_t *ref_len, const int32_t *qry_len, int kind,
                       int affine, bool mode16, int64_t arena_cap, Plan &plan, int target_waves = 1, bool wide10 = true)
{
    auto cols_of = [&](int m) { return mode16 ? cols16_for(m, wide10) : cols_for(m); };
    const bool need_aux = (kind == KIND_FIT) && !mode16;
    if (target_waves < 1) target_waves = 1;
    // fast path: every pair has the same shape (the common batch): closed-form plan
    if (nids > 0) {
        const int32_t p0 = ids ? ids[0] : 0;
        const int n0 = ref_len[p0], m0 = qry_len[p0];
        bool same = true;
        for (int64_t x = 1; x < nids && same; x++) {
            const int32_t p = ids ? ids[x] : (int32_t)x;
            same = (ref_len[p] == n0) & (qry_len[p] == m0);
        }
        if (same) {
            const int c = cols_of(m0);
            // 16-bit path: cb is a COUPLE's region (two pairs); 32-bit path: one pair's codes
            int64_t cb = mode16 ? ba16_ck_bytes(c, n0, ba16_npass(c, m0), !affine) : ba_code_bytes(affine, c, n0, m0);
            cb = (cb + 255) & ~(int64_t)255;
            const int unit = mode16 ? 2 : 1;   // items per cb
            const int32_t cap = slot_cap_for(n0, m0, affine);
            int64_t per_wave = cb > 0 ? (arena_cap / cb) * unit : nids;
            if (target_waves > 1) {
                // equal waves, whole warp claims (4 items) each
                int64_t even = (nids + target_waves - 1) / target_waves;
                even = (even + 7) & ~(int64_t)7;
                if (even < per_wave) per_wave = even;
            }
            if (per_wave < unit) per_wave = unit;
            if (per_wave > nids) per_wave = nids;
            for (int64_t first = 0; first < nids; first += per_wave) {
                const int64_t count = std::min(per_wave, nids - first);
                Plan::Wave w;
                w.first = first;
                w.count = count;
                w.code_bytes = ((count + unit - 1) / unit) * cb;
                w.aux_ints = need_aux ? count * ((int64_t)n0 + 1) : 0;
                w.slots = mode16 ? count * (int64_t)cap : 0;
                w.max_n_multipass = (m0 > 32 * c) ? n0 : 0;
                w.max_pair = 0;
                Plan::Launch L;
                L.cols = c;
                L.first = first;
                L.count = count;
                w.launches.push_back(L);
                for (int64_t k = 0; k < count; k++) {
                    WorkItem it;
                    it.pair = ids ? ids[first + k] : (int32_t)(first + k);
                    it.cols = c;
                    it.code_off = (k / unit) * cb;
                    it.n_ck = n0;
                    it.ck = mode16 ? ((int32_t)ba16_npass(c, m0) | (int32_t)((k & 1) << 16)) : 0;
                    plan.items[first + k] = it;
                    if (it.pair > w.max_pair) w.max_pair = it.pair;
                }
                if (need_aux)
                    for (int64_t k = 0; k < count; k++) plan.aux_off[first + k] = k * ((int64_t)n0 + 1);
                if (mode16)
                    for (int64_t k = 0; k < count; k++) {
                        plan.slot_off[first + k] = k * (int64_t)cap;
                        plan.slot_cap[first + k] = cap;
                    }
                plan.waves.push_back(w);
            }
            return;
        }
    }
    // pass 1: class sizes and whether a class is uniform (then no sort is needed)
    int64_t cnt[BA16_MAX_COLS + 1] = {0};
    int64_t first_cells[BA16_MAX_COLS + 1];
    bool uniform[BA16_MAX_COLS + 1];
    for (int c = 0; c <= BA16_MAX_COLS; c++) {
        first_cells[c] = -1;
        uniform[c] = true;
    }
    std::vector<uint8_t> cls((size_t)nids);
    int64_t total_cb = 0;
    int tmemo_n = -1, tmemo_m = -1;
    int64_t tmemo_cb = 0;
    for (int64_t x = 0; x < nids; x++) {
        const int32_t p = ids ? ids[x] : (int32_t)x;
        const int c = cols_of(qry_len[p]);
        cls[(size_t)x] = (uint8_t)c;
        if (target_waves > 1) {
            if (ref_len[p] != tmemo_n || qry_len[p] != tmemo_m) {
                tmemo_n = ref_len[p];
                tmemo_m = qry_len[p];
                tmemo_cb = mode16 ? ba16_ck_bytes(c, tmemo_n, ba16_npass(c, tmemo_m), !affine) / 2
                                  : ba_code_bytes(affine, c, tmemo_n, tmemo_m);
                tmemo_cb = (tmemo_cb + 255) & ~(int64_t)255;
            }
            total_cb += tmemo_cb;
        }
        const int64_t cells = (int64_t)ref_len[p] * (int64_t)qry_len[p];
        if (first_cells[c] < 0)
            first_cells[c] = cells;
        else if (first_cells[c] != cells)
            uniform[c] = false;
        cnt[c]++;
    }
    // pass 2: pair ids grouped by class (widest first), written into one array
    std::vector<int32_t> order((size_t)nids);
    int64_t start[BA16_MAX_COLS + 2];
    {
        int64_t acc = 0;
        for (int c = BA16_MAX_COLS; c >= 1; c--) {
            start[c] = acc;
            acc += cnt[c];
        }
        int64_t fill[BA16_MAX_COLS + 1];
        for (int c = 1; c <= BA16_MAX_COLS; c++) fill[c] = start[c];
        for (int64_t x = 0; x < nids; x++) {
            const int32_t p = ids ? ids[x] : (int32_t)x;
            order[(size_t)fill[cls[(size_t)x]]++] = p;
        }
    }
    // soft cap for the overlap waves (cuts fall on whole warp claims of four items)
    int64_t soft_cap = arena_cap;
    if (target_waves > 1) soft_cap = std::min(arena_cap, total_cb / target_waves + 1);
    Plan::Wave wave;
    wave.first = 0;
    wave.count = 0;
    wave.code_bytes = 0;
    wave.aux_ints = 0;
    wave.slots = 0;
    wave.max_n_multipass = 0;
    wave.max_pair = 0;
    auto flush = [&]() {
        if (wave.count > 0) plan.waves.push_back(wave);
        wave.first += wave.count;
        wave.count = 0;
        wave.code_bytes = 0;
        wave.aux_ints = 0;
        wave.slots = 0;
        wave.max_n_multipass = 0;
        wave.max_pair = 0;
        wave.launches.clear();
    };
    int64_t w = 0;  // next item slot
    for (int c = BA16_MAX_COLS; c >= 1; c--) {
        if (cnt[c] == 0) continue;
        int32_t *v = order.data() + start[c];
        if (!uniform[c]) {
            std::stable_sort(v, v + cnt[c], [&](int32_t a, int32_t b) {
                return (int64_t)ref_len[a] * qry_len[a] > (int64_t)ref_len[b] * qry_len[b];
            });
        }
        Plan::Launch L;
        L.cols = c;
        L.first = wave.first + wave.count;
        L.count = 0;
        int memo_n = -1, memo_m = -1;   // batches are often uniform: reuse the per-shape numbers
        int64_t memo_cb = 0;
        int32_t memo_cap = 0;
        // 16-bit path: items are placed a COUPLE at a time (items 2k, 2k+1 of a launch share one checkpoint
        // region laid out for the larger of the two; the list is sorted by size, so partners are alike)
        const int64_t step = mode16 ? 2 : 1;
        for (int64_t x = 0; x < cnt[c]; x += step) {
            const int32_t p = v[x];
            const bool has2 = mode16 && x + 1 < cnt[c];
            const int32_t p2 = has2 ? v[x + 1] : p;
            const int cn = std::max(ref_len[p], ref_len[p2]), cm = std::max(qry_len[p], qry_len[p2]);
            if (cn != memo_n || cm != memo_m) {
                memo_n = cn;
                memo_m = cm;
                memo_cb = mode16 ? ba16_ck_bytes(c, cn, ba16_npass(c, cm), !affine) : ba_code_bytes(affine, c, cn, cm);
                memo_cb = (memo_cb + 255) & ~(int64_t)255;
            }
            const int64_t cb = memo_cb;
            if (wave.count > 0 && (wave.code_bytes + cb > arena_cap ||
                                   (wave.code_bytes + cb > soft_cap && (L.count & 3) == 0))) {
                if (L.count > 0) wave.launches.push_back(L);
                flush();
                L.first = wave.first;
                L.count = 0;
            }
            for (int h = 0; h < (has2 ? 2 : 1); h++) {
                const int32_t q = h ? p2 : p;
                WorkItem it;
                it.pair = q;
                it.cols = c;
                it.code_off = wave.code_bytes;
                it.n_ck = cn;
                it.ck = mode16 ? ((int32_t)ba16_npass(c, cm) | (h << 16)) : 0;
                plan.items[w] = it;
                if (q > wave.max_pair) wave.max_pair = q;
                if (need_aux) {
                    plan.aux_off[w] = wave.aux_ints;
                    wave.aux_ints += (int64_t)ref_len[q] + 1;
                }
                if (mode16) {
                    memo_cap = slot_cap_for(ref_len[q], qry_len[q], affine);
                    plan.slot_off[w] = wave.slots;
                    plan.slot_cap[w] = memo_cap;
                    wave.slots += memo_cap;
                }
                w++;
                wave.count++;
                L.count++;
                if (qry_len[q] > 32 * c && ref_len[q] > wave.max_n_multipass) wave.max_n_multipass = ref_len[q];
            }
            wave.code_bytes += cb;
        }
        if (L.count > 0) wave.launches.push_back(L);
    }
    flush();
}

}  // namespace
