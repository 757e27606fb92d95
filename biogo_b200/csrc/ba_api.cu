// ba_api.cu -- C ABI of libbiogoalign.so (see include/biogo_align.h) and the host-side
// batch scheduler: scoring upload, length-bucketed work lists, wave sizing against the
// traceback-code arena, kernel launches and result gathering.  Plain CUDA runtime; no
// torch types, no CPU fallback.
#include "../../include/biogo_align.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "ba_common.cuh"
#ifndef BA_EMULATE
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>
#endif
#include "ba_fill32.cuh"
#include "ba_walk.cuh"
#include "ba_fill16.cuh"
#include "ba_trace16.cuh"
#include "ba_plan.h"
#include "ba_format.cuh"

// One spelling for kernel launches so the same scheduler source also drives the CPU
// emulator of the no-GPU test tier (tests/emu).
#ifdef BA_EMULATE
#define BA_LAUNCH(kernel, grid, block, stream, ...) emu::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
#define BA_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
#else
#define BA_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define BA_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct PinnedBlock {
    void *p;
    size_t cap;
    bool used;
};

struct ToInt64 {
    __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};

typedef void (*FillFn)(FillParams);
typedef void (*WalkFn)(WalkParams);

template <int KIND, bool AFFINE>
FillFn fill_for_cols(int cols)
{
    switch (cols) {
    case 1: return ba_fill32_kernel<KIND, AFFINE, 1>;
    case 2: return ba_fill32_kernel<KIND, AFFINE, 2>;
    case 3: return ba_fill32_kernel<KIND, AFFINE, 3>;
    case 4: return ba_fill32_kernel<KIND, AFFINE, 4>;
    case 5: return ba_fill32_kernel<KIND, AFFINE, 5>;
    case 6: return ba_fill32_kernel<KIND, AFFINE, 6>;
    case 7: return ba_fill32_kernel<KIND, AFFINE, 7>;
    default: return ba_fill32_kernel<KIND, AFFINE, 8>;
    }
}

FillFn pick_fill(int kind, int affine, int cols)
{
    if (affine) {
        if (kind == KIND_SW) return fill_for_cols<KIND_SW, true>(cols);
        if (kind == KIND_NW) return fill_for_cols<KIND_NW, true>(cols);
        return fill_for_cols<KIND_FIT, true>(cols);
    }
    if (kind == KIND_SW) return fill_for_cols<KIND_SW, false>(cols);
    if (kind == KIND_NW) return fill_for_cols<KIND_NW, false>(cols);
    return fill_for_cols<KIND_FIT, false>(cols);
}

template <bool WRITE>
WalkFn pick_walk(int kind, int affine)
{
    if (affine) {
        if (kind == KIND_SW) return ba_walk_kernel<KIND_SW, true, WRITE>;
        if (kind == KIND_NW) return ba_walk_kernel<KIND_NW, true, WRITE>;
        return ba_walk_kernel<KIND_FIT, true, WRITE>;
    }
    if (kind == KIND_SW) return ba_walk_kernel<KIND_SW, false, WRITE>;
    if (kind == KIND_NW) return ba_walk_kernel<KIND_NW, false, WRITE>;
    return ba_walk_kernel<KIND_FIT, false, WRITE>;
}

}  // namespace

struct ba_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;
    bool have_scoring = false;
    DeviceScoring h_sc;
    DeviceScoring *d_sc = nullptr;
    int64_t opt_arena_bytes = 0;
    // grow-only device buffers
    DevBuf letters, enc, ref_off, ref_len, qry_off, qry_len, status, err_pos, start;
    DevBuf items, aux_off, aux, bnd, seg_count, seg_base, segs, arena, scan_tmp, counters;
    DevBuf flags, slot_off, slot_cap, slots, seg_clip, tasks, prog;
    // Plan cache: batches of the same shape (same length arrays, same scoring) reuse the split, the
    // work lists and -- for single-wave plans -- their device copies.
    struct PlanCache {
        bool valid = false, plan_built = false;
        int64_t n = 0, cells = 0, arena_cap = 0, slot_cap_opt = 0;
        uint64_t epoch = 0;
        std::vector<int32_t> ref_len, qry_len, fast, slow0;
        Plan plan;
        const void *dev_items = nullptr, *dev_slot_off = nullptr, *dev_slot_cap = nullptr;  // where the device copy lives
    } pc;
    uint64_t scoring_epoch = 1;
    // the batch whose raw letters and offsets are still on the device (ba_format_batch with letters == NULL)
    int64_t last_n = -1, last_letters_bytes = 0;
    int last_stride = 0;
    DevBuf fmt_len, fmt_off, fmt_rows, fmt_segs, fmt_seg_start, fmt_seg_count;
    int64_t opt_plan_cache = 1;
    int64_t opt_force_path = 0, opt_slot_cap = 0, opt_long_mode = 1, opt_trace_mode = 1, opt_exclusive = 1, opt_sym = 1;
    int64_t cached_arena_cap = 0;                    // cudaMemGetInfo is slow: ask once per context
    std::vector<std::pair<const void *, int>> occ_cache;  // kernel -> resident CTAs per SM
    std::vector<PinnedBlock> pinned;
};

namespace {

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            char b_[512];                                                             \
            snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call,       \
                     cudaGetErrorString(e_));                                         \
            ctx->err = b_;                                                            \
            return (e_ == cudaErrorMemoryAllocation) ? BA_ERR_NOMEM : BA_ERR_CUDA;    \
        }                                                                             \
    } while (0)

int ensure(ba_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return BA_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        want = bytes;
        CK(cudaMalloc(&b.p, want));
    }
    b.cap = want;
    return BA_OK;
}

void *pinned_get(ba_ctx *ctx, size_t bytes)
{
    if (bytes == 0) bytes = 8;
    int bestk = -1;
    for (size_t k = 0; k < ctx->pinned.size(); k++) {
        PinnedBlock &b = ctx->pinned[k];
        if (!b.used && b.cap >= bytes && (bestk < 0 || b.cap < ctx->pinned[bestk].cap)) bestk = (int)k;
    }
    if (bestk >= 0) {
        ctx->pinned[bestk].used = true;
        return ctx->pinned[bestk].p;
    }
    void *p = nullptr;
    size_t cap = bytes + bytes / 4;
    if (cudaMallocHost(&p, cap) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    ctx->pinned.push_back({p, cap, true});
    return p;
}

void pinned_put(ba_ctx *ctx, void *p)
{
    for (auto &b : ctx->pinned)
        if (b.p == p) b.used = false;
}

// Work lists travel host -> device through a kernel that reads the pinned host buffer directly
// (zero-copy over PCIe) instead of the copy engine: an H2D copy-engine transfer would queue behind
// the 100+ MB letters upload another context may have in flight (exclusive_compute pipelining).
__global__ void ba_upload_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, int64_t words)
{
    for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < words; x += (int64_t)gridDim.x * blockDim.x)
        dst[x] = src[x];
}
int upload_words(ba_ctx *ctx, void *dst, const void *pinned_src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return BA_OK;
    const int64_t words = (int64_t)(bytes / 4);  // every work-list element is a multiple of 4 bytes
    int64_t blocks = (words + 255) / 256;
    if (blocks > 2 * (int64_t)ctx->sm_count) blocks = 2 * (int64_t)ctx->sm_count;
    BA_LAUNCH(ba_upload_kernel, (unsigned)blocks, 256, st, (uint32_t *)dst, (const uint32_t *)pinned_src, words);
    CK(cudaGetLastError());
    return BA_OK;
}

// seg_base = exclusive prefix sum of seg_count (int32 -> int64)
int exclusive_scan(ba_ctx *ctx, void *tmp, size_t &tmp_bytes, const int32_t *in, int64_t *out, int n,
                   cudaStream_t st)
{
#ifdef BA_EMULATE
    (void)st;
    if (!tmp) {
        tmp_bytes = 16;
        return BA_OK;
    }
    int64_t acc = 0;
    for (int k = 0; k < n; k++) {
        out[k] = acc;
        acc += in[k];
    }
    (void)ctx;
    return BA_OK;
#else
    auto itr = thrust::make_transform_iterator(in, ToInt64());
    CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, itr, out, n, st));
    return BA_OK;
#endif
}

// ---- everything the two execution paths share for one ba_align_batch call
struct Batch {
    const uint8_t *d_enc;
    int stride;
    const int64_t *d_ref_off, *d_qry_off;
    const int32_t *d_ref_len, *d_qry_len;
    const int32_t *h_ref_len, *h_qry_len;
    int64_t n;
    cudaStream_t st;
    int32_t *d_status;
    PairStart *d_start;
    ba_result *out;
    struct WaveOut {
        int32_t *segs;
        int64_t nsegs;
    };
    std::vector<WaveOut> wave_out;
    int64_t seg_total = 0;
    // exclusive_compute: the per-device kernel-phase token this call holds, released as soon as the
    // last kernel of the batch is enqueued (the result copies then overlap the next batch's kernels)
    std::unique_lock<std::mutex> *token = nullptr;
    bool more_kernels_later = false;   // pairs are already queued for the exact path after this one
    float ms_d2h = 0.f, ms_fill = 0.f, ms_trace = 0.f;
    double host_s = 0.0;
    int64_t launches = 0, fill_launches = 0, code_bytes = 0;
};

static bool g_timing = getenv("BA_TIMING") != nullptr;
// One batch's kernel phase at a time per device (option "exclusive_compute"): callers driving several
// contexts from several threads then pipeline -- the H2D copy of one batch runs under the kernels
// of another instead of all copies and all kernels coinciding.
static std::mutex g_compute_token[64];
#define BA_TSEC(name, t0) do { if (g_timing) fprintf(stderr, "[ba timing] %-22s %8.3f ms\n", name, (now_s() - (t0)) * 1e3); } while (0)

static inline double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int64_t arena_budget(ba_ctx *ctx, int64_t &cap)
{
    if (ctx->opt_arena_bytes > 0) {
        cap = ctx->opt_arena_bytes;
        return 0;
    }
    if (ctx->cached_arena_cap > 0) {
        cap = ctx->cached_arena_cap;
        return 0;
    }
    size_t free_b = 0, total_b = 0;
    cudaError_t e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return -1;
    cap = ctx->opt_arena_bytes;
    if (cap <= 0) {
        cap = (int64_t)((double)(free_b + ctx->arena.cap) * 0.6);
        const int64_t lim = 48LL << 30;
        if (cap > lim) cap = lim;
    }
    ctx->cached_arena_cap = cap;
    return 0;
}

template <typename Fn>
int cached_occupancy(ba_ctx *ctx, Fn fn, int threads, int &occ, size_t dyn_smem = 0)
{
    for (auto &e : ctx->occ_cache)
        if (e.first == (const void *)fn) {
            occ = e.second;
            return BA_OK;
        }
    occ = 1;
#ifndef BA_EMULATE
    if (dyn_smem > 0)  // beyond the 48 KB a kernel gets without opting in (static + dynamic)
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
#endif
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, dyn_smem));
    if (occ < 1) occ = 1;
    ctx->occ_cache.push_back({(const void *)fn, occ});
    return BA_OK;
}

// D2H of one wave's compact segments + per-item (count, base), then the host-side scatter of
// seg_start/seg_count by pair id.  `skip_overflow`: items whose count exceeds their slot
// capacity are left for the exact path and reported through `overflow`.
int gather_wave(ba_ctx *ctx, Batch &B, const Plan &plan, const Plan::Wave &w, int32_t *h_cnt, int64_t *h_base,
                int64_t wsegs, const int32_t *d_segs, const int32_t *h_raw_cnt, std::vector<int32_t> *overflow,
                bool release_token = false)
{
    cudaStream_t st = B.st;
    int32_t *h_segs = (int32_t *)pinned_get(ctx, sizeof(int32_t) * 5 * (size_t)(wsegs + 1));
    if (!h_segs) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    CK(cudaEventRecord(ctx->ev[2], st));
    if (wsegs > 0)
        CK(cudaMemcpyAsync(h_segs, d_segs, sizeof(int32_t) * 5 * (size_t)wsegs, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[3], st));
    if (release_token && B.token && B.token->owns_lock()) B.token->unlock();  // no kernel of this batch follows
    // the per-pair scatter only needs the counts (already on the host): do it while the segments travel
    const double t0 = now_s();
    for (int64_t k = 0; k < w.count; k++) {
        const int32_t p = plan.items[w.first + k].pair;
        if (overflow && (h_raw_cnt[k] < 0 || h_raw_cnt[k] > plan.slot_cap[w.first + k])) {
            overflow->push_back(p);
            continue;
        }
        B.out->seg_start[p] = B.seg_total + h_base[k];
        B.out->seg_count[p] = h_cnt[k];
    }
    BA_TSEC("gather scatter", t0);
    B.host_s += now_s() - t0;
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
    B.ms_d2h += ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    B.ms_fill += ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]));
    B.ms_trace += ms;
    B.wave_out.push_back({h_segs, wsegs});
    B.seg_total += wsegs;
    return BA_OK;
}

void plan_cache_drop(ba_ctx *ctx)
{
    ba_ctx::PlanCache &pc = ctx->pc;
    if (pc.plan_built) {
        pinned_put(ctx, pc.plan.items);
        pinned_put(ctx, pc.plan.slot_off);
        pinned_put(ctx, pc.plan.slot_cap);
    }
    pc.plan = Plan();
    pc.plan_built = false;
    pc.valid = false;
    pc.dev_items = pc.dev_slot_off = pc.dev_slot_cap = nullptr;
}

// ---- path A: 32-bit exact-code fill + two-pass walk (all kinds, any matrix) ---------------
int run_path32(ba_ctx *ctx, Batch &B, const int32_t *ids, int64_t nids)
{
    if (nids == 0) return BA_OK;
    const int kind = ctx->h_sc.kind, affine = ctx->h_sc.affine;
    cudaStream_t st = B.st;
    int rc;
    ctx->pc.dev_items = nullptr;  // this path reuses the work-list buffers
    int64_t arena_cap = 0;
    if (arena_budget(ctx, arena_cap) != 0) {
        ctx->err = "cudaMemGetInfo failed";
        return BA_ERR_CUDA;
    }
    Plan plan;
    const double tp0 = now_s();
    plan.items = (WorkItem *)pinned_get(ctx, sizeof(WorkItem) * (size_t)nids);
    plan.aux_off = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)nids);
    if (!plan.items || !plan.aux_off) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    build_plan(ids, nids, B.h_ref_len, B.h_qry_len, kind, affine, false, arena_cap, plan);
    B.host_s += now_s() - tp0;

    int64_t max_items = 0, max_code = 0, max_aux = 0;
    for (auto &w : plan.waves) {
        max_items = std::max(max_items, w.count);
        max_code = std::max(max_code, w.code_bytes);
        max_aux = std::max(max_aux, w.aux_ints);
    }
    if ((rc = ensure(ctx, ctx->items, sizeof(WorkItem) * (size_t)max_items)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->aux_off, sizeof(int64_t) * (size_t)(max_items + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->aux, sizeof(int32_t) * (size_t)(max_aux + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->seg_count, sizeof(int32_t) * (size_t)max_items)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->seg_base, sizeof(int64_t) * (size_t)(max_items + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->arena, (size_t)max_code + 256)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->counters, sizeof(int) * 128)) != BA_OK) return rc;
    size_t scan_bytes = 0;
    if ((rc = exclusive_scan(ctx, nullptr, scan_bytes, (const int32_t *)ctx->seg_count.p,
                             (int64_t *)ctx->seg_base.p, (int)max_items, st)) != BA_OK)
        return rc;
    if ((rc = ensure(ctx, ctx->scan_tmp, scan_bytes + 16)) != BA_OK) return rc;
    int32_t *h_cnt = (int32_t *)pinned_get(ctx, sizeof(int32_t) * (size_t)max_items);
    int64_t *h_base = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)(max_items + 1));
    if (!h_cnt || !h_base) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    WalkFn walk_count = pick_walk<false>(kind, affine);
    WalkFn walk_write = pick_walk<true>(kind, affine);

    for (auto &w : plan.waves) {
        WorkItem *d_items = (WorkItem *)ctx->items.p;
        if ((rc = upload_words(ctx, d_items, plan.items + w.first, sizeof(WorkItem) * (size_t)w.count, st)) != BA_OK)
            return rc;
        B.launches++;
        int64_t *d_aux_off = (int64_t *)ctx->aux_off.p;
        if (kind == KIND_FIT) {
            if ((rc = upload_words(ctx, d_aux_off, plan.aux_off + w.first, sizeof(int64_t) * (size_t)w.count, st)) !=
                BA_OK)
                return rc;
            B.launches++;
        }
        CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(int) * 128, st));

        // ---- fill
        CK(cudaEventRecord(ctx->ev[4], st));
        int li = 0;
        for (auto &L : w.launches) {
            FillFn fn = pick_fill(kind, affine, L.cols);
            int occ = 1;
            if ((rc = cached_occupancy(ctx, fn, BA_FILL_WARPS * 32, occ)) != BA_OK) return rc;
            int64_t blocks = (L.count + BA_FILL_WARPS - 1) / BA_FILL_WARPS;
            int64_t maxb = (int64_t)ctx->sm_count * occ;
            if (blocks > maxb) blocks = maxb;
            int64_t bnd_stride = 0;
            if (w.max_n_multipass > 0) {
                bnd_stride = (int64_t)(affine ? 3 : 1) * ((int64_t)w.max_n_multipass + 2);
                if ((rc = ensure(ctx, ctx->bnd, sizeof(int32_t) * (size_t)(bnd_stride * blocks * BA_FILL_WARPS))) !=
                    BA_OK)
                    return rc;
            }
            FillParams fp;
            fp.enc = B.d_enc;
            fp.stride = B.stride;
            fp.ref_off = B.d_ref_off;
            fp.ref_len = B.d_ref_len;
            fp.qry_off = B.d_qry_off;
            fp.qry_len = B.d_qry_len;
            fp.status = B.d_status;
            fp.items = d_items + (L.first - w.first);
            fp.n_items = (int)L.count;
            fp.counter = (int *)ctx->counters.p + (li++ % 64);
            fp.arena = (uint8_t *)ctx->arena.p;
            fp.bnd = (int32_t *)ctx->bnd.p;
            fp.bnd_stride = bnd_stride;
            fp.start = B.d_start;
            fp.aux = (int32_t *)ctx->aux.p;
            fp.aux_off = d_aux_off + (L.first - w.first);
            fp.sc = ctx->d_sc;
            BA_LAUNCH(fn, (unsigned)blocks, BA_FILL_WARPS * 32, st, fp);
            B.launches++;
            B.fill_launches++;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[5], st));
        B.code_bytes += w.code_bytes;

        // ---- walk: count, scan, write
        WalkParams wp;
        wp.enc = B.d_enc;
        wp.stride = B.stride;
        wp.ref_off = B.d_ref_off;
        wp.ref_len = B.d_ref_len;
        wp.qry_off = B.d_qry_off;
        wp.qry_len = B.d_qry_len;
        wp.status = B.d_status;
        wp.items = d_items;
        wp.n_items = (int)w.count;
        wp.arena = (const uint8_t *)ctx->arena.p;
        wp.start = B.d_start;
        wp.aux = (const int32_t *)ctx->aux.p;
        wp.aux_off = d_aux_off;
        wp.sc = ctx->d_sc;
        wp.seg_count = (int32_t *)ctx->seg_count.p;
        wp.seg_base = (const int64_t *)ctx->seg_base.p;
        wp.segs = nullptr;
        const unsigned wblocks = (unsigned)((w.count + 127) / 128);
        BA_LAUNCH(walk_count, wblocks, 128, st, wp);
        B.launches++;
        {
            size_t tb = ctx->scan_tmp.cap;
            if ((rc = exclusive_scan(ctx, ctx->scan_tmp.p, tb, (const int32_t *)ctx->seg_count.p,
                                     (int64_t *)ctx->seg_base.p, (int)w.count, st)) != BA_OK)
                return rc;
            B.launches += 2;
        }
        CK(cudaMemcpyAsync(h_cnt, ctx->seg_count.p, sizeof(int32_t) * (size_t)w.count, cudaMemcpyDeviceToHost,
                           st));
        CK(cudaMemcpyAsync(h_base, ctx->seg_base.p, sizeof(int64_t) * (size_t)w.count, cudaMemcpyDeviceToHost,
                           st));
        CK(cudaStreamSynchronize(st));
        const int64_t wsegs = h_base[w.count - 1] + h_cnt[w.count - 1];
        if ((rc = ensure(ctx, ctx->segs, sizeof(int32_t) * 5 * (size_t)(wsegs + 1))) != BA_OK) return rc;
        wp.segs = (int32_t *)ctx->segs.p;
        BA_LAUNCH(walk_write, wblocks, 128, st, wp);
        B.launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[6], st));
        if ((rc = gather_wave(ctx, B, plan, w, h_cnt, h_base, wsegs, (const int32_t *)ctx->segs.p, nullptr,
                              nullptr)) != BA_OK)
            return rc;
    }
    pinned_put(ctx, h_cnt);
    pinned_put(ctx, h_base);
    pinned_put(ctx, plan.items);
    pinned_put(ctx, plan.aux_off);
    return BA_OK;
}

// ---- path B: packed 16-bit DPX fill + checkpointed tile traceback (SWAffine, DNA-like) ------
}  // namespace
typedef void (*Fill16Fn)(Fill16Params);
// the kernel instances live in ba_fill16_inst.cu, one translation unit per (kind, linear) family
Fill16Fn ba_pick_fill16_0_0(int cols, bool multi, bool gsub, bool wavefront);
Fill16Fn ba_pick_fill16_1_0(int cols, bool multi, bool gsub, bool wavefront);
Fill16Fn ba_pick_fill16_2_0(int cols, bool multi, bool gsub, bool wavefront);
Fill16Fn ba_pick_fill16_0_1(int cols, bool multi, bool gsub, bool wavefront);
Fill16Fn ba_pick_fill16_1_1(int cols, bool multi, bool gsub, bool wavefront);
Fill16Fn ba_pick_fill16_2_1(int cols, bool multi, bool gsub, bool wavefront);
Fill16Fn ba_pick_fill16_0_2(int cols, bool multi, bool gsub, bool wavefront);
Fill16Fn ba_pick_fill16_1_2(int cols, bool multi, bool gsub, bool wavefront);
Fill16Fn ba_pick_fill16_2_2(int cols, bool multi, bool gsub, bool wavefront);
namespace {
bool g_sym_scoring(const DeviceScoring &s)
{
    // affine with one gap-extension value for both sequences and go + e <= 0: the SYM kernels apply
    const int er = s.la[1 * s.let], eq = s.la[1];
    return s.affine && er == eq && s.gap_open + er <= 0;
}
Fill16Fn pick_fill16(int kind, bool lin, int cols, bool multi, bool gsub, bool wavefront, bool sym = false)
{
    if (!lin && sym) {
        if (kind == KIND_SW) return ba_pick_fill16_0_2(cols, multi, gsub, wavefront);
        if (kind == KIND_NW) return ba_pick_fill16_1_2(cols, multi, gsub, wavefront);
        return ba_pick_fill16_2_2(cols, multi, gsub, wavefront);
    }
    if (lin) {
        if (kind == KIND_SW) return ba_pick_fill16_0_1(cols, multi, gsub, wavefront);
        if (kind == KIND_NW) return ba_pick_fill16_1_1(cols, multi, gsub, wavefront);
        return ba_pick_fill16_2_1(cols, multi, gsub, wavefront);
    }
    if (kind == KIND_SW) return ba_pick_fill16_0_0(cols, multi, gsub, wavefront);
    if (kind == KIND_NW) return ba_pick_fill16_1_0(cols, multi, gsub, wavefront);
    return ba_pick_fill16_2_0(cols, multi, gsub, wavefront);
}
typedef void (*Trace16Fn)(Trace16Params);
template <int NT>
Trace16Fn pick_trace16t(int kind, bool lin)
{
    if (lin) {
        if (kind == KIND_SW) return ba_trace16w_kernel<KIND_SW, true, NT>;
        if (kind == KIND_NW) return ba_trace16w_kernel<KIND_NW, true, NT>;
        return ba_trace16w_kernel<KIND_FIT, true, NT>;
    }
    if (kind == KIND_SW) return ba_trace16w_kernel<KIND_SW, false, NT>;
    if (kind == KIND_NW) return ba_trace16w_kernel<KIND_NW, false, NT>;
    return ba_trace16w_kernel<KIND_FIT, false, NT>;
}
Trace16Fn pick_trace16p(int kind, bool lin)
{
    if (lin) {
        if (kind == KIND_SW) return ba_trace16p_kernel<KIND_SW, true>;
        if (kind == KIND_NW) return ba_trace16p_kernel<KIND_NW, true>;
        return ba_trace16p_kernel<KIND_FIT, true>;
    }
    if (kind == KIND_SW) return ba_trace16p_kernel<KIND_SW, false>;
    if (kind == KIND_NW) return ba_trace16p_kernel<KIND_NW, false>;
    return ba_trace16p_kernel<KIND_FIT, false>;
}
Trace16Fn pick_trace16(int kind, bool lin)
{
    if (lin) {
        if (kind == KIND_SW) return ba_trace16_kernel<KIND_SW, true>;
        if (kind == KIND_NW) return ba_trace16_kernel<KIND_NW, true>;
        return ba_trace16_kernel<KIND_FIT, true>;
    }
    if (kind == KIND_SW) return ba_trace16_kernel<KIND_SW, false>;
    if (kind == KIND_NW) return ba_trace16_kernel<KIND_NW, false>;
    return ba_trace16_kernel<KIND_FIT, false>;
}

// Scoring-level eligibility of the packed 16-bit path (ba_fill16.cuh):
//   uniform gap entries over the non-gap letters (C1); affine SW additionally needs
//   non-positive gaps and go + gap + max(sub) < 0 (C2), linear SW strictly negative gaps (a cell
//   holding the table maximum then always came through the diagonal).  `gsub`: the alphabet is not the
//   5-letter gapped nucleotide one, so substitution scores come from shared-memory tables
//   instead of PRMT profiles.  maxsub / maxabs feed the per-pair 16-bit range checks.
struct Fast16 {
    bool ok = false, gsub = false;
    int maxsub = 0, maxabs = 0;
};
Fast16 fast16_scoring(const DeviceScoring &s)
{
    Fast16 f;
    if (s.let < 2 || s.let > BA_MAX_LET) return f;
    const int64_t gap_open = s.affine ? s.gap_open : 0;
    f.gsub = (s.let != 5);
    if (f.gsub && s.let > BA_MAX_LET - 1) return f;  // row `let` of the score tables is the virtual letter
    const int er = s.la[1 * s.let], eq = s.la[1];
    int maxsub = -1000000, maxabs = 0;
    for (int a = 1; a < s.let; a++) {
        if (s.la[a * s.let] != er || s.la[a] != eq) return f;
        for (int b = 1; b < s.let; b++) {
            const int v = s.la[a * s.let + b];
            if (v < -127 || v > 127) return f;
            maxsub = std::max(maxsub, v);
            maxabs = std::max(maxabs, v < 0 ? -v : v);
        }
    }
    if (er < -127 || er > 127 || eq < -127 || eq > 127) return f;
    if (gap_open < -4000 || gap_open > 4000) return f;
    maxabs = std::max(maxabs, std::max(er < 0 ? -er : er, eq < 0 ? -eq : eq));
    if (s.kind == KIND_SW && s.affine) {
        if (er > 0 || eq > 0 || maxsub <= 0) return f;
        if (gap_open + er + maxsub >= 0 || gap_open + eq + maxsub >= 0) return f;
    }
    if (s.kind == KIND_SW && !s.affine && (er >= 0 || eq >= 0)) return f;
    f.maxsub = maxsub;
    f.maxabs = maxabs;
    f.ok = true;
    return f;
}
// Per-pair 16-bit range check.
bool fast16_pair_ok(const DeviceScoring &s, const Fast16 &f, int n, int m)
{
    const int64_t lo = std::min(n, m), hi = std::max(n, m);
    if (lo <= 0 || n >= (1 << 19)) return false;
    if (lo * std::max(f.maxsub, 0) >= 30000) return false;
    if (s.kind != KIND_SW) {
        const int64_t go = s.affine ? s.gap_open : 0;
        const int64_t ago = go < 0 ? -go : go;
        if ((hi + 3) * f.maxabs + 2 * ago > 14000) return false;
    }
    return true;
}

int run_path16(ba_ctx *ctx, Batch &B, const int32_t *ids, int64_t nids, bool gsub, std::vector<int32_t> &overflow,
               bool use_cache)
{
    const int kind = ctx->h_sc.kind;
    const bool lin = !ctx->h_sc.affine;
    const int flag_mask = gsub ? 2 : 1;
    if (nids == 0) return BA_OK;
    cudaStream_t st = B.st;
    int rc;
    int64_t arena_cap = 0;
    if (arena_budget(ctx, arena_cap) != 0) {
        ctx->err = "cudaMemGetInfo failed";
        return BA_ERR_CUDA;
    }
    Plan local_plan;
    ba_ctx::PlanCache &pc = ctx->pc;
    const double tp0 = now_s();
    if (use_cache && pc.plan_built && (pc.arena_cap != arena_cap || pc.slot_cap_opt != ctx->opt_slot_cap)) {
        pinned_put(ctx, pc.plan.items);
        pinned_put(ctx, pc.plan.slot_off);
        pinned_put(ctx, pc.plan.slot_cap);
        pc.plan = Plan();
        pc.plan_built = false;
        pc.dev_items = nullptr;
    }
    Plan &plan = use_cache ? pc.plan : local_plan;
    if (!(use_cache && pc.plan_built)) {
        plan.items = (WorkItem *)pinned_get(ctx, sizeof(WorkItem) * (size_t)nids);
        plan.slot_off = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)nids);
        plan.slot_cap = (int32_t *)pinned_get(ctx, sizeof(int32_t) * (size_t)nids);
        if (!plan.items || !plan.slot_off || !plan.slot_cap) {
            ctx->err = "pinned host allocation failed";
            return BA_ERR_NOMEM;
        }
        const double tp1 = now_s();
        build_plan(ids, nids, B.h_ref_len, B.h_qry_len, kind, lin ? 0 : 1, true, arena_cap, plan);
        BA_TSEC("plan16 pinned_get", tp0);
        BA_TSEC("plan16 build", tp1);
        if (ctx->opt_slot_cap > 0) {  // test hook: force the overflow -> exact-path fallback
            for (auto &w : plan.waves) {
                w.slots = 0;
                for (int64_t k = 0; k < w.count; k++) {
                    plan.slot_off[w.first + k] = w.slots;
                    plan.slot_cap[w.first + k] = (int32_t)ctx->opt_slot_cap;
                    w.slots += ctx->opt_slot_cap;
                }
            }
        }
        if (use_cache) {
            pc.plan_built = true;
            pc.arena_cap = arena_cap;
            pc.slot_cap_opt = ctx->opt_slot_cap;
            pc.dev_items = nullptr;
        }
    }
    B.host_s += now_s() - tp0;
    int64_t max_items = 0, max_code = 0, max_slots = 0;
    for (auto &w : plan.waves) {
        max_items = std::max(max_items, w.count);
        max_code = std::max(max_code, w.code_bytes);
        max_slots = std::max(max_slots, w.slots);
    }
    if ((rc = ensure(ctx, ctx->items, sizeof(WorkItem) * (size_t)max_items)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->slot_off, sizeof(int64_t) * (size_t)(max_items + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->slot_cap, sizeof(int32_t) * (size_t)(max_items + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->slots, sizeof(int32_t) * 5 * (size_t)(max_slots + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->seg_count, sizeof(int32_t) * (size_t)max_items)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->seg_clip, sizeof(int32_t) * (size_t)max_items)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->seg_base, sizeof(int64_t) * (size_t)(max_items + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->arena, (size_t)max_code + 256)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->counters, sizeof(int) * 128)) != BA_OK) return rc;
    size_t scan_bytes = 0;
    if ((rc = exclusive_scan(ctx, nullptr, scan_bytes, (const int32_t *)ctx->seg_clip.p,
                             (int64_t *)ctx->seg_base.p, (int)max_items, st)) != BA_OK)
        return rc;
    if ((rc = ensure(ctx, ctx->scan_tmp, scan_bytes + 16)) != BA_OK) return rc;
    int32_t *h_long_err = (int32_t *)pinned_get(ctx, 64);
    if (!h_long_err) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    memset(h_long_err, 0, 64);
    std::vector<WorkItem *> long_tasks;
    int32_t *h_cnt = (int32_t *)pinned_get(ctx, sizeof(int32_t) * (size_t)max_items);
    int32_t *h_raw = (int32_t *)pinned_get(ctx, sizeof(int32_t) * (size_t)max_items);
    int64_t *h_base = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)(max_items + 1));
    if (!h_cnt || !h_raw || !h_base) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }

    for (auto &w : plan.waves) {
        WorkItem *d_items = (WorkItem *)ctx->items.p;
        const bool dev_ok = use_cache && plan.waves.size() == 1 && pc.dev_items == ctx->items.p &&
                            pc.dev_slot_off == ctx->slot_off.p && pc.dev_slot_cap == ctx->slot_cap.p;
        if (!dev_ok) {
            if ((rc = upload_words(ctx, d_items, plan.items + w.first, sizeof(WorkItem) * (size_t)w.count, st)) !=
                BA_OK)
                return rc;
            if ((rc = upload_words(ctx, ctx->slot_off.p, plan.slot_off + w.first, sizeof(int64_t) * (size_t)w.count,
                                   st)) != BA_OK)
                return rc;
            if ((rc = upload_words(ctx, ctx->slot_cap.p, plan.slot_cap + w.first, sizeof(int32_t) * (size_t)w.count,
                                   st)) != BA_OK)
                return rc;
            B.launches += 3;
            if (use_cache && plan.waves.size() == 1) {
                pc.dev_items = ctx->items.p;
                pc.dev_slot_off = ctx->slot_off.p;
                pc.dev_slot_cap = ctx->slot_cap.p;
            } else {
                pc.dev_items = nullptr;
            }
        }
        CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(int) * 128, st));

        // ---- fill (values + checkpoints)
        CK(cudaEventRecord(ctx->ev[4], st));
        int li = 0;
        for (auto &L : w.launches) {
            const bool multi = w.max_n_multipass > 0;
            Fill16Params fp;
            fp.enc = B.d_enc;
            fp.stride = B.stride;
            fp.ref_off = B.d_ref_off;
            fp.ref_len = B.d_ref_len;
            fp.qry_off = B.d_qry_off;
            fp.qry_len = B.d_qry_len;
            fp.status = B.d_status;
            fp.acgt_only = (const uint8_t *)ctx->flags.p;
            fp.flag_mask = flag_mask;
            fp.counter = (int *)ctx->counters.p + (li++ % 64);
            fp.arena = (uint8_t *)ctx->arena.p;
            fp.sc = ctx->d_sc;
            fp.long_couples = 0;
            fp.long_err = nullptr;

            // ---- LONG: too few pairs to fill the GPU pair by pair, and many passes each: spread the
            // passes of every pair over warps (wavefront).  Items of a launch are sorted by cells.
            const int W16 = 32 * L.cols;
            int maxp = 0, max_n = 0;
            if (multi && L.cols >= 7 && ctx->opt_long_mode != 0) {
                for (int64_t k = 0; k < L.count; k++) {
                    const int32_t p = plan.items[L.first + k].pair;
                    maxp = std::max(maxp, (B.h_qry_len[p] + W16 - 1) / W16);
                    max_n = std::max(max_n, B.h_ref_len[p]);
                }
            }
            const bool sym = !lin && ctx->opt_sym != 0 && g_sym_scoring(ctx->h_sc);
            Fill16Fn fn = pick_fill16(kind, lin, L.cols, multi, gsub, false, sym);
            // GSUB kernels keep their score tables in dynamic shared memory (the size depends on the
            // matrix only, so the per-kernel cache entry stays valid until the scoring changes)
            const size_t dyn = gsub ? (size_t)BA16_WARPS * ba16_gsub_bytes_per_warp(ctx->h_sc.let) : 0;
            int occ = 1;
            if ((rc = cached_occupancy(ctx, fn, BA16_WARPS * 32, occ, dyn)) != BA_OK) return rc;
            int64_t maxb = (int64_t)ctx->sm_count * occ;
            const int64_t pair_warps = (L.count + 3) / 4;
            const bool use_long = maxp >= 4 && (ctx->opt_long_mode == 2 || pair_warps < maxb * BA16_WARPS) &&
                                  (int64_t)maxp * (L.count + 4) < (1 << 22);
            if (use_long) {
                fn = pick_fill16(kind, lin, L.cols, true, gsub, true, sym);
                if ((rc = cached_occupancy(ctx, fn, BA16_WARPS * 32, occ, dyn)) != BA_OK) return rc;
                maxb = (int64_t)ctx->sm_count * occ;
                const int64_t ncp = ((L.count + 3) / 4) * 2;         // couples per pass, even
                const int64_t ntasks = (int64_t)maxp * 2 * ncp;
                WorkItem *h_tasks = (WorkItem *)pinned_get(ctx, sizeof(WorkItem) * (size_t)ntasks);
                if (!h_tasks) {
                    ctx->err = "pinned host allocation failed";
                    return BA_ERR_NOMEM;
                }
                for (int ps = 0; ps < maxp; ps++)
                    for (int64_t k = 0; k < 2 * ncp; k++) {
                        WorkItem t;
                        t.pair = -1;
                        t.cols = L.cols;
                        t.code_off = 0;
                        if (k < L.count) {
                            const WorkItem &it = plan.items[L.first + k];
                            if (ps < (B.h_qry_len[it.pair] + W16 - 1) / W16) t = it;
                        }
                        h_tasks[(int64_t)ps * 2 * ncp + k] = t;
                    }
                const int64_t bnd_stride = 4 * ((int64_t)max_n + 2);
                if ((rc = ensure(ctx, ctx->tasks, sizeof(WorkItem) * (size_t)ntasks)) != BA_OK) return rc;
                if ((rc = ensure(ctx, ctx->prog, 64)) != BA_OK) return rc;
                if ((rc = ensure(ctx, ctx->bnd, sizeof(int32_t) * (size_t)(bnd_stride * ncp))) != BA_OK) return rc;
                if ((rc = upload_words(ctx, ctx->tasks.p, h_tasks, sizeof(WorkItem) * (size_t)ntasks, st)) != BA_OK)
                    return rc;
                B.launches++;
                CK(cudaMemsetAsync(ctx->prog.p, 0, 64, st));
                // row tags start out invalid
                CK(cudaMemsetAsync(ctx->bnd.p, 0, sizeof(int32_t) * (size_t)(bnd_stride * ncp), st));
                fp.items = (const WorkItem *)ctx->tasks.p;
                fp.n_items = (int)ntasks;
                fp.bnd = (int32_t *)ctx->bnd.p;
                fp.bnd_stride = bnd_stride;
                fp.long_couples = (int)ncp;
                fp.long_err = (int *)ctx->prog.p;
                int64_t blocks = (ntasks + 4 * BA16_WARPS - 1) / (4 * BA16_WARPS);
                if (blocks > maxb) blocks = maxb;
                BA_LAUNCH_SMEM(fn, (unsigned)blocks, BA16_WARPS * 32, dyn, st, fp);
                CK(cudaMemcpyAsync(h_long_err + (long_tasks.size() & 15), ctx->prog.p, sizeof(int), cudaMemcpyDeviceToHost,
                                   st));
                long_tasks.push_back(h_tasks);
                B.out->wavefront_launches++;
                B.launches++;
                B.fill_launches++;
                continue;
            }
            int64_t blocks = (L.count + 2 * BA16_GROUPS - 1) / (2 * BA16_GROUPS);  // two pairs per group
            if (blocks > maxb) blocks = maxb;
            int64_t bnd_stride = 0;
            if (multi) {
                bnd_stride = 4 * ((int64_t)w.max_n_multipass + 2);
                if ((rc = ensure(ctx, ctx->bnd, sizeof(int32_t) * (size_t)(bnd_stride * blocks * BA16_GROUPS))) != BA_OK)
                    return rc;
            }
            fp.items = d_items + (L.first - w.first);
            fp.n_items = (int)L.count;
            fp.bnd = (int32_t *)ctx->bnd.p;
            fp.bnd_stride = bnd_stride;
            BA_LAUNCH_SMEM(fn, (unsigned)blocks, BA16_WARPS * 32, dyn, st, fp);
            B.launches++;
            B.fill_launches++;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[5], st));
        B.code_bytes += w.code_bytes;
        BA_TSEC("P16 fill launched", tp0);

        // ---- tile traceback into slots, clip, scan, compact
        Trace16Params tp;
        tp.enc = B.d_enc;
        tp.stride = B.stride;
        tp.ref_off = B.d_ref_off;
        tp.ref_len = B.d_ref_len;
        tp.qry_off = B.d_qry_off;
        tp.qry_len = B.d_qry_len;
        tp.status = B.d_status;
        tp.acgt_only = (const uint8_t *)ctx->flags.p;
        tp.flag_mask = flag_mask;
        tp.items = d_items;
        tp.n_items = (int)w.count;
        tp.arena = (const uint8_t *)ctx->arena.p;
        tp.sc = ctx->d_sc;
        tp.seg_count = (int32_t *)ctx->seg_count.p;
        tp.slot_off = (const int64_t *)ctx->slot_off.p;
        tp.slot_cap = (const int32_t *)ctx->slot_cap.p;
        tp.slots = (int32_t *)ctx->slots.p;
        tp.counter = (int *)ctx->counters.p + 64;  // zeroed with the fill counters at the start of the wave
        // one thread per pair when the batch fills the GPU that way, else one warp per pair
        // (window of speculatively rebuilt tiles, see ba_trace16w_kernel), a whole CTA per pair when
        // even that leaves SMs idle; a mixed batch gives warps to its longest pairs (the head of
        // the size-sorted work list) as far as thread slots are free
        const bool cta_trace = ctx->opt_trace_mode == 3 ||
                               (ctx->opt_trace_mode == 1 && w.count <= 2 * (int64_t)ctx->sm_count);
        const bool warp_trace = ctx->opt_trace_mode == 2 ||
                                (ctx->opt_trace_mode == 1 &&
                                 w.count * 32 <= (int64_t)ctx->sm_count * BA_TRACE_MINB * BA_TRACE_THREADS * 2);
        if (cta_trace) {
            BA_LAUNCH(pick_trace16t<BA_TRACE_THREADS>(kind, lin), (unsigned)w.count, BA_TRACE_THREADS, st, tp);
            B.out->warp_trace_launches++;
        } else if (warp_trace) {
            const unsigned tb = (unsigned)((w.count + BA_TRACE_THREADS / 32 - 1) / (BA_TRACE_THREADS / 32));
            BA_LAUNCH(pick_trace16t<32>(kind, lin), tb, BA_TRACE_THREADS, st, tp);
            B.out->warp_trace_launches++;
        } else {
            const int64_t slots_free = (int64_t)ctx->sm_count * BA_TRACE_MINB * BA_TRACE_THREADS - w.count;
            auto path_len = [&](int64_t k) {
                const WorkItem &x = plan.items[w.first + k];
                return (int64_t)B.h_ref_len[x.pair] + B.h_qry_len[x.pair];
            };
            int64_t head = 0;
            if (ctx->opt_trace_mode == 1) {
                if (path_len(0) >= 3 * path_len(w.count / 2)) {
                    // strongly mixed lengths: with a thread per pair the few longest walks would run
                    // alone at the end; warps shorten exactly those
                    head = w.count;
                } else if (slots_free > 0) {
                    head = std::min(w.count, slots_free / 31);
                    if (head < w.count && path_len(0) < 2 * path_len(head)) head = 0;  // the head is not longer
                }
            }
            if (ctx->opt_trace_mode == 4) head = (w.count + 3) / 4;  // test hook: always split
            if (head > 0) {
                Trace16Params th = tp;
                th.n_items = (int)head;
                const unsigned tb = (unsigned)((head + BA_TRACE_THREADS / 32 - 1) / (BA_TRACE_THREADS / 32));
                BA_LAUNCH(pick_trace16t<32>(kind, lin), tb, BA_TRACE_THREADS, st, th);
                B.out->warp_trace_launches++;
                B.launches++;
            }
            if (head < w.count) {
                Trace16Params tt = tp;
                tt.items = tp.items + head;
                tt.n_items = (int)(w.count - head);
                tt.seg_count = tp.seg_count + head;
                tt.slot_off = tp.slot_off + head;
                tt.slot_cap = tp.slot_cap + head;
                int64_t tblocks = (w.count - head + BA_TRACE_THREADS - 1) / BA_TRACE_THREADS;
                if (ctx->opt_trace_mode != 5) {
                    // static assignment: thread k of the grid walks pair k
                    BA_LAUNCH(pick_trace16(kind, lin), (unsigned)tblocks, BA_TRACE_THREADS, st, tt);
                } else {
                    // persistent grid, pairs claimed from a counter (largest first).  Measured on B200:
                    // +5 % on configs 1 and 3, -9 % on config 2 against the static kernel, so it is
                    // opt-in (trace_mode 5) until the state machine's overhead is trimmed
                    tblocks = std::min<int64_t>(tblocks, (int64_t)ctx->sm_count * BA_TRACE_MINB);
                    BA_LAUNCH(pick_trace16p(kind, lin), (unsigned)tblocks, BA_TRACE_THREADS, st, tt);
                }
            }
        }
        const unsigned cblocks = (unsigned)((w.count + 255) / 256);
        BA_LAUNCH(ba_clip_counts_kernel, cblocks, 256, st, (const int32_t *)ctx->seg_count.p,
                  (const int32_t *)ctx->slot_cap.p, (int32_t *)ctx->seg_clip.p, (int)w.count);
        B.launches += 2;
        {
            size_t tb = ctx->scan_tmp.cap;
            if ((rc = exclusive_scan(ctx, ctx->scan_tmp.p, tb, (const int32_t *)ctx->seg_clip.p,
                                     (int64_t *)ctx->seg_base.p, (int)w.count, st)) != BA_OK)
                return rc;
            B.launches += 2;
        }
        CK(cudaMemcpyAsync(h_cnt, ctx->seg_clip.p, sizeof(int32_t) * (size_t)w.count, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_raw, ctx->seg_count.p, sizeof(int32_t) * (size_t)w.count, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_base, ctx->seg_base.p, sizeof(int64_t) * (size_t)w.count, cudaMemcpyDeviceToHost,
                           st));
        BA_TSEC("P16 trace launched", tp0);
        CK(cudaStreamSynchronize(st));
        BA_TSEC("P16 counts synced", tp0);
        for (WorkItem *t : long_tasks) pinned_put(ctx, t);
        long_tasks.clear();
        for (int k = 0; k < 16; k++)
            if (h_long_err[k] != 0) {
                ctx->err = "fill16 wavefront: a pass waited too long for its left neighbour";
                return BA_ERR_CUDA;
            }
        const int64_t wsegs = h_base[w.count - 1] + h_cnt[w.count - 1];
        if ((rc = ensure(ctx, ctx->segs, sizeof(int32_t) * 5 * (size_t)(wsegs + 1))) != BA_OK) return rc;
        BA_LAUNCH(ba_compact_kernel, cblocks, 256, st, (const int32_t *)ctx->slots.p,
                  (const int64_t *)ctx->slot_off.p, (const int32_t *)ctx->slot_cap.p,
                  (const int32_t *)ctx->seg_count.p, (const int64_t *)ctx->seg_base.p, (int32_t *)ctx->segs.p,
                  (int)w.count);
        B.launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[6], st));
        // last wave, nothing queued for the exact path and no pair overflowed its slots or failed the
        // letter-class check: the kernel phase of this batch ends here
        bool last_kernels = B.token && &w == &plan.waves.back() && !B.more_kernels_later && overflow.empty();
        if (last_kernels)
            for (int64_t k = 0; k < w.count && last_kernels; k++)
                last_kernels = h_raw[k] >= 0 && h_raw[k] <= plan.slot_cap[w.first + k];
        if ((rc = gather_wave(ctx, B, plan, w, h_cnt, h_base, wsegs, (const int32_t *)ctx->segs.p, h_raw,
                              &overflow, last_kernels)) != BA_OK)
            return rc;
        BA_TSEC("P16 wave gathered", tp0);
    }
    pinned_put(ctx, h_long_err);
    pinned_put(ctx, h_cnt);
    pinned_put(ctx, h_raw);
    pinned_put(ctx, h_base);
    if (!use_cache) {
        pinned_put(ctx, plan.items);
        pinned_put(ctx, plan.slot_off);
        pinned_put(ctx, plan.slot_cap);
    }
    return BA_OK;
}

int run_device(ba_ctx *ctx, const uint8_t *d_letters, int64_t letters_bytes, int stride,
               const int64_t *d_ref_off, const int32_t *d_ref_len, const int64_t *d_qry_off,
               const int32_t *d_qry_len, const int32_t *h_ref_len, const int32_t *h_qry_len,
               int64_t n, cudaStream_t st, ba_result *out, std::unique_lock<std::mutex> *token = nullptr)
{
    const int kind = ctx->h_sc.kind, affine = ctx->h_sc.affine;
    const double t_call = now_s();

    // ---- result arrays (pinned, pooled)
    out->n = n;
    out->status = (int32_t *)pinned_get(ctx, sizeof(int32_t) * (size_t)n);
    out->err_pos = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)n);
    out->seg_start = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)n);
    out->seg_count = (int32_t *)pinned_get(ctx, sizeof(int32_t) * (size_t)n);
    out->segs = nullptr;
    out->n_segs = 0;
    if (!out->status || !out->err_pos || !out->seg_start || !out->seg_count) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    if (n == 0) return BA_OK;

    CK(cudaEventRecord(ctx->ev[1], st));

    Batch B;
    B.stride = stride;
    B.d_ref_off = d_ref_off;
    B.d_qry_off = d_qry_off;
    B.d_ref_len = d_ref_len;
    B.d_qry_len = d_qry_len;
    B.h_ref_len = h_ref_len;
    B.h_qry_len = h_qry_len;
    B.n = n;
    B.st = st;
    B.out = out;
    B.token = token;

    // ---- encode + validate
    int rc;
    if ((rc = ensure(ctx, ctx->enc, (size_t)letters_bytes + 16)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->status, sizeof(int32_t) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->err_pos, sizeof(int64_t) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->start, sizeof(PairStart) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->flags, (size_t)n + 16)) != BA_OK) return rc;
    uint8_t *d_enc = (uint8_t *)ctx->enc.p;
    B.d_enc = d_enc;
    B.d_status = (int32_t *)ctx->status.p;
    B.d_start = (PairStart *)ctx->start.p;
    int64_t *d_err_pos = (int64_t *)ctx->err_pos.p;
    const Fast16 f16 = fast16_scoring(ctx->h_sc);
    const bool fast16 = ctx->opt_force_path != 1 && f16.ok;
    if (letters_bytes > 0) {
        int64_t blocks = (letters_bytes / 16 + 255) / 256;
        if (blocks < 1) blocks = 1;
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        BA_LAUNCH(ba_encode_kernel, (unsigned)blocks, 256, st, d_letters, d_enc, letters_bytes, ctx->d_sc);
        B.launches++;
    }
    {
        int64_t blocks = (n + 7) / 8;
        if (blocks > ctx->sm_count * 32) blocks = ctx->sm_count * 32;
        BA_LAUNCH(ba_validate_kernel, (unsigned)blocks, 256, st, d_enc, stride, d_ref_off, d_ref_len, d_qry_off,
                  d_qry_len, n, kind, affine, B.d_status, d_err_pos, fast16 ? (uint8_t *)ctx->flags.p : nullptr);
        B.launches++;
    }
    CK(cudaGetLastError());
    BA_TSEC("T result+enc launch", t_call);

    // ---- split the batch between the two paths (cached for batches of the same shape)
    std::vector<int32_t> slow;
    if (fast16) {
        // size eligibility is known on the host; letter eligibility (ACGT only) is decided on the
        // device: ineligible pairs come back from the traceback marked for the exact path
        const double ts0 = now_s();
        ba_ctx::PlanCache &pc = ctx->pc;
        const bool hit = ctx->opt_plan_cache && pc.valid && pc.n == n && pc.epoch == ctx->scoring_epoch &&
                         memcmp(pc.ref_len.data(), h_ref_len, sizeof(int32_t) * (size_t)n) == 0 &&
                         memcmp(pc.qry_len.data(), h_qry_len, sizeof(int32_t) * (size_t)n) == 0;
        if (!hit) {
            plan_cache_drop(ctx);
            pc.n = n;
            pc.epoch = ctx->scoring_epoch;
            pc.ref_len.assign(h_ref_len, h_ref_len + n);
            pc.qry_len.assign(h_qry_len, h_qry_len + n);
            pc.fast.clear();
            pc.slow0.clear();
            pc.fast.reserve((size_t)n);
            pc.cells = 0;
            for (int64_t p = 0; p < n; p++) {
                pc.cells += (int64_t)h_ref_len[p] * (int64_t)h_qry_len[p];
                const bool ok = fast16_pair_ok(ctx->h_sc, f16, h_ref_len[p], h_qry_len[p]);
                (ok ? pc.fast : pc.slow0).push_back((int32_t)p);
            }
            pc.valid = true;
        }
        out->cells = pc.cells;
        slow = pc.slow0;
        BA_TSEC("partition", ts0);
        B.host_s += now_s() - ts0;
        const size_t nslow0 = slow.size();
        B.more_kernels_later = nslow0 > 0;
        if ((rc = run_path16(ctx, B, pc.fast.data(), (int64_t)pc.fast.size(), f16.gsub, slow,
                             ctx->opt_plan_cache != 0)) != BA_OK)
            return rc;
        out->pairs_fast16 = (int64_t)pc.fast.size() - (int64_t)(slow.size() - nslow0);
        if ((rc = run_path32(ctx, B, slow.data(), (int64_t)slow.size())) != BA_OK) return rc;
    } else {
        int64_t cells = 0;
        for (int64_t p = 0; p < n; p++) cells += (int64_t)h_ref_len[p] * (int64_t)h_qry_len[p];
        out->cells = cells;
        if ((rc = run_path32(ctx, B, nullptr, n)) != BA_OK) return rc;
    }

    BA_TSEC("T paths done", t_call);
    // ---- gather
    CK(cudaEventRecord(ctx->ev[2], st));
    CK(cudaMemcpyAsync(out->status, B.d_status, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->err_pos, d_err_pos, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[3], st));
    CK(cudaStreamSynchronize(st));
    {
        float ms = 0.f, all = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        B.ms_d2h += ms;
        CK(cudaEventElapsedTime(&all, ctx->ev[1], ctx->ev[3]));
        out->ms_d2h = B.ms_d2h;
        out->ms_kernels = all - B.ms_d2h;
    }
    if (B.wave_out.size() == 1) {
        out->segs = B.wave_out[0].segs;
    } else {
        out->segs = (int32_t *)pinned_get(ctx, sizeof(int32_t) * 5 * (size_t)(B.seg_total + 1));
        if (!out->segs) {
            ctx->err = "pinned host allocation failed";
            return BA_ERR_NOMEM;
        }
        int64_t o = 0;
        for (auto &wo : B.wave_out) {
            memcpy(out->segs + 5 * o, wo.segs, sizeof(int32_t) * 5 * (size_t)wo.nsegs);
            o += wo.nsegs;
            pinned_put(ctx, wo.segs);
        }
    }
    for (int64_t p = 0; p < n; p++)
        if (out->status[p] != 0) {
            out->seg_start[p] = 0;
            out->seg_count[p] = 0;
        }
    out->n_segs = B.seg_total;
    out->gpu_launches = B.launches;
    out->fill_launches = B.fill_launches;
    out->code_bytes = B.code_bytes;
    out->ms_fill = B.ms_fill;
    out->ms_trace = B.ms_trace;
    out->ms_host = (float)(B.host_s * 1e3);
    BA_TSEC("T call total", t_call);
    return BA_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------
extern "C" {

int ba_abi_version(void) { return BA_ABI_VERSION; }

const char *ba_build_info(void)
{
    return "libbiogoalign sm_100a (fill16 DPX + checkpointed trace, fill32 exact-code kernels; nvcc " __DATE__ ")";
}

int ba_create(int device, ba_ctx **out)
{
    if (!out) return BA_ERR_ARG;
    *out = nullptr;
    ba_ctx *ctx = new ba_ctx();
    ctx->device = device;
    auto fail = [&](cudaError_t e, const char *what) {
        // keep the ctx alive so the caller can read ba_last_error(); it holds no device state
        char b[256];
        snprintf(b, sizeof b, "%s: %s", what, cudaGetErrorString(e));
        ctx->err = b;
        *out = ctx;
        return BA_ERR_CUDA;
    };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) {
        ctx->err = "libbiogoalign is built for sm_100a (B200) only";
        *out = ctx;
        return BA_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(e, "cudaStreamCreate");
    for (int k = 0; k < 8; k++) {
        e = cudaEventCreate(&ctx->ev[k]);
        if (e != cudaSuccess) return fail(e, "cudaEventCreate");
    }
    e = cudaMalloc((void **)&ctx->d_sc, sizeof(DeviceScoring));
    if (e != cudaSuccess) return fail(e, "cudaMalloc");
    *out = ctx;
    return BA_OK;
}

void ba_destroy(ba_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    DevBuf *bufs[] = {&ctx->letters, &ctx->enc,  &ctx->ref_off,   &ctx->ref_len,  &ctx->qry_off, &ctx->qry_len,
                      &ctx->status,  &ctx->err_pos, &ctx->start,  &ctx->items,    &ctx->aux_off, &ctx->aux,
                      &ctx->bnd,     &ctx->seg_count, &ctx->seg_base, &ctx->segs, &ctx->arena,   &ctx->scan_tmp,
                      &ctx->counters, &ctx->flags,  &ctx->slot_off,  &ctx->slot_cap, &ctx->slots, &ctx->seg_clip,
                      &ctx->tasks,   &ctx->prog,     &ctx->fmt_len, &ctx->fmt_off, &ctx->fmt_rows,
                      &ctx->fmt_segs, &ctx->fmt_seg_start, &ctx->fmt_seg_count};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (auto &b : ctx->pinned) cudaFreeHost(b.p);
    if (ctx->d_sc) cudaFree(ctx->d_sc);
    for (int k = 0; k < 8; k++)
        if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *ba_last_error(const ba_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ba_set_option(ba_ctx *ctx, const char *key, int64_t value)
{
    if (!ctx || !key) return BA_ERR_ARG;
    if (!strcmp(key, "arena_bytes")) {
        ctx->opt_arena_bytes = value;
        return BA_OK;
    }
    if (!strcmp(key, "slot_cap")) {
        ctx->opt_slot_cap = value;
        return BA_OK;
    }
    if (!strcmp(key, "long_mode")) {  // 0 = never, 1 = automatic, 2 = whenever a launch has >= 4 passes
        ctx->opt_long_mode = value;
        return BA_OK;
    }
    if (!strcmp(key, "trace_mode")) {  // 0 = thread per pair (static), 1 = automatic, 2 = warp per pair, 3 = CTA per pair,
                                       // 4 = split (test hook), 5 = thread per pair (persistent)
        ctx->opt_trace_mode = value;
        return BA_OK;
    }
    if (!strcmp(key, "exclusive_compute")) {
        ctx->opt_exclusive = value;
        return BA_OK;
    }
    if (!strcmp(key, "sym_kernels")) {  // 0: never use the symmetric-gap fill variant (test hook)
        ctx->opt_sym = value;
        return BA_OK;
    }
    if (!strcmp(key, "plan_cache")) {
        ctx->opt_plan_cache = value;
        plan_cache_drop(ctx);
        return BA_OK;
    }
    if (!strcmp(key, "force_path")) {
        ctx->opt_force_path = value;
        return BA_OK;
    }
    ctx->err = std::string("unknown option: ") + key;
    return BA_ERR_ARG;
}

int ba_set_scoring(ba_ctx *ctx, int kind, int affine, const int64_t *la, int let, int alpha_len,
                   int64_t gap_open, const int32_t *index)
{
    if (!ctx || !la || !index || let < 1 || kind < BA_SW || kind > BA_FITTED) return BA_ERR_ARG;
    if (!ctx->d_sc) {
        ctx->err = "context has no device (ba_create failed)";
        return BA_ERR_CUDA;
    }
    // align/sw_letters.go:70-72 (absent in fitted_letters.go:71-78)
    if (kind != BA_FITTED && let < alpha_len) {
        ctx->err = "align: scoring matrix size does not match alphabet length";
        return BA_ERR_MATRIX_SIZE;
    }
    if (let > BA_MAX_LET) {
        ctx->err = "scoring matrix larger than 32x32 is not supported";
        return BA_ERR_RANGE;
    }
    int64_t maxabs = 0;
    for (int k = 0; k < let * let; k++) {
        int64_t a = la[k] < 0 ? -la[k] : la[k];
        if (a > maxabs) maxabs = a;
    }
    const int64_t ago = gap_open < 0 ? -gap_open : gap_open;
    if (maxabs > (1 << 20) || ago > (1 << 20)) {
        ctx->err = "scoring values exceed the 32-bit device range";
        return BA_ERR_RANGE;
    }
    DeviceScoring s;
    memset(&s, 0, sizeof s);
    for (int k = 0; k < let * let; k++) s.la[k] = (int32_t)la[k];
    s.let = let;
    s.gap_open = affine ? (int32_t)gap_open : 0;
    s.alpha_len = alpha_len;
    s.kind = kind;
    s.affine = affine ? 1 : 0;
    for (int k = 0; k < 256; k++) s.index[k] = (index[k] < 0 || index[k] > 253) ? 0xFF : (uint8_t)index[k];
    // the matrix is re-read every call (callers may mutate it), but an unchanged one
    // keeps the device copy and the cached plans
    if (ctx->have_scoring && memcmp(&s, &ctx->h_sc, sizeof s) == 0) return BA_OK;
    ctx->h_sc = s;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d_sc, &ctx->h_sc, sizeof s, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_scoring = true;
    ctx->scoring_epoch++;  // cached plans depend on the scoring (eligibility, kind)
    ctx->occ_cache.clear(); // dynamic shared memory of the GSUB kernels depends on the matrix size
    return BA_OK;
}

static int check_range(ba_ctx *ctx, const int32_t *ref_len, const int32_t *qry_len, int64_t n)
{
    int64_t maxabs = 0;
    const DeviceScoring &s = ctx->h_sc;
    for (int k = 0; k < s.let * s.let; k++) {
        int64_t a = s.la[k] < 0 ? -(int64_t)s.la[k] : s.la[k];
        if (a > maxabs) maxabs = a;
    }
    int64_t maxlen = 0;
    for (int64_t p = 0; p < n; p++) {
        if (ref_len[p] < 0 || qry_len[p] < 0) {
            ctx->err = "negative sequence length";
            return BA_ERR_ARG;
        }
        int64_t l = (int64_t)ref_len[p] + qry_len[p];
        if (l > maxlen) maxlen = l;
    }
    const int64_t ago = s.gap_open < 0 ? -(int64_t)s.gap_open : s.gap_open;
    const int64_t bound = ago + (maxlen + 2) * maxabs;
    if (bound * BA_SCALE >= (1LL << 26)) {
        ctx->err = "score bound exceeds the 32-bit device lanes";
        return BA_ERR_RANGE;
    }
    return BA_OK;
}

int ba_align_batch_device(ba_ctx *ctx, const uint8_t *d_letters, int64_t letters_bytes, int stride,
                          const int64_t *d_ref_off, const int32_t *d_ref_len, const int64_t *d_qry_off,
                          const int32_t *d_qry_len, const int32_t *h_ref_len, const int32_t *h_qry_len,
                          int64_t n, void *cuda_stream, ba_result *out)
{
    if (!ctx || !out || n < 0 || (stride != 1 && stride != 2)) return BA_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (!ctx->have_scoring) {
        ctx->err = "ba_set_scoring has not been called";
        return BA_ERR_NO_SCORING;
    }
    if (n > 0 && (!d_ref_off || !d_ref_len || !d_qry_off || !d_qry_len || !h_ref_len || !h_qry_len))
        return BA_ERR_ARG;
    if (n > 0x7fffffff) return BA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_range(ctx, h_ref_len, h_qry_len, n);
    if (rc != BA_OK) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], st));
    ctx->last_n = -1;  // the context's own letters buffer no longer mirrors the last batch (ba_format_batch)
    {
        // same kernel-phase token as ba_align_batch: with several contexts the result copies and
        // the host gather of one batch run under the kernels of the next
        std::unique_lock<std::mutex> token(g_compute_token[ctx->device & 63], std::defer_lock);
        if (ctx->opt_exclusive) token.lock();
        rc = run_device(ctx, d_letters, letters_bytes, stride, d_ref_off, d_ref_len, d_qry_off, d_qry_len,
                        h_ref_len, h_qry_len, n, st, out, ctx->opt_exclusive ? &token : nullptr);
    }
    out->ms_h2d = 0.f;
    if (rc != BA_OK) ba_free_result(ctx, out);
    return rc;
}

int ba_align_batch(ba_ctx *ctx, const uint8_t *letters, int stride, const int64_t *ref_off,
                   const int32_t *ref_len, const int64_t *qry_off, const int32_t *qry_len, int64_t n,
                   ba_result *out)
{
    if (!ctx || !out || n < 0 || (stride != 1 && stride != 2)) return BA_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (!ctx->have_scoring) {
        ctx->err = "ba_set_scoring has not been called";
        return BA_ERR_NO_SCORING;
    }
    if (n > 0 && (!ref_off || !ref_len || !qry_off || !qry_len)) return BA_ERR_ARG;
    if (n > 0x7fffffff) return BA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_range(ctx, ref_len, qry_len, n);
    if (rc != BA_OK) return rc;
    int64_t letters_bytes = 0;
    for (int64_t p = 0; p < n; p++) {
        if (ref_off[p] < 0 || qry_off[p] < 0) {
            ctx->err = "negative sequence offset";
            return BA_ERR_ARG;
        }
        int64_t e1 = ref_len[p] ? ref_off[p] + ((int64_t)ref_len[p] - 1) * stride + 1 : 0;
        int64_t e2 = qry_len[p] ? qry_off[p] + ((int64_t)qry_len[p] - 1) * stride + 1 : 0;
        if (e1 > letters_bytes) letters_bytes = e1;
        if (e2 > letters_bytes) letters_bytes = e2;
    }
    if (letters_bytes > 0 && !letters) return BA_ERR_ARG;
    cudaStream_t st = ctx->stream;
    ctx->last_n = -1;
    if ((rc = ensure(ctx, ctx->letters, (size_t)letters_bytes + 16)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->ref_off, sizeof(int64_t) * (size_t)(n + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->qry_off, sizeof(int64_t) * (size_t)(n + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->ref_len, sizeof(int32_t) * (size_t)(n + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->qry_len, sizeof(int32_t) * (size_t)(n + 1))) != BA_OK) return rc;
    CK(cudaEventRecord(ctx->ev[0], st));
    // The offset/length arrays are the caller's pageable memory: a copy straight from them would
    // block inside the driver behind the letters copy.  Stage them in pinned memory and send them
    // first; the letters buffer may be pinned (ba_alloc_host) or not.
    uint8_t *stage = nullptr;
    if (n > 0) {
        const size_t b8 = sizeof(int64_t) * (size_t)n, b4 = sizeof(int32_t) * (size_t)n;
        stage = (uint8_t *)pinned_get(ctx, 2 * b8 + 2 * b4);
        if (!stage) {
            ctx->err = "pinned host allocation failed";
            return BA_ERR_NOMEM;
        }
        memcpy(stage, ref_off, b8);
        memcpy(stage + b8, qry_off, b8);
        memcpy(stage + 2 * b8, ref_len, b4);
        memcpy(stage + 2 * b8 + b4, qry_len, b4);
        CK(cudaMemcpyAsync(ctx->ref_off.p, stage, b8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->qry_off.p, stage + b8, b8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->ref_len.p, stage + 2 * b8, b4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->qry_len.p, stage + 2 * b8 + b4, b4, cudaMemcpyHostToDevice, st));
    }
    // in pieces: the small H2D copies of another context's kernel phase (work lists) share the copy
    // engine and must not queue behind one 100+ MB transfer
    for (int64_t o = 0; o < letters_bytes; o += (4 << 20)) {
        const size_t len = (size_t)std::min<int64_t>(4 << 20, letters_bytes - o);
        CK(cudaMemcpyAsync((uint8_t *)ctx->letters.p + o, letters + o, len, cudaMemcpyHostToDevice, st));
    }
    {
        std::unique_lock<std::mutex> token(g_compute_token[ctx->device & 63], std::defer_lock);
        if (ctx->opt_exclusive) {
            // inputs on the device first, then wait for the GPU's kernel phase to be free
            CK(cudaEventRecord(ctx->ev[7], st));
            CK(cudaEventSynchronize(ctx->ev[7]));
            token.lock();
        }
        rc = run_device(ctx, (const uint8_t *)ctx->letters.p, letters_bytes, stride, (const int64_t *)ctx->ref_off.p,
                        (const int32_t *)ctx->ref_len.p, (const int64_t *)ctx->qry_off.p,
                        (const int32_t *)ctx->qry_len.p, ref_len, qry_len, n, st, out,
                        ctx->opt_exclusive ? &token : nullptr);
    }
    if (rc != BA_OK) cudaStreamSynchronize(st);  // the staged copies may still be in flight
    if (stage) pinned_put(ctx, stage);           // on success run_device returned with the stream drained
    if (rc != BA_OK) {
        ba_free_result(ctx, out);
        return rc;
    }
    if (n > 0) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        out->ms_h2d = ms;
    }
    ctx->last_n = n;
    ctx->last_stride = stride;
    ctx->last_letters_bytes = letters_bytes;
    return BA_OK;
}

int ba_format_batch(ba_ctx *ctx, const uint8_t *letters, int stride, const int64_t *ref_off,
                    const int32_t *ref_len, const int64_t *qry_off, const int32_t *qry_len, int64_t n,
                    const int64_t *seg_start, const int32_t *seg_count, const int32_t *segs, int64_t n_segs,
                    uint8_t gap, ba_formatted *out)
{
    if (!ctx || !out || n < 0 || n_segs < 0 || (stride != 1 && stride != 2)) return BA_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (!ctx->d_sc) {
        ctx->err = "context has no device (ba_create failed)";
        return BA_ERR_CUDA;
    }
    if (n > 0 && (!seg_start || !seg_count || (n_segs > 0 && !segs))) return BA_ERR_ARG;
    if (n > 0 && (!ref_len || !qry_len)) return BA_ERR_ARG;
    if (n > 0x7fffffff) return BA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    // every segment must lie inside its pair (the kernels trust the coordinates)
    for (int64_t p = 0; p < n; p++) {
        if (seg_count[p] < 0 || seg_start[p] < 0 || seg_start[p] + seg_count[p] > n_segs) {
            ctx->err = "segment range outside the segment array";
            return BA_ERR_ARG;
        }
        const int32_t *s = segs + 5 * seg_start[p];
        for (int k = 0; k < seg_count[p]; k++, s += 5)
            if (s[0] < 0 || s[1] < s[0] || s[1] > ref_len[p] || s[2] < 0 || s[3] < s[2] || s[3] > qry_len[p]) {
                ctx->err = "segment outside its sequences";
                return BA_ERR_ARG;
            }
    }
    const uint8_t *d_letters;
    const int64_t *d_ref_off, *d_qry_off;
    if (!letters) {
        if (ctx->last_n != n || ctx->last_stride != stride) {
            ctx->err = "letters == NULL needs a preceding ba_align_batch call of the same batch on this context";
            return BA_ERR_ARG;
        }
    } else {
        if (n > 0 && (!ref_off || !qry_off)) return BA_ERR_ARG;
        int64_t letters_bytes = 0;
        for (int64_t p = 0; p < n; p++) {
            if (ref_off[p] < 0 || qry_off[p] < 0 || ref_len[p] < 0 || qry_len[p] < 0) return BA_ERR_ARG;
            int64_t e1 = ref_len[p] ? ref_off[p] + ((int64_t)ref_len[p] - 1) * stride + 1 : 0;
            int64_t e2 = qry_len[p] ? qry_off[p] + ((int64_t)qry_len[p] - 1) * stride + 1 : 0;
            letters_bytes = std::max(letters_bytes, std::max(e1, e2));
        }
        if ((rc = ensure(ctx, ctx->letters, (size_t)letters_bytes + 16)) != BA_OK) return rc;
        if ((rc = ensure(ctx, ctx->ref_off, sizeof(int64_t) * (size_t)(n + 1))) != BA_OK) return rc;
        if ((rc = ensure(ctx, ctx->qry_off, sizeof(int64_t) * (size_t)(n + 1))) != BA_OK) return rc;
        ctx->last_n = -1;  // the device copy now belongs to this call
        if (letters_bytes > 0)
            CK(cudaMemcpyAsync(ctx->letters.p, letters, (size_t)letters_bytes, cudaMemcpyHostToDevice, st));
        if (n > 0) {
            CK(cudaMemcpyAsync(ctx->ref_off.p, ref_off, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(ctx->qry_off.p, qry_off, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
        }
    }
    d_letters = (const uint8_t *)ctx->letters.p;
    d_ref_off = (const int64_t *)ctx->ref_off.p;
    d_qry_off = (const int64_t *)ctx->qry_off.p;

    out->n = n;
    out->off_a = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)(n + 1));
    out->off_b = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)(n + 1));
    if (!out->off_a || !out->off_b) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    out->off_a[0] = out->off_b[0] = 0;
    if (n == 0) return BA_OK;

    if ((rc = ensure(ctx, ctx->fmt_segs, sizeof(int32_t) * 5 * (size_t)(n_segs + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->fmt_seg_start, sizeof(int64_t) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->fmt_seg_count, sizeof(int32_t) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->fmt_len, sizeof(int32_t) * 2 * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->fmt_off, sizeof(int64_t) * 2 * (size_t)(n + 1))) != BA_OK) return rc;
    if (n_segs > 0)
        CK(cudaMemcpyAsync(ctx->fmt_segs.p, segs, sizeof(int32_t) * 5 * (size_t)n_segs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->fmt_seg_start.p, seg_start, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->fmt_seg_count.p, seg_count, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[1], st));

    FormatParams fp;
    fp.letters = d_letters;
    fp.stride = stride;
    fp.ref_off = d_ref_off;
    fp.qry_off = d_qry_off;
    fp.seg_start = (const int64_t *)ctx->fmt_seg_start.p;
    fp.seg_count = (const int32_t *)ctx->fmt_seg_count.p;
    fp.segs = (const int32_t *)ctx->fmt_segs.p;
    fp.n = n;
    fp.len_a = (int32_t *)ctx->fmt_len.p;
    fp.len_b = fp.len_a + n;
    int64_t *d_off_a = (int64_t *)ctx->fmt_off.p, *d_off_b = d_off_a + (n + 1);
    fp.off_a = d_off_a;
    fp.off_b = d_off_b;
    fp.row_a = fp.row_b = nullptr;
    fp.gap = gap;
    BA_LAUNCH(ba_format_size_kernel, (unsigned)((n + 255) / 256), 256, st, fp);
    size_t scan_bytes = 0;
    if ((rc = exclusive_scan(ctx, nullptr, scan_bytes, fp.len_a, d_off_a, (int)n, st)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->scan_tmp, scan_bytes + 16)) != BA_OK) return rc;
    size_t tb = ctx->scan_tmp.cap;
    if ((rc = exclusive_scan(ctx, ctx->scan_tmp.p, tb, fp.len_a, d_off_a, (int)n, st)) != BA_OK) return rc;
    tb = ctx->scan_tmp.cap;
    if ((rc = exclusive_scan(ctx, ctx->scan_tmp.p, tb, fp.len_b, d_off_b, (int)n, st)) != BA_OK) return rc;
    // row sizes to the host: the last offsets + the last lengths give the totals
    int32_t *h_len = (int32_t *)pinned_get(ctx, sizeof(int32_t) * 2 * (size_t)n);
    if (!h_len) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    CK(cudaMemcpyAsync(out->off_a, d_off_a, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->off_b, d_off_b, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_len, fp.len_a, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out->bytes_a = out->off_a[n - 1] + h_len[n - 1];
    out->bytes_b = out->off_b[n - 1] + h_len[2 * n - 1];
    out->off_a[n] = out->bytes_a;
    out->off_b[n] = out->bytes_b;
    pinned_put(ctx, h_len);
    if ((rc = ensure(ctx, ctx->fmt_rows, (size_t)(out->bytes_a + out->bytes_b) + 32)) != BA_OK) return rc;
    fp.row_a = (uint8_t *)ctx->fmt_rows.p;
    fp.row_b = fp.row_a + out->bytes_a;
    int64_t wblocks = (n + 7) / 8;
    if (wblocks > (int64_t)ctx->sm_count * 8) wblocks = (int64_t)ctx->sm_count * 8;
    BA_LAUNCH(ba_format_write_kernel, (unsigned)wblocks, 256, st, fp);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[2], st));
    out->row_a = (uint8_t *)pinned_get(ctx, (size_t)out->bytes_a + 1);
    out->row_b = (uint8_t *)pinned_get(ctx, (size_t)out->bytes_b + 1);
    if (!out->row_a || !out->row_b) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    if (out->bytes_a > 0)
        CK(cudaMemcpyAsync(out->row_a, fp.row_a, (size_t)out->bytes_a, cudaMemcpyDeviceToHost, st));
    if (out->bytes_b > 0)
        CK(cudaMemcpyAsync(out->row_b, fp.row_b, (size_t)out->bytes_b, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
    out->ms_kernels = ms;
    return BA_OK;
}

void ba_free_formatted(ba_ctx *ctx, ba_formatted *f)
{
    if (!ctx || !f) return;
    if (f->off_a) pinned_put(ctx, f->off_a);
    if (f->off_b) pinned_put(ctx, f->off_b);
    if (f->row_a) pinned_put(ctx, f->row_a);
    if (f->row_b) pinned_put(ctx, f->row_b);
    memset(f, 0, sizeof *f);
}

void ba_free_result(ba_ctx *ctx, ba_result *res)
{
    if (!ctx || !res) return;
    if (res->status) pinned_put(ctx, res->status);
    if (res->err_pos) pinned_put(ctx, res->err_pos);
    if (res->seg_start) pinned_put(ctx, res->seg_start);
    if (res->seg_count) pinned_put(ctx, res->seg_count);
    if (res->segs) pinned_put(ctx, res->segs);
    memset(res, 0, sizeof *res);
}

void *ba_alloc_host(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}

void ba_free_host(void *p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"

#include "ba_seqio.inl"
