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
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
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
#include "ba_matrices.inl"
#include "ba_pals.cuh"

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
    uint64_t stamp;   // last release (pool trimming frees the least recently used blocks first)
};

struct ToInt64 {
    __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};

#define BA_ERR_LONG_TIMEOUT (-100)   // internal: a wavefront pass gave up waiting; the batch is re-run without LONG

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
    cudaStream_t stream = nullptr;   // main stream: encode, validate, fills of the even waves, exact path
    cudaStream_t s_fill2 = nullptr;  // fills of the odd waves (the tail of one fill overlaps the head of the next)
    cudaStream_t s_trace = nullptr;  // traceback + compaction of wave w beside the fill of wave w+1 (high priority)
    cudaStream_t s_h2d = nullptr;    // input upload, in groups the first waves do not have to wait for
    cudaStream_t s_d2h = nullptr;    // result copies
    cudaStream_t s_prep = nullptr;   // encode + validate of upload groups (host-buffer calls)
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // pooled events of one call (timing / no timing), handed out in order and reused by the next call
    std::vector<cudaEvent_t> ev_timed, ev_plain, ev_upload;
    size_t ev_timed_next = 0, ev_plain_next = 0;
    std::string err;
    bool have_scoring = false;
    DeviceScoring h_sc;
    DeviceScoring *d_sc = nullptr;        // the current scoring's device copy (one of sc_cache)
    DeviceScoring *d_sc_first = nullptr;  // allocated by ba_create (a context without it has no device)
    struct ScoringSlot {
        DeviceScoring h;
        DeviceScoring *d;
        uint64_t stamp = 0;
    };
    std::vector<ScoringSlot> sc_cache;
    int64_t scoring_uploads = 0;
    int64_t opt_arena_bytes = 0;
    // grow-only device buffers
    DevBuf letters, enc, ref_off, ref_len, qry_off, qry_len, status, err_pos, start;
    DevBuf items, aux_off, aux, bnd, seg_count, seg_base, segs, arena, scan_tmp, counters;
    DevBuf flags, slot_off, slot_cap, slots, seg_clip, tasks, prog;
    DevBuf pair_seg_start, pair_seg_count, meta, ovf;
    DevBuf pals_meta, pals_traps, pals_vec, pals_stack, pals_hits;   // align/pals/dp batches (ba_pals.inl)
    DevBuf segs32;   // exact path's per-wave segments (ctx->segs keeps the 16-bit path's until they are fetched)
    // Plan cache: batches of the same shape (same length arrays, same scoring) reuse the split, the
    // work lists and their device copies.
    struct PlanCache {
        bool valid = false, plan_built = false;
        int64_t n = 0, cells = 0, arena_cap = 0, slot_cap_opt = 0, waves_opt = 0;
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
    // overlap scheduling (see run_path16): waves per batch (0 = automatic), fill streams (1 or 2), claims a
    // fill warp takes before its CTA retires (0 = persistent grid), resident fill CTAs per SM (0 = all that
    // fit), threads per thread-per-pair traceback CTA (0 = automatic), upload groups (0 = automatic)
    int64_t opt_waves = 0, opt_fill_streams = 2, opt_fill_claims = 1, opt_fill_cap = 0, opt_trace_cta = 0;
    int64_t opt_upload_groups = 0, opt_chunk_pairs = 8192, opt_test_long_timeout = 0, opt_wide10 = 1;
    int64_t cached_arena_cap = 0;                    // cudaMemGetInfo is slow: ask once per context
    std::vector<std::pair<const void *, int>> occ_cache;  // kernel -> resident CTAs per SM
    std::vector<PinnedBlock> pinned;
    int64_t deferred_dev_segs = 0;                   // multi-device: segments of the last batch still only in ctx->segs
    std::mutex pinned_mu;                            // ba_free_result may run on another thread (GC finalizers)
    int64_t pinned_budget = 0;                       // bytes of unused pinned blocks kept for reuse (0 = default)
    size_t pinned_peak = 0;                          // most pinned bytes in use at once so far
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

// Pinned host memory pool of a context.  Blocks are reused best-fit; unused blocks beyond a budget
// (default 1 GiB, option "pinned_budget") are returned to the driver, least recently used first, so a
// long-lived context fed batches of varying size does not accumulate page-locked memory.
void *pinned_get(ba_ctx *ctx, size_t bytes)
{
    if (bytes == 0) bytes = 8;
    std::lock_guard<std::mutex> lk(ctx->pinned_mu);
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
    ctx->pinned.push_back({p, cap, true, 0});
    return p;
}

void pinned_put(ba_ctx *ctx, void *p)
{
    if (!p) return;
    static std::atomic<uint64_t> clock_{1};
    std::lock_guard<std::mutex> lk(ctx->pinned_mu);
    size_t idle = 0;
    for (auto &b : ctx->pinned) {
        if (b.p == p) {
            b.used = false;
            b.stamp = clock_++;
        }
        if (!b.used) idle += b.cap;
    }
    // default budget: what the largest batch so far had in use at once (its blocks come back on the next
    // call of the same size), at least 1 GiB
    size_t in_use = 0;
    for (auto &b : ctx->pinned)
        if (b.used || b.p == p) in_use += b.cap;
    ctx->pinned_peak = std::max(ctx->pinned_peak, in_use);
    const size_t budget = ctx->pinned_budget > 0 ? (size_t)ctx->pinned_budget
                                                 : std::max((size_t)1 << 30, ctx->pinned_peak + ctx->pinned_peak / 2);
    while (idle > budget) {
        int old = -1;
        for (size_t k = 0; k < ctx->pinned.size(); k++)
            if (!ctx->pinned[k].used && (old < 0 || ctx->pinned[k].stamp < ctx->pinned[old].stamp)) old = (int)k;
        if (old < 0) break;
        idle -= ctx->pinned[old].cap;
        cudaFreeHost(ctx->pinned[old].p);
        ctx->pinned.erase(ctx->pinned.begin() + old);
    }
}

// Pinned blocks a call holds: released on every exit path unless handed to the caller (ba_result).
struct Lease {
    ba_ctx *ctx;
    std::vector<void *> held;
    explicit Lease(ba_ctx *c) : ctx(c) {}
    ~Lease()
    {
        for (void *p : held) pinned_put(ctx, p);
    }
    void *get(size_t bytes)
    {
        void *p = pinned_get(ctx, bytes);
        if (p) held.push_back(p);
        return p;
    }
    void keep(void *p)   // ownership moves to the caller's result
    {
        for (size_t k = 0; k < held.size(); k++)
            if (held[k] == p) {
                held.erase(held.begin() + k);
                return;
            }
    }
    void drop(void *p)
    {
        keep(p);
        pinned_put(ctx, p);
    }
};

cudaEvent_t ev_new(ba_ctx *ctx, bool timed)
{
    std::vector<cudaEvent_t> &pool = timed ? ctx->ev_timed : ctx->ev_plain;
    size_t &next = timed ? ctx->ev_timed_next : ctx->ev_plain_next;
    if (next == pool.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, timed ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        pool.push_back(e);
    }
    return pool[next++];
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
// Input upload groups: pairs [p_lo, p_hi) reference only letters below byte b_hi, and the group's own
// upload (and encode) covers bytes [b_lo, b_hi).  A wave may start as soon as the groups up to its
// largest pair id are on the device, encoded and validated -- the rest of the batch uploads under its
// kernels (ba_align_batch); device-resident callers have one group that is ready from the start.
struct UploadGroup {
    int64_t p_lo, p_hi, b_lo, b_hi;
    cudaEvent_t ready = nullptr;    // recorded on the upload stream after this group's copy (nullptr: resident)
    cudaEvent_t checked = nullptr;  // recorded on the main stream after encode + validate of the group
};

struct Batch {
    const uint8_t *d_letters = nullptr;
    uint8_t *d_enc = nullptr;
    int stride = 1;
    const int64_t *d_ref_off = nullptr, *d_qry_off = nullptr;
    const int32_t *d_ref_len = nullptr, *d_qry_len = nullptr;
    const int32_t *h_ref_len = nullptr, *h_qry_len = nullptr;
    int64_t n = 0;
    cudaStream_t st = nullptr;  // main stream (the caller's for ba_align_batch_device)
    // encode + validate stream: the main stream, or -- while the input is still uploading in groups -- a
    // stream of its own, so that the preparation of group g+1 is not queued behind the fill of chunk g
    cudaStream_t prep_st = nullptr;
    int32_t *d_status = nullptr;
    int64_t *d_err_pos = nullptr;
    uint8_t *d_flags = nullptr;   // letter-class flags per pair (16-bit path eligibility), or nullptr
    PairStart *d_start = nullptr;
    ba_result *out = nullptr;
    Lease *lease = nullptr;
    std::vector<UploadGroup> groups;
    size_t next_group = 0;
    // groups whose copy (and `ready` record) has been ENQUEUED; nullptr when all of them were before the
    // call started.  A feeder thread gathers and enqueues the groups of a scattered share while the
    // kernels of the first groups are already running; waiting on an event that has not been recorded
    // yet would not wait at all, so the consumer first waits here on the host.
    std::atomic<int64_t> *groups_enqueued = nullptr;
    // all segments of the batch in ONE pinned block (grown in the rare case the estimate was short)
    int32_t *h_segs = nullptr;
    int64_t h_segs_cap = 0;     // in segments
    int64_t seg_total = 0;      // segments produced so far == next free position
    // multi-device batches: the compacted segments stay on the device until every device knows where its
    // share starts in the common result block; then they travel straight to their final place
    bool defer_segs = false;
    int64_t dev_segs = 0;       // segments compacted on the device by the 16-bit path (prefix of the batch's segments)
    cudaEvent_t ev_d2h0 = nullptr;  // first result copy of the call (ms_d2h)
    float ms_fill = 0.f, ms_trace = 0.f;
    double host_s = 0.0;
    int64_t launches = 0, fill_launches = 0, code_bytes = 0, waves = 0;
};

static bool g_timing = getenv("BA_TIMING") != nullptr;
// exclusive_compute: ba_align_batch calls on one device run their FILL phases one after the other
// even when they come from several contexts on several threads -- as a DEVICE-side dependency (the
// first fill of a batch waits for the event the previous batch recorded after its last fill), so no
// host thread ever blocks for it.  Uploads, encode/validate, the tail traceback and the result
// copies of neighbouring batches overlap freely.
struct DeviceGate {
    std::mutex mu;
    cudaEvent_t last_fill = nullptr;
    bool armed = false;
};
static DeviceGate g_gate[64];
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

// Encode + validate every upload group up to the one holding pair `pmax` (main stream, in order).
int prep_pairs(ba_ctx *ctx, Batch &B, int64_t pmax)
{
    const int kind = ctx->h_sc.kind, affine = ctx->h_sc.affine;
    cudaStream_t PS = B.prep_st ? B.prep_st : B.st;
    while (B.next_group < B.groups.size() && B.groups[B.next_group].p_lo <= pmax) {
        UploadGroup &g = B.groups[B.next_group++];
        if (B.groups_enqueued)
            while (B.groups_enqueued->load(std::memory_order_acquire) < (int64_t)B.next_group) std::this_thread::yield();
        if (g.ready) CK(cudaStreamWaitEvent(PS, g.ready, 0));
        const int64_t nb = g.b_hi - g.b_lo, np = g.p_hi - g.p_lo;
        if (nb > 0) {
            int64_t blocks = (nb / 16 + 255) / 256;
            if (blocks < 1) blocks = 1;
            if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
            BA_LAUNCH(ba_encode_kernel, (unsigned)blocks, 256, PS, B.d_letters + g.b_lo, B.d_enc + g.b_lo, nb,
                      ctx->d_sc);
            B.launches++;
        }
        if (np > 0) {
            int64_t blocks = (np + 7) / 8;
            if (blocks > ctx->sm_count * 32) blocks = ctx->sm_count * 32;
            BA_LAUNCH(ba_validate_kernel, (unsigned)blocks, 256, PS, B.d_enc, B.stride, B.d_ref_off + g.p_lo,
                      B.d_ref_len + g.p_lo, B.d_qry_off + g.p_lo, B.d_qry_len + g.p_lo, np, kind, affine,
                      B.d_status + g.p_lo, B.d_err_pos + g.p_lo, B.d_flags ? B.d_flags + g.p_lo : nullptr);
            B.launches++;
        }
        CK(cudaGetLastError());
        g.checked = ev_new(ctx, false);
        if (!g.checked) {
            ctx->err = "cudaEventCreate failed";
            return BA_ERR_CUDA;
        }
        CK(cudaEventRecord(g.checked, PS));
    }
    return BA_OK;
}

// Make `stream` see everything prepared so far (no-op when it IS the preparation stream).
int sync_prep(ba_ctx *ctx, Batch &B, cudaStream_t stream)
{
    if (!B.prep_st || B.prep_st == stream || B.next_group == 0) return BA_OK;
    CK(cudaStreamWaitEvent(stream, B.groups[B.next_group - 1].checked, 0));
    return BA_OK;
}

// Room for `need` segments in the batch's pinned segment block; copies already enqueued on the
// result stream land first when the block has to move.
int segs_reserve(ba_ctx *ctx, Batch &B, int64_t need, int64_t valid)
{
    if (need <= B.h_segs_cap) return BA_OK;
    const int64_t cap = need + need / 8 + 1024;
    int32_t *nb = (int32_t *)B.lease->get(sizeof(int32_t) * 5 * (size_t)cap);
    if (!nb) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    if (B.h_segs) {
        CK(cudaStreamSynchronize(ctx->s_d2h));
        if (valid > 0) memcpy(nb, B.h_segs, sizeof(int32_t) * 5 * (size_t)valid);
        B.lease->drop(B.h_segs);
    }
    B.h_segs = nb;
    B.h_segs_cap = cap;
    return BA_OK;
}

// Exact path: D2H of one wave's compact segments behind the segments the batch already has, and the
// host-side scatter of seg_start/seg_count by pair id.
int gather_wave(ba_ctx *ctx, Batch &B, const Plan &plan, const Plan::Wave &w, const int32_t *h_cnt,
                const int64_t *h_base, int64_t wsegs, const int32_t *d_segs)
{
    cudaStream_t st = B.st;
    int rc;
    if ((rc = segs_reserve(ctx, B, B.seg_total + wsegs, B.seg_total)) != BA_OK) return rc;
    if (!B.ev_d2h0) {
        B.ev_d2h0 = ev_new(ctx, true);
        if (B.ev_d2h0) CK(cudaEventRecord(B.ev_d2h0, st));
    }
    if (wsegs > 0)
        CK(cudaMemcpyAsync(B.h_segs + 5 * B.seg_total, d_segs, sizeof(int32_t) * 5 * (size_t)wsegs,
                           cudaMemcpyDeviceToHost, st));
    // the per-pair scatter only needs the counts (already on the host): do it while the segments travel
    const double t0 = now_s();
    for (int64_t k = 0; k < w.count; k++) {
        const int32_t p = plan.items[w.first + k].pair;
        B.out->seg_start[p] = B.seg_total + h_base[k];
        B.out->seg_count[p] = h_cnt[k];
    }
    BA_TSEC("gather scatter", t0);
    B.host_s += now_s() - t0;
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    B.ms_fill += ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]));
    B.ms_trace += ms;
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
    if ((rc = prep_pairs(ctx, B, B.n - 1)) != BA_OK) return rc;
    if ((rc = sync_prep(ctx, B, st)) != BA_OK) return rc;
    ctx->pc.dev_items = nullptr;  // this path reuses the work-list buffers
    int64_t arena_cap = 0;
    if (arena_budget(ctx, arena_cap) != 0) {
        ctx->err = "cudaMemGetInfo failed";
        return BA_ERR_CUDA;
    }
    Plan plan;
    const double tp0 = now_s();
    plan.items = (WorkItem *)B.lease->get(sizeof(WorkItem) * (size_t)nids);
    plan.aux_off = (int64_t *)B.lease->get(sizeof(int64_t) * (size_t)nids);
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
    int32_t *h_cnt = (int32_t *)B.lease->get(sizeof(int32_t) * (size_t)max_items);
    int64_t *h_base = (int64_t *)B.lease->get(sizeof(int64_t) * (size_t)(max_items + 1));
    if (!h_cnt || !h_base) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    WalkFn walk_count = pick_walk<false>(kind, affine);
    WalkFn walk_write = pick_walk<true>(kind, affine);

    for (auto &w : plan.waves) {
        B.waves++;
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
        if ((rc = ensure(ctx, ctx->segs32, sizeof(int32_t) * 5 * (size_t)(wsegs + 1))) != BA_OK) return rc;
        wp.segs = (int32_t *)ctx->segs32.p;
        BA_LAUNCH(walk_write, wblocks, 128, st, wp);
        B.launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[6], st));
        if ((rc = gather_wave(ctx, B, plan, w, h_cnt, h_base, wsegs, (const int32_t *)ctx->segs32.p)) != BA_OK)
            return rc;
    }
    B.lease->drop(h_cnt);
    B.lease->drop(h_base);
    B.lease->drop(plan.items);
    B.lease->drop(plan.aux_off);
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
    // (SW: the biased kernels add -(go + e) to every value; the head room of condition C4 covers up to 2000)
    return s.affine && er == eq && s.gap_open + er <= 0 && (s.kind != KIND_SW || s.gap_open + er >= -2000);
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
template <int NT, int MC>
Trace16Fn pick_trace16t_mc(int kind, bool lin)
{
    if (lin) {
        if (kind == KIND_SW) return ba_trace16w_kernel<KIND_SW, true, NT, MC>;
        if (kind == KIND_NW) return ba_trace16w_kernel<KIND_NW, true, NT, MC>;
        return ba_trace16w_kernel<KIND_FIT, true, NT, MC>;
    }
    if (kind == KIND_SW) return ba_trace16w_kernel<KIND_SW, false, NT, MC>;
    if (kind == KIND_NW) return ba_trace16w_kernel<KIND_NW, false, NT, MC>;
    return ba_trace16w_kernel<KIND_FIT, false, NT, MC>;
}
// mc: widest strip of the wave (8, or 10 when it contains the 257..320-column single-pass class)
template <int NT>
Trace16Fn pick_trace16t(int kind, bool lin, int mc = BA_MAX_COLS)
{
    return mc > BA_MAX_COLS ? pick_trace16t_mc<NT, BA16_MAX_COLS>(kind, lin) : pick_trace16t_mc<NT, BA_MAX_COLS>(kind, lin);
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
template <int TS, int MC>
Trace16Fn pick_trace16ts(int kind, bool lin)
{
    if (lin) {
        if (kind == KIND_SW) return ba_trace16_kernel<KIND_SW, true, TS, MC>;
        if (kind == KIND_NW) return ba_trace16_kernel<KIND_NW, true, TS, MC>;
        return ba_trace16_kernel<KIND_FIT, true, TS, MC>;
    }
    if (kind == KIND_SW) return ba_trace16_kernel<KIND_SW, false, TS, MC>;
    if (kind == KIND_NW) return ba_trace16_kernel<KIND_NW, false, TS, MC>;
    return ba_trace16_kernel<KIND_FIT, false, TS, MC>;
}
// thread-per-pair traceback with CTAs of `ts` threads (64: fits the SM slot a retiring fill CTA frees)
Trace16Fn pick_trace16(int kind, bool lin, int ts, int mc = BA_MAX_COLS)
{
    if (mc > BA_MAX_COLS) return pick_trace16ts<BA_TRACE_THREADS, BA16_MAX_COLS>(kind, lin);   // (128 threads only)
    return ts == 64 ? pick_trace16ts<64, BA_MAX_COLS>(kind, lin) : pick_trace16ts<BA_TRACE_THREADS, BA_MAX_COLS>(kind, lin);
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

// How one fill launch of a wave runs (decided for every launch before anything is enqueued, so all
// buffers are sized once and no allocation happens between overlapping kernels).
struct FillCfg {
    Fill16Fn fn = nullptr;
    size_t dyn = 0;
    int64_t blocks = 0;
    int max_claims = 0;
    bool use_long = false;
    int maxp = 0, max_n = 0;
    int64_t ncp = 0, ntasks = 0, bnd_stride = 0, bnd_ints = 0;
    int64_t maxb = 0;   // resident CTAs of this kernel on the whole GPU
};

int run_path16(ba_ctx *ctx, Batch &B, const int32_t *ids, int64_t nids, bool gsub, std::vector<int32_t> &overflow,
               bool use_cache, int64_t cells)
{
    const int kind = ctx->h_sc.kind;
    const bool lin = !ctx->h_sc.affine;
    const int flag_mask = gsub ? 2 : 1;
    if (nids == 0) return BA_OK;
    cudaStream_t st = B.st;
    Lease &lease = *B.lease;
    int rc;
    int64_t arena_cap = 0;
    if (arena_budget(ctx, arena_cap) != 0) {
        ctx->err = "cudaMemGetInfo failed";
        return BA_ERR_CUDA;
    }
    // ---- waves: the traceback of wave w runs beside the fill of wave w+1 (different pipes bind them:
    // the fill saturates the integer pipe, the walk waits on checkpoint loads), so a batch that is
    // large enough is cut into several waves even when its checkpoints would fit the arena at once.
    // Measured on B200 (profiles/r02_overlap_sweep.md): with the present traceback kernel, whose
    // latency per launch is ~1.4 ms whatever the wave size, cutting config 2 into waves LOSES
    // (10.3 ms -> 10.9 / 12.2 / 13.7 ms for 2 / 4 / 8 waves), so the automatic setting keeps one wave
    // per arena-load; "waves" > 1 remains for experiments and is covered by the parity tests.
    (void)cells;
    int64_t tw = ctx->opt_waves;
    if (tw <= 0) tw = 1;
    const int64_t wave_cap = tw > 1 ? std::max<int64_t>(arena_cap / 3, 1) : arena_cap;
    Plan local_plan;
    ba_ctx::PlanCache &pc = ctx->pc;
    const double tp0 = now_s();
    if (use_cache && pc.plan_built &&
        (pc.arena_cap != arena_cap || pc.slot_cap_opt != ctx->opt_slot_cap || pc.waves_opt != tw)) {
        pinned_put(ctx, pc.plan.items);
        pinned_put(ctx, pc.plan.slot_off);
        pinned_put(ctx, pc.plan.slot_cap);
        pc.plan = Plan();
        pc.plan_built = false;
        pc.dev_items = nullptr;
    }
    Plan &plan = use_cache ? pc.plan : local_plan;
    if (!(use_cache && pc.plan_built)) {
        if (use_cache) {
            plan.items = (WorkItem *)pinned_get(ctx, sizeof(WorkItem) * (size_t)nids);
            plan.slot_off = (int64_t *)pinned_get(ctx, sizeof(int64_t) * (size_t)nids);
            plan.slot_cap = (int32_t *)pinned_get(ctx, sizeof(int32_t) * (size_t)nids);
        } else {
            plan.items = (WorkItem *)lease.get(sizeof(WorkItem) * (size_t)nids);
            plan.slot_off = (int64_t *)lease.get(sizeof(int64_t) * (size_t)nids);
            plan.slot_cap = (int32_t *)lease.get(sizeof(int32_t) * (size_t)nids);
        }
        if (!plan.items || !plan.slot_off || !plan.slot_cap) {
            if (use_cache) {
                pinned_put(ctx, plan.items);
                pinned_put(ctx, plan.slot_off);
                pinned_put(ctx, plan.slot_cap);
                plan = Plan();
            }
            ctx->err = "pinned host allocation failed";
            return BA_ERR_NOMEM;
        }
        const double tp1 = now_s();
        build_plan(ids, nids, B.h_ref_len, B.h_qry_len, kind, lin ? 0 : 1, true, wave_cap, plan, (int)tw, ctx->opt_wide10 != 0);
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
            pc.waves_opt = tw;
            pc.dev_items = nullptr;
        }
    }
    B.host_s += now_s() - tp0;
    const int64_t nw = (int64_t)plan.waves.size();
    // Overlap mode (non-persistent fill grids, 64-thread traceback CTAs, alternating fill streams) only when the
    // caller asked for waves: a batch that is merely too large for the arena runs its waves with the
    // single-wave kernels' settings (measured: 1 M pairs of config 2 on one GPU, 94.5 -> see DESIGN 4.4)
    const bool overlap = nw > 1 && tw > 1;
    int64_t region_bytes = 0, region_slots = 0, total_slots = 0;
    for (auto &w : plan.waves) {
        region_bytes = std::max(region_bytes, w.code_bytes);
        region_slots = std::max(region_slots, w.slots);
        total_slots += w.slots;
    }
    region_bytes = (region_bytes + 255) & ~(int64_t)255;
    // checkpoint regions: wave w uses region w % R; with R >= 3 a fill never waits for a traceback
    int64_t R = region_bytes > 0 ? arena_cap / region_bytes : nw;
    R = std::max<int64_t>(1, std::min<int64_t>(R, nw));
    const bool two_streams = overlap && ctx->opt_fill_streams >= 2 && R >= 2;

    // ---- launch configurations
    const bool sym = !lin && ctx->opt_sym != 0 && g_sym_scoring(ctx->h_sc);
    const size_t dyn = gsub ? (size_t)BA16_WARPS * ba16_gsub_bytes_per_warp(ctx->h_sc.let) : 0;
    std::vector<std::vector<FillCfg>> cfgs((size_t)nw);
    int64_t max_bnd_ints = 0, max_tasks = 0;
    for (int64_t wi = 0; wi < nw; wi++) {
        const Plan::Wave &w = plan.waves[(size_t)wi];
        const bool multi = w.max_n_multipass > 0;
        int64_t wave_bnd = 0, wave_tasks = 0;
        for (auto &L : w.launches) {
            FillCfg c;
            c.dyn = dyn;
            // ---- LONG: too few pairs to fill the GPU pair by pair, and many passes each: spread the
            // passes of every pair over warps (wavefront).  Items of a launch are sorted by cells.
            const int W16 = 32 * L.cols;
            if (multi && L.cols >= 7 && ctx->opt_long_mode != 0) {
                for (int64_t k = 0; k < L.count; k++) {
                    const int32_t p = plan.items[L.first + k].pair;
                    c.maxp = std::max(c.maxp, (B.h_qry_len[p] + W16 - 1) / W16);
                    c.max_n = std::max(c.max_n, B.h_ref_len[p]);
                }
            }
            c.fn = pick_fill16(kind, lin, L.cols, multi, gsub, false, sym);
            // GSUB kernels keep their score tables in dynamic shared memory (the size depends on the
            // matrix only, so the per-kernel cache entry stays valid until the scoring changes)
            int occ = 1;
            if ((rc = cached_occupancy(ctx, c.fn, BA16_WARPS * 32, occ, dyn)) != BA_OK) return rc;
            if (overlap && ctx->opt_fill_cap > 0) occ = (int)std::min<int64_t>(occ, ctx->opt_fill_cap);
            int64_t maxb = (int64_t)ctx->sm_count * occ;
            const int64_t pair_warps = (L.count + 3) / 4;
            c.use_long = c.maxp >= 4 && (ctx->opt_long_mode == 2 || pair_warps < maxb * BA16_WARPS) &&
                         (int64_t)c.maxp * (L.count + 4) < (1 << 22);
            if (c.use_long) {
                c.fn = pick_fill16(kind, lin, L.cols, true, gsub, true, sym);
                if ((rc = cached_occupancy(ctx, c.fn, BA16_WARPS * 32, occ, dyn)) != BA_OK) return rc;
                maxb = (int64_t)ctx->sm_count * occ;
                c.ncp = ((L.count + 3) / 4) * 2;         // couples per pass, even
                c.ntasks = (int64_t)c.maxp * 2 * c.ncp;
                c.bnd_stride = 4 * ((int64_t)c.max_n + 2);
                c.bnd_ints = c.bnd_stride * c.ncp;
                c.blocks = std::min(maxb, (c.ntasks + 4 * BA16_WARPS - 1) / (4 * BA16_WARPS));
                wave_tasks += c.ntasks;
            } else {
                const int64_t need = (L.count + 2 * BA16_GROUPS - 1) / (2 * BA16_GROUPS);  // two pairs per group
                c.maxb = maxb;
                // single-pass launches of an overlapped batch retire their CTAs after `fill_claims` claims,
                // so the high-priority traceback CTAs of the previous wave find SM slots all along
                if (overlap && !multi && ctx->opt_fill_claims > 0) {
                    c.max_claims = (int)ctx->opt_fill_claims;
                    c.blocks = (need + c.max_claims - 1) / c.max_claims;
                } else {
                    c.blocks = std::min(need, maxb);
                }
                if (multi) {
                    c.bnd_stride = 4 * ((int64_t)w.max_n_multipass + 2);
                    c.bnd_ints = c.bnd_stride * c.blocks * BA16_GROUPS;
                }
            }
            wave_bnd += c.bnd_ints;
            cfgs[(size_t)wi].push_back(c);
        }
        max_bnd_ints = std::max(max_bnd_ints, wave_bnd);
        max_tasks = std::max(max_tasks, wave_tasks);
    }

    // ---- device buffers (sized for the whole batch; nothing is allocated once kernels overlap)
    if ((rc = ensure(ctx, ctx->items, sizeof(WorkItem) * (size_t)nids)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->slot_off, sizeof(int64_t) * (size_t)(nids + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->slot_cap, sizeof(int32_t) * (size_t)(nids + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->slots, sizeof(int32_t) * 5 * (size_t)(R * region_slots + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->seg_count, sizeof(int32_t) * (size_t)nids)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->seg_clip, sizeof(int32_t) * (size_t)nids)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->seg_base, sizeof(int64_t) * (size_t)(nids + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->arena, (size_t)(R * region_bytes) + 256)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->counters, sizeof(int) * 256 * (size_t)nw)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->segs, sizeof(int32_t) * 5 * (size_t)(total_slots + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->ovf, sizeof(int32_t) * (size_t)(nids + 1))) != BA_OK) return rc;
    const size_t meta_words = BA_META_HEAD + BA_META_WAVE * (size_t)nw;
    if ((rc = ensure(ctx, ctx->meta, sizeof(long long) * meta_words)) != BA_OK) return rc;
    // two fills may be in flight (alternating streams): their boundary rows / task lists alternate too
    if (max_bnd_ints > 0 && (rc = ensure(ctx, ctx->bnd, sizeof(int32_t) * (size_t)(2 * max_bnd_ints))) != BA_OK) return rc;
    if (max_tasks > 0) {
        if ((rc = ensure(ctx, ctx->tasks, sizeof(WorkItem) * (size_t)(2 * max_tasks))) != BA_OK) return rc;
        if ((rc = ensure(ctx, ctx->prog, 64 * (size_t)nw)) != BA_OK) return rc;
    }
    size_t scan_bytes = 0;
    {
        int64_t max_items = 0;
        for (auto &w : plan.waves) max_items = std::max(max_items, w.count);
        if ((rc = exclusive_scan(ctx, nullptr, scan_bytes, (const int32_t *)ctx->seg_clip.p,
                                 (int64_t *)ctx->seg_base.p, (int)max_items, st)) != BA_OK)
            return rc;
    }
    if ((rc = ensure(ctx, ctx->scan_tmp, scan_bytes + 16)) != BA_OK) return rc;
    long long *h_meta = (long long *)lease.get(sizeof(long long) * meta_words);
    int32_t *h_long_err = (int32_t *)lease.get(sizeof(int32_t) * (size_t)(nw + 1));
    if (!h_meta || !h_long_err) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    memset(h_long_err, 0, sizeof(int32_t) * (size_t)(nw + 1));
    std::vector<WorkItem *> long_tasks;

    // ---- one-time set-up on the main stream: inputs of the first wave, work lists, counters
    // (ahead of the device gate: encode + validate run beside the previous batch's fills.  A batch whose
    // upload is still in flight only takes its first group here; the chunked fill asks for the rest.)
    if ((rc = prep_pairs(ctx, B, (!overlap && B.groups.size() > 1) ? B.groups[0].p_lo : plan.waves[0].max_pair)) != BA_OK)
        return rc;
    WorkItem *d_items = (WorkItem *)ctx->items.p;
    const bool dev_ok = use_cache && pc.dev_items == ctx->items.p && pc.dev_slot_off == ctx->slot_off.p &&
                        pc.dev_slot_cap == ctx->slot_cap.p;
    if (!dev_ok) {
        if ((rc = upload_words(ctx, d_items, plan.items, sizeof(WorkItem) * (size_t)nids, st)) != BA_OK) return rc;
        if ((rc = upload_words(ctx, ctx->slot_off.p, plan.slot_off, sizeof(int64_t) * (size_t)nids, st)) != BA_OK)
            return rc;
        if ((rc = upload_words(ctx, ctx->slot_cap.p, plan.slot_cap, sizeof(int32_t) * (size_t)nids, st)) != BA_OK)
            return rc;
        B.launches += 3;
        if (use_cache) {
            pc.dev_items = ctx->items.p;
            pc.dev_slot_off = ctx->slot_off.p;
            pc.dev_slot_cap = ctx->slot_cap.p;
        }
    }
    CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(int) * 256 * (size_t)nw, st));
    CK(cudaMemsetAsync(ctx->meta.p, 0, sizeof(long long) * meta_words, st));
    if (max_tasks > 0) CK(cudaMemsetAsync(ctx->prog.p, 0, 64 * (size_t)nw, st));
    CK(cudaMemsetAsync(ctx->pair_seg_start.p, 0, sizeof(int64_t) * (size_t)B.n, st));
    CK(cudaMemsetAsync(ctx->pair_seg_count.p, 0, sizeof(int32_t) * (size_t)B.n, st));
    // the fills of this batch follow the fills of the previous batch on this device
    DeviceGate &gate = g_gate[ctx->device & 63];
    std::unique_lock<std::mutex> gate_lock(gate.mu, std::defer_lock);
    // gate_busy: the previous batch on this device is still filling, so this batch's fill cannot start before its
    // own upload has (very likely) landed -- then one fill launch per class beats the chunked fill, whose only
    // purpose is to start early on a partly uploaded batch (two pipelined callers: every batch but the first)
    bool gate_busy = false;
    if (ctx->opt_exclusive) {
        gate_lock.lock();
        if (gate.armed) {
            if (cudaEventQuery(gate.last_fill) == cudaErrorNotReady) {
                gate_busy = true;
                (void)cudaGetLastError();
            }
            CK(cudaStreamWaitEvent(st, gate.last_fill, 0));
        }
    }
    cudaEvent_t ev_setup = ev_new(ctx, true);   // also the start of the fill span
    if (!ev_setup) {
        ctx->err = "cudaEventCreate failed";
        return BA_ERR_CUDA;
    }
    CK(cudaEventRecord(ev_setup, st));
    CK(cudaStreamWaitEvent(ctx->s_trace, ev_setup, 0));
    if (two_streams || B.groups.size() > 1) CK(cudaStreamWaitEvent(ctx->s_fill2, ev_setup, 0));
    std::vector<cudaEvent_t> ev_fill_end((size_t)nw), ev_trace_end((size_t)nw);
    cudaEvent_t ev_trace_start = nullptr;

    for (int64_t wi = 0; wi < nw; wi++) {
        const Plan::Wave &w = plan.waves[(size_t)wi];
        B.waves++;
        cudaStream_t F = (two_streams && (wi & 1)) ? ctx->s_fill2 : st;
        // Chunked fill: while parts of the input are still on their way (ba_align_batch uploads in
        // groups), the fill starts on the pairs that are already here -- one launch per chunk of the
        // work list, alternating between the two fill streams so the tail of one overlaps the head of
        // the next; the traceback follows once, after the last chunk.
        const bool chunked = !overlap && !gate_busy && B.groups.size() > 1 && B.next_group < B.groups.size();
        bool used_alt = false;
        if (!chunked) {
            if ((rc = prep_pairs(ctx, B, w.max_pair)) != BA_OK) return rc;
            if ((rc = sync_prep(ctx, B, F)) != BA_OK) return rc;
            if (F != st && B.next_group > 0) CK(cudaStreamWaitEvent(F, B.groups[B.next_group - 1].checked, 0));
        }
        if (wi >= R) {
            CK(cudaStreamWaitEvent(F, ev_trace_end[(size_t)(wi - R)], 0));  // its region is free again
            // a chunked fill alternates its chunks between F and the second fill stream: that one must not write
            // into the region either before the traceback of wave wi - R has read it (found by tools/fuzz_oracle.py:
            // arena-limited waves + grouped upload + two fill streams gave wrong segments or an illegal access, one
            // round in ~1500)
            if (chunked && ctx->opt_fill_streams >= 2 && F != ctx->s_fill2)
                CK(cudaStreamWaitEvent(ctx->s_fill2, ev_trace_end[(size_t)(wi - R)], 0));
        }
        const int64_t reg = wi % R;
        uint8_t *arena_w = (uint8_t *)ctx->arena.p + reg * region_bytes;
        int32_t *slots_w = (int32_t *)ctx->slots.p + 5 * reg * region_slots;
        const int par = (int)(wi & 1);

        // ---- fill (values + checkpoints)
        int li = 0;
        int64_t bnd_used = 0, tasks_used = 0;
        for (size_t lx = 0; lx < w.launches.size(); lx++) {
            const Plan::Launch &L = w.launches[lx];
            const FillCfg &c = cfgs[(size_t)wi][lx];
            Fill16Params fp;
            fp.enc = B.d_enc;
            fp.stride = B.stride;
            fp.ref_off = B.d_ref_off;
            fp.ref_len = B.d_ref_len;
            fp.qry_off = B.d_qry_off;
            fp.qry_len = B.d_qry_len;
            fp.status = B.d_status;
            fp.acgt_only = B.d_flags;
            fp.flag_mask = flag_mask;
            fp.counter = (int *)ctx->counters.p + 256 * wi + (li % 192);
            fp.arena = arena_w;
            fp.sc = ctx->d_sc;
            fp.long_couples = 0;
            fp.long_err = nullptr;
            fp.max_claims = c.max_claims;
            int32_t *bnd_w = (int32_t *)ctx->bnd.p + par * max_bnd_ints + bnd_used;
            bnd_used += c.bnd_ints;
            if (c.use_long) {
                if (chunked) {   // a wavefront launch needs all its pairs
                    int32_t mp = 0;
                    for (int64_t k = 0; k < L.count; k++) mp = std::max(mp, plan.items[L.first + k].pair);
                    if ((rc = prep_pairs(ctx, B, mp)) != BA_OK) return rc;
                    if ((rc = sync_prep(ctx, B, F)) != BA_OK) return rc;
                }
                WorkItem *h_tasks = (WorkItem *)lease.get(sizeof(WorkItem) * (size_t)c.ntasks);
                if (!h_tasks) {
                    ctx->err = "pinned host allocation failed";
                    return BA_ERR_NOMEM;
                }
                const int W16 = 32 * L.cols;
                for (int ps = 0; ps < c.maxp; ps++)
                    for (int64_t k = 0; k < 2 * c.ncp; k++) {
                        WorkItem t;
                        t.pair = -1;
                        t.cols = L.cols;
                        t.code_off = 0;
                        t.n_ck = 0;
                        t.ck = 0;
                        if (k < L.count) {
                            const WorkItem &it = plan.items[L.first + k];
                            if (ps < (B.h_qry_len[it.pair] + W16 - 1) / W16) t = it;
                        }
                        h_tasks[(int64_t)ps * 2 * c.ncp + k] = t;
                    }
                WorkItem *d_tasks = (WorkItem *)ctx->tasks.p + par * max_tasks + tasks_used;
                tasks_used += c.ntasks;
                if ((rc = upload_words(ctx, d_tasks, h_tasks, sizeof(WorkItem) * (size_t)c.ntasks, F)) != BA_OK)
                    return rc;
                B.launches++;
                // row tags start out invalid
                CK(cudaMemsetAsync(bnd_w, 0, sizeof(int32_t) * (size_t)c.bnd_ints, F));
                fp.items = d_tasks;
                fp.n_items = (int)c.ntasks;
                fp.bnd = bnd_w;
                fp.bnd_stride = c.bnd_stride;
                fp.long_couples = (int)c.ncp;
                fp.long_err = (int *)((char *)ctx->prog.p + 64 * wi);
                BA_LAUNCH_SMEM(c.fn, (unsigned)c.blocks, BA16_WARPS * 32, c.dyn, F, fp);
                CK(cudaMemcpyAsync(h_long_err + wi, fp.long_err, sizeof(int), cudaMemcpyDeviceToHost, F));
                long_tasks.push_back(h_tasks);
                B.out->wavefront_launches++;
                B.launches++;
                B.fill_launches++;
                li++;
                continue;
            }
            fp.bnd = bnd_w;
            fp.bnd_stride = c.bnd_stride;
            int64_t nchunk = 1;
            if (chunked && li + 32 < 192)
                nchunk = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((int64_t)(B.groups.size() - B.next_group) + 1, 8),
                                                                L.count / std::max<int64_t>(ctx->opt_chunk_pairs, 8)));
            int64_t per = (L.count + nchunk - 1) / nchunk;
            per = (per + 7) & ~(int64_t)7;   // whole CTAs of two warps x four pairs
            for (int64_t c0 = 0; c0 < L.count; c0 += per) {
                const int64_t cnt = std::min(per, L.count - c0);
                cudaStream_t FS = F;
                if (chunked) {
                    int32_t mp = 0;
                    for (int64_t k = 0; k < cnt; k++) mp = std::max(mp, plan.items[L.first + c0 + k].pair);
                    if ((rc = prep_pairs(ctx, B, mp)) != BA_OK) return rc;
                    // boundary rows of multi-pass launches are per launch: those stay on one stream
                    if (ctx->opt_fill_streams >= 2 && c.bnd_ints == 0 && ((c0 / per) & 1)) {
                        FS = ctx->s_fill2;
                        used_alt = true;
                    }
                    if (B.next_group > 0 && FS != (B.prep_st ? B.prep_st : B.st))
                        CK(cudaStreamWaitEvent(FS, B.groups[B.next_group - 1].checked, 0));
                }
                const int64_t need = (cnt + 2 * BA16_GROUPS - 1) / (2 * BA16_GROUPS);
                const int64_t blocks = c.max_claims > 0 ? (need + c.max_claims - 1) / c.max_claims
                                                        : std::min(need, std::max<int64_t>(c.maxb, 1));
                fp.items = d_items + L.first + c0;
                fp.n_items = (int)cnt;
                fp.counter = (int *)ctx->counters.p + 256 * wi + (li % 192);
                li++;
                BA_LAUNCH_SMEM(c.fn, (unsigned)blocks, BA16_WARPS * 32, c.dyn, FS, fp);
                B.launches++;
                B.fill_launches++;
            }
        }
        CK(cudaGetLastError());
        if (used_alt) {   // join the second fill stream before the wave's end-of-fill event
            cudaEvent_t ej = ev_new(ctx, false);
            if (!ej) {
                ctx->err = "cudaEventCreate failed";
                return BA_ERR_CUDA;
            }
            CK(cudaEventRecord(ej, ctx->s_fill2));
            CK(cudaStreamWaitEvent(F, ej, 0));
        }
        ev_fill_end[(size_t)wi] = ev_new(ctx, true);
        if (!ev_fill_end[(size_t)wi]) {
            ctx->err = "cudaEventCreate failed";
            return BA_ERR_CUDA;
        }
        CK(cudaEventRecord(ev_fill_end[(size_t)wi], F));
        B.code_bytes += w.code_bytes;

        // ---- tile traceback into slots, then clip, scan, compact (trace stream)
        cudaStream_t T = ctx->s_trace;
        CK(cudaStreamWaitEvent(T, ev_fill_end[(size_t)wi], 0));
        if (wi == 0) {
            ev_trace_start = ev_new(ctx, true);
            if (ev_trace_start) CK(cudaEventRecord(ev_trace_start, T));
        }
        Trace16Params tp;
        tp.enc = B.d_enc;
        tp.stride = B.stride;
        tp.ref_off = B.d_ref_off;
        tp.ref_len = B.d_ref_len;
        tp.qry_off = B.d_qry_off;
        tp.qry_len = B.d_qry_len;
        tp.status = B.d_status;
        tp.acgt_only = B.d_flags;
        tp.flag_mask = flag_mask;
        tp.items = d_items + w.first;
        tp.n_items = (int)w.count;
        tp.arena = arena_w;
        tp.sc = ctx->d_sc;
        tp.seg_count = (int32_t *)ctx->seg_count.p + w.first;
        tp.slot_off = (const int64_t *)ctx->slot_off.p + w.first;
        tp.slot_cap = (const int32_t *)ctx->slot_cap.p + w.first;
        tp.slots = slots_w;
        tp.counter = (int *)ctx->counters.p + 256 * wi + 255;
        tp.row_bias = sym ? ctx->h_sc.gap_open + ctx->h_sc.la[1 * ctx->h_sc.let] : 0;
        tp.val_bias = (sym && kind == KIND_SW) ? -tp.row_bias : 0;
        // one thread per pair when the wave fills the GPU that way, else one warp per pair
        // (window of speculatively rebuilt tiles, see ba_trace16w_kernel), a whole CTA per pair when
        // even that leaves SMs idle; a mixed wave gives warps to its longest pairs (the head of
        // the size-sorted work list) as far as thread slots are free
        const bool cta_trace = ctx->opt_trace_mode == 3 ||
                               (ctx->opt_trace_mode == 1 && w.count <= 2 * (int64_t)ctx->sm_count);
        const bool warp_trace = ctx->opt_trace_mode == 2 ||
                                (ctx->opt_trace_mode == 1 &&
                                 w.count * 32 <= (int64_t)ctx->sm_count * BA_TRACE_MINB * BA_TRACE_THREADS * 2);
        // thread-per-pair CTAs: 64 threads beside a running fill (they fit the slot of one fill CTA)
        int wave_mc = BA_MAX_COLS;
        for (auto &L : w.launches) wave_mc = std::max(wave_mc, L.cols);
        const int tcta = wave_mc > BA_MAX_COLS ? BA_TRACE_THREADS
                         : (ctx->opt_trace_cta == 64 || (ctx->opt_trace_cta == 0 && overlap && wi + 1 < nw)) ? 64
                                                                                                             : BA_TRACE_THREADS;
        const size_t tsm128 = 0, tsm = 0;   // the traceback kernels use static shared memory only
        auto trace_launch = [&](Trace16Fn fn, unsigned grid, int threads, size_t smem, const Trace16Params &prm) -> int {
            int occ_unused = 0;
            int r = cached_occupancy(ctx, fn, threads, occ_unused, smem);   // opts the kernel in to > 48 KB once
            if (r != BA_OK) return r;
            BA_LAUNCH_SMEM(fn, grid, threads, smem, T, prm);
            return BA_OK;
        };
        if (cta_trace) {
            if ((rc = trace_launch(pick_trace16t<BA_TRACE_THREADS>(kind, lin, wave_mc), (unsigned)w.count, BA_TRACE_THREADS, tsm128, tp)) != BA_OK)
                return rc;
            B.out->warp_trace_launches++;
        } else if (warp_trace) {
            const unsigned tb = (unsigned)((w.count + BA_TRACE_THREADS / 32 - 1) / (BA_TRACE_THREADS / 32));
            if ((rc = trace_launch(pick_trace16t<32>(kind, lin, wave_mc), tb, BA_TRACE_THREADS, tsm128, tp)) != BA_OK) return rc;
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
                if ((rc = trace_launch(pick_trace16t<32>(kind, lin, wave_mc), tb, BA_TRACE_THREADS, tsm128, th)) != BA_OK) return rc;
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
                if (ctx->opt_trace_mode != 5 || wave_mc > BA_MAX_COLS) {   // (the persistent variant has 8-column tiles only)
                    // static assignment: thread k of the grid walks pair k
                    const int64_t tblocks = (w.count - head + tcta - 1) / tcta;
                    if ((rc = trace_launch(pick_trace16(kind, lin, tcta, wave_mc), (unsigned)tblocks, tcta, tsm, tt)) != BA_OK) return rc;
                } else {
                    // persistent grid, pairs claimed from a counter (largest first).  Measured on B200:
                    // +5 % on configs 1 and 3, -9 % on config 2 against the static kernel, so it is
                    // opt-in (trace_mode 5) until the state machine's overhead is trimmed
                    int64_t tblocks = (w.count - head + BA_TRACE_THREADS - 1) / BA_TRACE_THREADS;
                    tblocks = std::min<int64_t>(tblocks, (int64_t)ctx->sm_count * BA_TRACE_MINB);
                    if ((rc = trace_launch(pick_trace16p(kind, lin), (unsigned)tblocks, BA_TRACE_THREADS, tsm128, tt)) != BA_OK)
                        return rc;
                }
            }
        }
        const unsigned cblocks = (unsigned)((w.count + 255) / 256);
        long long *d_meta = (long long *)ctx->meta.p;
        BA_LAUNCH(ba_clip_counts_kernel, cblocks, 256, T, (const int32_t *)ctx->seg_count.p + w.first,
                  (const int32_t *)ctx->slot_cap.p + w.first, (const WorkItem *)d_items + w.first,
                  (int32_t *)ctx->seg_clip.p + w.first, d_meta, (int32_t *)ctx->ovf.p, (int)w.count);
        {
            size_t tb = ctx->scan_tmp.cap;
            if ((rc = exclusive_scan(ctx, ctx->scan_tmp.p, tb, (const int32_t *)ctx->seg_clip.p + w.first,
                                     (int64_t *)ctx->seg_base.p + w.first, (int)w.count, T)) != BA_OK)
                return rc;
        }
        BA_LAUNCH(ba_compact_kernel, cblocks, 256, T, (const int32_t *)slots_w, (const int64_t *)ctx->slot_off.p + w.first,
                  (const int32_t *)ctx->slot_cap.p + w.first, (const int32_t *)ctx->seg_count.p + w.first,
                  (const int64_t *)ctx->seg_base.p + w.first, (const WorkItem *)d_items + w.first,
                  (const long long *)d_meta, (int32_t *)ctx->segs.p, (int64_t *)ctx->pair_seg_start.p,
                  (int32_t *)ctx->pair_seg_count.p, (int)w.count);
        BA_LAUNCH(ba_wave_total_kernel, 1, 32, T, (const int64_t *)ctx->seg_base.p + w.first,
                  (const int32_t *)ctx->seg_clip.p + w.first, (int)w.count, d_meta, (int)wi);
        B.launches += 6;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_meta + BA_META_HEAD + BA_META_WAVE * wi, d_meta + BA_META_HEAD + BA_META_WAVE * wi,
                           sizeof(long long) * BA_META_WAVE, cudaMemcpyDeviceToHost, T));
        ev_trace_end[(size_t)wi] = ev_new(ctx, true);
        if (!ev_trace_end[(size_t)wi]) {
            ctx->err = "cudaEventCreate failed";
            return BA_ERR_CUDA;
        }
        CK(cudaEventRecord(ev_trace_end[(size_t)wi], T));
    }
    // every fill of this batch is enqueued: the next batch on this device may queue its fills behind them
    if (ctx->opt_exclusive) {
        if (!gate.last_fill) CK(cudaEventCreateWithFlags(&gate.last_fill, cudaEventDisableTiming));
        if (two_streams && nw >= 2)   // the last fill on the second stream (largest odd wave)
            CK(cudaStreamWaitEvent(st, ev_fill_end[(size_t)(((nw - 1) & 1) ? nw - 1 : nw - 2)], 0));
        CK(cudaEventRecord(gate.last_fill, st));
        gate.armed = true;
        gate_lock.unlock();
    }
    BA_TSEC("P16 all waves enqueued", tp0);

    // ---- follow the waves on the host: as soon as a wave is compacted its segments travel
    for (int64_t wi = 0; wi < nw; wi++) {
        CK(cudaEventSynchronize(ev_trace_end[(size_t)wi]));
        const long long *rec = h_meta + BA_META_HEAD + BA_META_WAVE * wi;
        const int64_t base = rec[0], wsegs = rec[1];
        int64_t want = base + wsegs;
        if (wi == 0 && nw > 1) {
            // size the block for the whole batch from the first wave's segments per pair
            const double per = (double)wsegs / (double)std::max<int64_t>(plan.waves[0].count, 1);
            want = std::max<int64_t>(want, (int64_t)(per * 1.1 * (double)nids) + 4096);
            want = std::min<int64_t>(want, total_slots);
        }
        if (!B.defer_segs) {
            if ((rc = segs_reserve(ctx, B, want, base)) != BA_OK) return rc;
            if (!B.ev_d2h0) {
                B.ev_d2h0 = ev_new(ctx, true);
                if (B.ev_d2h0) CK(cudaEventRecord(B.ev_d2h0, ctx->s_d2h));
            }
            if (wsegs > 0)
                CK(cudaMemcpyAsync(B.h_segs + 5 * base, (const int32_t *)ctx->segs.p + 5 * base,
                                   sizeof(int32_t) * 5 * (size_t)wsegs, cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
        B.seg_total = base + wsegs;
        B.dev_segs = B.seg_total;
    }
    BA_TSEC("P16 waves done", tp0);
    const bool long_tasks_empty = long_tasks.empty();
    for (WorkItem *t : long_tasks) lease.drop(t);
    // spans on the device (the kernels of different waves overlap)
    {
        float ms = 0.f, ms2 = 0.f;
        CK(cudaEventElapsedTime(&ms, ev_setup, ev_fill_end[(size_t)(nw - 1)]));
        if (nw >= 2) CK(cudaEventElapsedTime(&ms2, ev_setup, ev_fill_end[(size_t)(nw - 2)]));
        B.ms_fill += std::max(ms, ms2);
        if (ev_trace_start) {
            CK(cudaEventElapsedTime(&ms, ev_trace_start, ev_trace_end[(size_t)(nw - 1)]));
            B.ms_trace += ms;
        }
    }
    if (ctx->opt_test_long_timeout && !long_tasks_empty) h_long_err[0] = 1;   // test hook: exercise the retry path
    for (int64_t wi = 0; wi < nw; wi++)
        if (h_long_err[wi] != 0) {
            ctx->err = "fill16 wavefront: a pass waited too long for its left neighbour";
            return BA_ERR_LONG_TIMEOUT;
        }
    // pairs for the exact path (slot overflow, letters outside the kernel's alphabet class)
    const int64_t novf = h_meta[BA_META_HEAD + BA_META_WAVE * (nw - 1) + 2];
    // per-pair results in pair order straight from the device
    CK(cudaMemcpyAsync(B.out->seg_start, ctx->pair_seg_start.p, sizeof(int64_t) * (size_t)B.n, cudaMemcpyDeviceToHost,
                       ctx->s_d2h));
    CK(cudaMemcpyAsync(B.out->seg_count, ctx->pair_seg_count.p, sizeof(int32_t) * (size_t)B.n, cudaMemcpyDeviceToHost,
                       ctx->s_d2h));
    int32_t *h_ovf = nullptr;
    if (novf > 0) {
        h_ovf = (int32_t *)lease.get(sizeof(int32_t) * (size_t)novf);
        if (!h_ovf) {
            ctx->err = "pinned host allocation failed";
            return BA_ERR_NOMEM;
        }
        CK(cudaMemcpyAsync(h_ovf, ctx->ovf.p, sizeof(int32_t) * (size_t)novf, cudaMemcpyDeviceToHost, ctx->s_d2h));
    }
    CK(cudaStreamSynchronize(ctx->s_d2h));
    if (novf > 0) {
        std::sort(h_ovf, h_ovf + novf);
        overflow.insert(overflow.end(), h_ovf, h_ovf + novf);
        lease.drop(h_ovf);
    }
    lease.drop(h_meta);
    lease.drop(h_long_err);
    if (!use_cache) {
        lease.drop(plan.items);
        lease.drop(plan.slot_off);
        lease.drop(plan.slot_cap);
    }
    BA_TSEC("P16 done", tp0);
    return BA_OK;
}

int run_device_once(ba_ctx *ctx, Batch &B, int64_t letters_bytes)
{
    const int kind = ctx->h_sc.kind, affine = ctx->h_sc.affine;
    (void)kind;
    (void)affine;
    const double t_call = now_s();
    ba_result *out = B.out;
    const int64_t n = B.n;
    cudaStream_t st = B.st;
    int rc;
    ctx->ev_timed_next = ctx->ev_plain_next = 0;
    B.next_group = 0;
    B.seg_total = 0;
    B.dev_segs = 0;
    B.waves = 0;
    B.ev_d2h0 = nullptr;

    if ((rc = ensure(ctx, ctx->enc, (size_t)letters_bytes + 16)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->status, sizeof(int32_t) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->err_pos, sizeof(int64_t) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->start, sizeof(PairStart) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->flags, (size_t)n + 16)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->pair_seg_start, sizeof(int64_t) * (size_t)n)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->pair_seg_count, sizeof(int32_t) * (size_t)n)) != BA_OK) return rc;
    B.d_enc = (uint8_t *)ctx->enc.p;
    B.d_status = (int32_t *)ctx->status.p;
    B.d_start = (PairStart *)ctx->start.p;
    B.d_err_pos = (int64_t *)ctx->err_pos.p;
    const Fast16 f16 = fast16_scoring(ctx->h_sc);
    const bool fast16 = ctx->opt_force_path != 1 && f16.ok;
    B.d_flags = fast16 ? (uint8_t *)ctx->flags.p : nullptr;
    CK(cudaEventRecord(ctx->ev[1], st));

    // ---- split the batch between the two paths (cached for batches of the same shape)
    std::vector<int32_t> slow;
    const int32_t *h_ref_len = B.h_ref_len, *h_qry_len = B.h_qry_len;
    bool host_arrays_final = false;   // out->seg_start / seg_count already hold the device's pair-order results
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
        if (!pc.fast.empty()) {
            if ((rc = run_path16(ctx, B, pc.fast.data(), (int64_t)pc.fast.size(), f16.gsub, slow,
                                 ctx->opt_plan_cache != 0, pc.cells)) != BA_OK)
                return rc;
            host_arrays_final = true;
        }
        out->pairs_fast16 = (int64_t)pc.fast.size() - (int64_t)(slow.size() - nslow0);
        if ((rc = run_path32(ctx, B, slow.data(), (int64_t)slow.size())) != BA_OK) return rc;
    } else {
        int64_t cells = 0;
        for (int64_t p = 0; p < n; p++) cells += (int64_t)h_ref_len[p] * (int64_t)h_qry_len[p];
        out->cells = cells;
        if ((rc = run_path32(ctx, B, nullptr, n)) != BA_OK) return rc;
    }
    (void)host_arrays_final;
    if ((rc = prep_pairs(ctx, B, n - 1)) != BA_OK) return rc;   // groups no wave asked for (nothing to align)
    if ((rc = sync_prep(ctx, B, st)) != BA_OK) return rc;

    BA_TSEC("T paths done", t_call);
    // ---- per-pair status (the traceback may have changed it) and error positions
    if (!B.ev_d2h0) {
        B.ev_d2h0 = ev_new(ctx, true);
        if (B.ev_d2h0) CK(cudaEventRecord(B.ev_d2h0, st));
    }
    CK(cudaMemcpyAsync(out->status, B.d_status, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->err_pos, B.d_err_pos, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[3], st));
    CK(cudaStreamSynchronize(st));
    {
        float ms = 0.f, all = 0.f;
        if (B.ev_d2h0) CK(cudaEventElapsedTime(&ms, B.ev_d2h0, ctx->ev[3]));
        CK(cudaEventElapsedTime(&all, ctx->ev[1], ctx->ev[3]));
        out->ms_d2h = ms;          // span from the first result copy to the last (overlaps the kernels of later waves)
        out->ms_kernels = all;     // whole device span of the call
    }
    for (int64_t p = 0; p < n; p++)
        if (out->status[p] != 0) {
            out->seg_start[p] = 0;
            out->seg_count[p] = 0;
        }
    out->segs = B.h_segs;
    if (B.h_segs) B.lease->keep(B.h_segs);
    B.h_segs = nullptr;
    out->n_segs = B.seg_total;
    ctx->deferred_dev_segs = B.defer_segs ? B.dev_segs : 0;
    out->gpu_launches = B.launches;
    out->fill_launches = B.fill_launches;
    out->code_bytes = B.code_bytes;
    out->waves = B.waves;
    out->ms_fill = B.ms_fill;
    out->ms_trace = B.ms_trace;
    out->ms_host = (float)(B.host_s * 1e3);
    BA_TSEC("T call total", t_call);
    return BA_OK;
}

// Drain every stream of the context (error paths: nothing may still be writing into pooled memory).
void drain(ba_ctx *ctx, cudaStream_t extra)
{
    cudaStream_t all[] = {ctx->stream, ctx->s_fill2, ctx->s_trace, ctx->s_h2d, ctx->s_d2h, ctx->s_prep, extra};
    for (cudaStream_t s : all)
        if (s) cudaStreamSynchronize(s);
    (void)cudaGetLastError();
}

int run_device(ba_ctx *ctx, Batch &B, int64_t letters_bytes)
{
    ba_result *out = B.out;
    const int64_t n = B.n;
    Lease lease(ctx);
    B.lease = &lease;
    // ---- result arrays (pinned, pooled)
    out->n = n;
    out->status = (int32_t *)lease.get(sizeof(int32_t) * (size_t)n);
    out->err_pos = (int64_t *)lease.get(sizeof(int64_t) * (size_t)n);
    out->seg_start = (int64_t *)lease.get(sizeof(int64_t) * (size_t)n);
    out->seg_count = (int32_t *)lease.get(sizeof(int32_t) * (size_t)n);
    out->segs = nullptr;
    out->n_segs = 0;
    if (!out->status || !out->err_pos || !out->seg_start || !out->seg_count) {
        out->status = nullptr;
        out->err_pos = nullptr;
        out->seg_start = nullptr;
        out->seg_count = nullptr;
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    int rc = BA_OK;
    if (n > 0) {
        memset(out->seg_start, 0, sizeof(int64_t) * (size_t)n);
        memset(out->seg_count, 0, sizeof(int32_t) * (size_t)n);
        rc = run_device_once(ctx, B, letters_bytes);
        if (rc == BA_ERR_LONG_TIMEOUT) {
            // a wavefront pass gave up waiting (never observed; the spin is bounded so that a scheduling
            // anomaly cannot hang the GPU): run the batch again with the pass-by-pass kernels
            drain(ctx, B.st);
            const int64_t keep = ctx->opt_long_mode;
            ctx->opt_long_mode = 0;
            plan_cache_drop(ctx);
            const int64_t wl = out->wavefront_launches;
            B.launches = B.fill_launches = B.code_bytes = 0;
            B.ms_fill = B.ms_trace = 0.f;
            if (B.h_segs) lease.drop(B.h_segs);
            B.h_segs = nullptr;
            B.h_segs_cap = 0;
            rc = run_device_once(ctx, B, letters_bytes);
            out->wavefront_launches = wl;
            ctx->opt_long_mode = keep;
            plan_cache_drop(ctx);
            if (rc == BA_ERR_LONG_TIMEOUT) rc = BA_ERR_CUDA;
        }
    }
    if (rc != BA_OK) {
        drain(ctx, B.st);
        plan_cache_drop(ctx);   // its device copy / waves may be half-built
        out->status = nullptr;
        out->err_pos = nullptr;
        out->seg_start = nullptr;
        out->seg_count = nullptr;
        out->segs = nullptr;
        return rc;              // the lease returns every pinned block
    }
    lease.keep(out->status);
    lease.keep(out->err_pos);
    lease.keep(out->seg_start);
    lease.keep(out->seg_count);
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
    int prio_lo = 0, prio_hi = 0;
    e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e != cudaSuccess) return fail(e, "cudaDeviceGetStreamPriorityRange");
    e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_lo);
    if (e != cudaSuccess) return fail(e, "cudaStreamCreate");
    e = cudaStreamCreateWithPriority(&ctx->s_fill2, cudaStreamNonBlocking, prio_lo);
    if (e != cudaSuccess) return fail(e, "cudaStreamCreate");
    e = cudaStreamCreateWithPriority(&ctx->s_trace, cudaStreamNonBlocking, prio_hi);  // CTAs of the traceback go first
    if (e != cudaSuccess) return fail(e, "cudaStreamCreate");
    e = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(e, "cudaStreamCreate");
    e = cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(e, "cudaStreamCreate");
    e = cudaStreamCreateWithPriority(&ctx->s_prep, cudaStreamNonBlocking, prio_hi);   // short kernels between fill CTAs
    if (e != cudaSuccess) return fail(e, "cudaStreamCreate");
    for (int k = 0; k < 8; k++) {
        e = cudaEventCreate(&ctx->ev[k]);
        if (e != cudaSuccess) return fail(e, "cudaEventCreate");
    }
    e = cudaMalloc((void **)&ctx->d_sc_first, sizeof(DeviceScoring));
    if (e != cudaSuccess) return fail(e, "cudaMalloc");
    ctx->d_sc = ctx->d_sc_first;
    *out = ctx;
    return BA_OK;
}

void ba_destroy(ba_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStream_t streams[] = {ctx->stream, ctx->s_fill2, ctx->s_trace, ctx->s_h2d, ctx->s_d2h, ctx->s_prep};
    for (cudaStream_t s : streams)
        if (s) cudaStreamSynchronize(s);
    plan_cache_drop(ctx);
    DevBuf *bufs[] = {&ctx->letters, &ctx->enc,  &ctx->ref_off,   &ctx->ref_len,  &ctx->qry_off, &ctx->qry_len,
                      &ctx->status,  &ctx->err_pos, &ctx->start,  &ctx->items,    &ctx->aux_off, &ctx->aux,
                      &ctx->bnd,     &ctx->seg_count, &ctx->seg_base, &ctx->segs, &ctx->arena,   &ctx->scan_tmp,
                      &ctx->counters, &ctx->flags,  &ctx->slot_off,  &ctx->slot_cap, &ctx->slots, &ctx->seg_clip,
                      &ctx->tasks,   &ctx->prog,     &ctx->fmt_len, &ctx->fmt_off, &ctx->fmt_rows,
                      &ctx->fmt_segs, &ctx->fmt_seg_start, &ctx->fmt_seg_count, &ctx->pair_seg_start,
                      &ctx->pair_seg_count, &ctx->meta, &ctx->ovf, &ctx->segs32, &ctx->pals_meta,
                      &ctx->pals_traps, &ctx->pals_vec, &ctx->pals_stack, &ctx->pals_hits};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (auto &b : ctx->pinned) cudaFreeHost(b.p);
    for (size_t k = 1; k < ctx->sc_cache.size(); k++)
        if (ctx->sc_cache[k].d) cudaFree(ctx->sc_cache[k].d);
    if (ctx->d_sc_first) cudaFree(ctx->d_sc_first);
    for (int k = 0; k < 8; k++)
        if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    for (cudaEvent_t e : ctx->ev_timed) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_plain) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_upload) cudaEventDestroy(e);
    for (cudaStream_t s : streams)
        if (s) cudaStreamDestroy(s);
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
    // overlap scheduling of the 16-bit path (run_path16)
    if (!strcmp(key, "waves")) {           // waves per batch: 0 = automatic, 1 = no overlap, k = about k waves
        ctx->opt_waves = value;
        return BA_OK;
    }
    if (!strcmp(key, "fill_streams")) {    // 1: all fills on one stream; 2 (default): alternate two
        ctx->opt_fill_streams = value;
        return BA_OK;
    }
    if (!strcmp(key, "fill_claims")) {     // claims per fill warp before its CTA retires; 0 = persistent grid
        ctx->opt_fill_claims = value;
        return BA_OK;
    }
    if (!strcmp(key, "fill_cap")) {        // resident fill CTAs per SM while waves overlap; 0 = all that fit
        ctx->opt_fill_cap = value;
        return BA_OK;
    }
    if (!strcmp(key, "trace_cta")) {       // threads per thread-per-pair traceback CTA: 0 automatic, 64, 128
        ctx->opt_trace_cta = value;
        return BA_OK;
    }
    if (!strcmp(key, "upload_groups")) {   // pieces the letters upload is cut into: 0 automatic, 1 = one piece
        ctx->opt_upload_groups = value;
        return BA_OK;
    }
    if (!strcmp(key, "chunk_pairs")) {     // fewest pairs a fill chunk of a still-uploading batch gets (0 = default)
        ctx->opt_chunk_pairs = value > 0 ? value : 8192;
        return BA_OK;
    }
    if (!strcmp(key, "wide10")) {          // 0: no 10-column single-pass class for queries of 257..320 letters
        ctx->opt_wide10 = value;
        plan_cache_drop(ctx);
        return BA_OK;
    }
    if (!strcmp(key, "test_long_timeout")) {   // test hook: report the first wavefront launch as timed out
        ctx->opt_test_long_timeout = value;
        return BA_OK;
    }
    if (!strcmp(key, "pinned_budget")) {   // bytes of idle pinned host memory a context keeps for reuse
        ctx->pinned_budget = value;
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
    CK(cudaSetDevice(ctx->device));
    // device copies of the last few scorings stay resident (callers that alternate between aligners or
    // named tables -- ba_set_scoring_named -- switch by pointer instead of uploading again)
    static std::atomic<uint64_t> clock_{1};
    ba_ctx::ScoringSlot *slot = nullptr;
    for (auto &e : ctx->sc_cache)
        if (memcmp(&s, &e.h, sizeof s) == 0) slot = &e;
    if (!slot) {
        if (ctx->sc_cache.size() < 8) {
            ba_ctx::ScoringSlot e;
            e.d = nullptr;
            if (ctx->sc_cache.empty())
                e.d = ctx->d_sc_first;   // the buffer ba_create allocated
            else
                CK(cudaMalloc((void **)&e.d, sizeof(DeviceScoring)));
            ctx->sc_cache.push_back(e);
            slot = &ctx->sc_cache.back();
        } else {
            slot = &ctx->sc_cache[0];
            for (auto &e : ctx->sc_cache)
                if (e.stamp < slot->stamp) slot = &e;
        }
        slot->h = s;
        // no kernel of this context is in flight between calls, so the least recently used copy may be rewritten
        CK(cudaMemcpyAsync(slot->d, &slot->h, sizeof s, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->scoring_uploads++;
    }
    slot->stamp = clock_++;
    ctx->d_sc = slot->d;
    ctx->h_sc = s;
    ctx->have_scoring = true;
    ctx->scoring_epoch++;  // cached plans depend on the scoring (eligibility, kind)
    ctx->occ_cache.clear(); // dynamic shared memory of the GSUB kernels depends on the matrix size
    return BA_OK;
}

// ---- align/matrix as a library: the 78 tables of align/matrix/matrices.go by name
int ba_matrix_count(void) { return ba_matrix_table_n; }

const char *ba_matrix_name(int k) { return (k >= 0 && k < ba_matrix_table_n) ? ba_matrix_table[k].name : nullptr; }

int ba_matrix_get(const char *name, int *let, const char **letters, const int16_t **table)
{
    if (!name) return BA_ERR_ARG;
    for (int k = 0; k < ba_matrix_table_n; k++)
        if (!strcmp(ba_matrix_table[k].name, name)) {
            if (let) *let = ba_matrix_table[k].let;
            if (letters) *letters = ba_matrix_table[k].letters;
            if (table) *table = (const int16_t *)ba_matrix_table[k].data;
            return BA_OK;
        }
    return BA_ERR_ARG;
}

int ba_set_scoring_named(ba_ctx *ctx, int kind, int affine, const char *name, int64_t gap, int alpha_len,
                         int64_t gap_open, const int32_t *index)
{
    if (!ctx) return BA_ERR_ARG;
    int let = 0;
    const int16_t *t = nullptr;
    if (ba_matrix_get(name, &let, nullptr, &t) != BA_OK) {
        ctx->err = std::string("unknown scoring matrix: ") + (name ? name : "(null)");
        return BA_ERR_ARG;
    }
    std::vector<int64_t> la((size_t)let * let);
    for (int k = 0; k < let * let; k++) la[(size_t)k] = t[k];
    // the published tables carry no gap penalty (matrices.go:10-13): write `gap` into the gap letter's
    // row and column, [0][0] stays
    for (int k = 1; k < let; k++) {
        la[(size_t)k] = gap;
        la[(size_t)k * let] = gap;
    }
    return ba_set_scoring(ctx, kind, affine, la.data(), let, alpha_len, gap_open, index);
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
    ctx->last_n = -1;  // the context's own letters buffer no longer mirrors the last batch (ba_format_batch)
    Batch B;
    B.d_letters = d_letters;
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
    UploadGroup g;   // everything is resident already
    g.p_lo = 0;
    g.p_hi = n;
    g.b_lo = 0;
    g.b_hi = letters_bytes;
    B.groups.push_back(g);
    rc = run_device(ctx, B, letters_bytes);
    out->ms_h2d = 0.f;
    if (rc != BA_OK) memset(out, 0, sizeof *out);
    return rc;
}

// A share of a larger batch whose letters are scattered over the caller's buffer (multi-device
// batches): pair k of the call is pair ids[k] of the caller.  The call's own layout is "reference then
// query of every pair, in call order"; each upload group is gathered into the pinned staging buffer
// right before its copy is enqueued, so the gather of group g+1 runs under the DMA of group g.
struct GatherSrc {
    const uint8_t *letters;
    const int64_t *ref_off, *qry_off;   // the caller's arrays (indexed by ids[k])
    const int64_t *ids;
    uint8_t *stage;                     // pinned, the call's letters_bytes long
};

static int align_host(ba_ctx *ctx, const uint8_t *letters, int stride, const int64_t *ref_off,
                      const int32_t *ref_len, const int64_t *qry_off, const int32_t *qry_len, int64_t n,
                      ba_result *out, const GatherSrc *gs, bool defer_segs);

int ba_align_batch(ba_ctx *ctx, const uint8_t *letters, int stride, const int64_t *ref_off,
                   const int32_t *ref_len, const int64_t *qry_off, const int32_t *qry_len, int64_t n,
                   ba_result *out)
{
    return align_host(ctx, letters, stride, ref_off, ref_len, qry_off, qry_len, n, out, nullptr, false);
}

static int align_host(ba_ctx *ctx, const uint8_t *letters, int stride, const int64_t *ref_off,
                      const int32_t *ref_len, const int64_t *qry_off, const int32_t *qry_len, int64_t n,
                      ba_result *out, const GatherSrc *gs, bool defer_segs)
{
    if (gs) letters = gs->stage;
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
    // ---- upload groups: pairs [p_lo, p_hi) need only letters below b_hi (a running maximum of the
    // sequence ends, so any layout is legal; batches packed in pair order split evenly).  The first
    // waves start on the first groups while the rest of the letters are still on their way.
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
    Batch B;
    {
        int64_t ng = ctx->opt_upload_groups;
        if (ng <= 0) ng = std::min<int64_t>(16, letters_bytes / (8 << 20));  // pieces of >= 8 MB
        if (ng < 1) ng = 1;
        const int64_t piece = (letters_bytes + ng - 1) / ng;
        int64_t hi = 0, p_lo = 0, b_lo = 0;
        for (int64_t p = 0; p < n; p++) {
            int64_t e1 = ref_len[p] ? ref_off[p] + ((int64_t)ref_len[p] - 1) * stride + 1 : 0;
            int64_t e2 = qry_len[p] ? qry_off[p] + ((int64_t)qry_len[p] - 1) * stride + 1 : 0;
            hi = std::max(hi, std::max(e1, e2));
            const bool last = (p + 1 == n);
            if (last || hi - b_lo >= piece) {
                UploadGroup g;
                g.p_lo = p_lo;
                g.p_hi = p + 1;
                g.b_lo = b_lo;
                // groups meet on 16-byte boundaries (vectorised encode); the last one ends the buffer
                g.b_hi = last ? letters_bytes : std::min(letters_bytes, (hi + 15) & ~(int64_t)15);
                B.groups.push_back(g);
                p_lo = p + 1;
                b_lo = g.b_hi;
            }
        }
    }
    cudaStream_t st = ctx->stream, up = ctx->s_h2d;
    ctx->last_n = -1;
    if ((rc = ensure(ctx, ctx->letters, (size_t)letters_bytes + 16)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->ref_off, sizeof(int64_t) * (size_t)(n + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->qry_off, sizeof(int64_t) * (size_t)(n + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->ref_len, sizeof(int32_t) * (size_t)(n + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->qry_len, sizeof(int32_t) * (size_t)(n + 1))) != BA_OK) return rc;
    CK(cudaEventRecord(ctx->ev[0], up));
    // The offset/length arrays are the caller's pageable memory: a copy straight from them would
    // block inside the driver behind the letters copy.  Stage them in pinned memory and send them
    // first; the letters buffer may be pinned (ba_alloc_host) or not.
    Lease stage_lease(ctx);
    ctx->ev_timed_next = ctx->ev_plain_next = 0;
    if (n > 0) {
        const size_t b8 = sizeof(int64_t) * (size_t)n, b4 = sizeof(int32_t) * (size_t)n;
        uint8_t *stage = (uint8_t *)stage_lease.get(2 * b8 + 2 * b4);
        if (!stage) {
            ctx->err = "pinned host allocation failed";
            return BA_ERR_NOMEM;
        }
        memcpy(stage, ref_off, b8);
        memcpy(stage + b8, qry_off, b8);
        memcpy(stage + 2 * b8, ref_len, b4);
        memcpy(stage + 2 * b8 + b4, qry_len, b4);
        CK(cudaMemcpyAsync(ctx->ref_off.p, stage, b8, cudaMemcpyHostToDevice, up));
        CK(cudaMemcpyAsync(ctx->qry_off.p, stage + b8, b8, cudaMemcpyHostToDevice, up));
        CK(cudaMemcpyAsync(ctx->ref_len.p, stage + 2 * b8, b4, cudaMemcpyHostToDevice, up));
        CK(cudaMemcpyAsync(ctx->qry_len.p, stage + 2 * b8 + b4, b4, cudaMemcpyHostToDevice, up));
    }
    // the letters, group by group, each in pieces of 4 MB (small copies of other contexts -- work lists,
    // per-wave records -- share the copy engines and must not queue behind one 100+ MB transfer)
    // pooled events: run_device resets the pool cursor, so these come from a pool of their own
    while (ctx->ev_upload.size() < B.groups.size()) {
        cudaEvent_t e = nullptr;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_upload.push_back(e);
    }
    std::vector<cudaEvent_t> ready(ctx->ev_upload.begin(), ctx->ev_upload.begin() + B.groups.size());
    std::atomic<int64_t> enqueued{0};
    cudaError_t feed_err = cudaSuccess;
    auto feed = [&, letters]() {
        int64_t gathered = 0;   // pairs whose letters are in the staging buffer (gs)
        cudaError_t e = cudaSuccess;
        for (size_t gi = 0; gi < B.groups.size(); gi++) {
            const UploadGroup &g = B.groups[gi];
            if (gs) {
                // every pair that starts below the end of this piece
                while (gathered < n && ref_off[gathered] < g.b_hi) {
                    const int64_t p = gs->ids[gathered];
                    const size_t rb = (size_t)ref_len[gathered] * stride, qb = (size_t)qry_len[gathered] * stride;
                    if (rb) memcpy(gs->stage + ref_off[gathered], gs->letters + gs->ref_off[p], rb);
                    if (qb) memcpy(gs->stage + qry_off[gathered], gs->letters + gs->qry_off[p], qb);
                    gathered++;
                }
            }
            for (int64_t o = g.b_lo; o < g.b_hi && e == cudaSuccess; o += (4 << 20)) {
                const size_t len = (size_t)std::min<int64_t>(4 << 20, g.b_hi - o);
                e = cudaMemcpyAsync((uint8_t *)ctx->letters.p + o, letters + o, len, cudaMemcpyHostToDevice, up);
            }
            if (e == cudaSuccess) e = cudaEventRecord(ready[gi], up);
            enqueued.store((int64_t)gi + 1, std::memory_order_release);   // (also on failure: nobody may wait forever)
        }
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[7], up);
        feed_err = e;
    };
    // A scattered share in several groups: the gather is host work (a few GB/s), so it runs on a thread of
    // its own beside the launches and the kernels of the groups that are already on the device.
    std::thread feeder;
    struct Join {
        std::thread &t;
        ~Join() { if (t.joinable()) t.join(); }
    } join_feeder{feeder};
    if (gs && B.groups.size() > 1) {
        B.groups_enqueued = &enqueued;
        const int dev = ctx->device;
        feeder = std::thread([&feed, dev]() {
            cudaSetDevice(dev);
            feed();
        });
    } else {
        feed();
        if (feed_err != cudaSuccess) {
            ctx->err = std::string("upload: ") + cudaGetErrorString(feed_err);
            return BA_ERR_CUDA;
        }
    }
    for (size_t gi = 0; gi < B.groups.size(); gi++) B.groups[gi].ready = ready[gi];
    B.d_letters = (const uint8_t *)ctx->letters.p;
    B.stride = stride;
    B.d_ref_off = (const int64_t *)ctx->ref_off.p;
    B.d_qry_off = (const int64_t *)ctx->qry_off.p;
    B.d_ref_len = (const int32_t *)ctx->ref_len.p;
    B.d_qry_len = (const int32_t *)ctx->qry_len.p;
    B.h_ref_len = ref_len;
    B.h_qry_len = qry_len;
    B.n = n;
    B.st = st;
    B.prep_st = B.groups.size() > 1 ? ctx->s_prep : nullptr;
    B.out = out;
    B.defer_segs = defer_segs;
    rc = run_device(ctx, B, letters_bytes);
    if (feeder.joinable()) {
        feeder.join();
        if (rc == BA_OK && feed_err != cudaSuccess) {
            ctx->err = std::string("upload: ") + cudaGetErrorString(feed_err);
            rc = BA_ERR_CUDA;
        }
    }
    if (rc != BA_OK) {
        drain(ctx, nullptr);   // the staged copies may still be in flight
        memset(out, 0, sizeof *out);
        return rc;
    }
    CK(cudaStreamSynchronize(up));  // (already drained: every group was waited for)
    if (n > 0) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[7]));
        out->ms_h2d = ms;   // span of the upload stream; it runs beside the kernels of the first waves
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

#ifdef BA_TRACE_STATS
extern "C" int ba_debug_trace_stats(unsigned long long *out8, int reset)
{
    if (cudaMemcpyFromSymbol(out8, ba_trace_stats, sizeof(unsigned long long) * 8) != cudaSuccess) return BA_ERR_CUDA;
    if (reset) {
        unsigned long long z[8] = {0};
        if (cudaMemcpyToSymbol(ba_trace_stats, z, sizeof z) != cudaSuccess) return BA_ERR_CUDA;
    }
    return BA_OK;
}
#endif
#include "ba_seqio.inl"
#include "ba_multi.inl"
#include "ba_pals.inl"
