// ba_multi.inl -- ONE batch over several GPUs of one box, under the C ABI (include/biogo_align.h).
//
// Pairs never communicate, so there is no collective and no NCCL: the batch is dealt to the
// devices by length-bucketed work lists, every device runs its share through its own ba_ctx from its own
// host thread, and the host gathers the results by pair id.
//   * dealing: pairs are bucketed by cells (quarter-octave classes, counting sort, largest first) and
//     dealt to the devices in snake order, so every device gets the same cells within ~1 % and the same
//     mix of length classes; a batch of equal-shaped pairs is simply cut into contiguous ranges.
//   * each share's letters are gathered piece by piece into that device's pinned staging buffer while the
//     previous piece is on its way to the GPU (align_host / GatherSrc); contiguous ranges need no gather.
//   * results: the per-pair arrays are scattered by pair id by the device threads; the segments of a device
//     stay on that device until all devices know their totals, then travel straight to their final place
//     in the one result block (no host-side copy of segments, no scan, no sort).
// Included at the end of ba_api.cu (inside its translation unit: it uses the scheduler's internals).

#include <condition_variable>
#include <functional>
#include <thread>

struct ba_multi {
    std::vector<int> devices;
    std::vector<ba_ctx *> ctx;
    std::string err;
    std::mutex err_mu;
    int64_t opt_contiguous = 1;   // cut equal-shaped batches into contiguous ranges (no gather)
    // one persistent host thread per device
    struct Worker {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::function<void()> job;
        bool has_job = false, quit = false, done = true;
    };
    std::vector<Worker *> workers;
};

namespace {

void multi_worker_main(ba_multi::Worker *w)
{
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) return;
            job = std::move(w->job);
            w->has_job = false;
        }
        job();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->done = true;
        }
        w->cv.notify_all();
    }
}

// run fn(d) for every device, each on its own thread, and wait for all
void multi_parallel(ba_multi *m, const std::function<void(int)> &fn)
{
    const int D = (int)m->ctx.size();
#ifdef BA_EMULATE
    for (int d = 0; d < D; d++) fn(d);   // the emulator keeps global fiber state: one device at a time
#else
    if (D == 1) {
        fn(0);
        return;
    }
    for (int d = 0; d < D; d++) {
        ba_multi::Worker *w = m->workers[(size_t)d];
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->job = [&fn, d] { fn(d); };
            w->has_job = true;
            w->done = false;
        }
        w->cv.notify_all();
    }
    for (int d = 0; d < D; d++) {
        ba_multi::Worker *w = m->workers[(size_t)d];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
    }
#endif
}

void multi_fail(ba_multi *m, int d, const std::string &msg)
{
    std::lock_guard<std::mutex> lk(m->err_mu);
    if (m->err.empty()) m->err = "device " + std::to_string(m->devices[(size_t)d]) + ": " + msg;
}

// Length-bucketed dealing.  shares[d] = pair ids of device d.  allow_ranges (all pairs of one shape): the
// shares are contiguous ranges of the caller's order, share d is [cut[d], cut[d+1]); returns that flag.
bool multi_deal(const int32_t *ref_len, const int32_t *qry_len, int64_t n, int D, bool allow_ranges,
                std::vector<std::vector<int64_t>> &shares, std::vector<int64_t> &cut, bool wide10 = true)
{
    shares.assign((size_t)D, std::vector<int64_t>());
    cut.assign((size_t)D + 1, 0);
    if (allow_ranges) {   // the caller has checked that all pairs have one shape
        for (int d = 0; d <= D; d++) cut[(size_t)d] = n * d / D;
        return true;
    }
    // quarter-octave classes of cells: key = 4 * log2(cells) from the exponent and two mantissa bits
    // in front of it the planner's own first key, the strip width of the pair's kernel instance (widest
    // first, ba_plan.h), so that a share is in the order its device works through it
    constexpr int NBC = 4 * 64 + 5;
    constexpr int NB = (BA16_MAX_COLS + 1) * NBC - 1;
    auto bucket = [wide10](int64_t cells, int m) -> int {
        const int cls = cols16_for(m, wide10);
        if (cells <= 0) return cls * NBC;
        const int e = 63 - __builtin_clzll((unsigned long long)cells);
        const int frac = e >= 2 ? (int)((cells >> (e - 2)) & 3) : 0;
        return cls * NBC + 4 * e + frac + 1;
    };
    std::vector<int64_t> cnt(NB + 1, 0);
    std::vector<uint16_t> bk((size_t)n);
    for (int64_t p = 0; p < n; p++) {
        const int b = bucket((int64_t)ref_len[p] * (int64_t)qry_len[p], qry_len[p]);
        bk[(size_t)p] = (uint16_t)b;
        cnt[(size_t)b]++;
    }
    // positions in largest-class-first order
    std::vector<int64_t> start(NB + 1, 0);
    int64_t acc = 0;
    for (int b = NB; b >= 0; b--) {
        start[(size_t)b] = acc;
        acc += cnt[(size_t)b];
    }
    for (int d = 0; d < D; d++) shares[(size_t)d].reserve((size_t)(n / D + 2));
    // pair p takes position start[bucket]++ of the largest-first order; position x goes to device snake(x).
    // A share keeps that order: it is (nearly) the order the device's own planner works in, so the fill can
    // start on the first uploaded pieces of the share while the rest is still being gathered and copied.
    std::vector<int64_t> ord((size_t)n);
    for (int64_t p = 0; p < n; p++) ord[(size_t)start[bk[(size_t)p]]++] = p;
    for (int64_t x = 0; x < n; x++) {
        const int64_t r = x % (2 * D);
        const int d = (int)(r < D ? r : 2 * D - 1 - r);
        shares[(size_t)d].push_back(ord[(size_t)x]);
    }
    return false;
}

}  // namespace

extern "C" {

int ba_create_multi(const int *devices, int n_devices, ba_multi **out)
{
    if (!out) return BA_ERR_ARG;
    *out = nullptr;
    ba_multi *m = new ba_multi();
    *out = m;
    if (n_devices <= 0 || !devices) {
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt < 1) {
            (void)cudaGetLastError();
            m->err = "no CUDA device";
            return BA_ERR_CUDA;
        }
        for (int d = 0; d < cnt; d++) m->devices.push_back(d);
    } else {
        m->devices.assign(devices, devices + n_devices);
    }
    for (int dev : m->devices) {
        ba_ctx *c = nullptr;
        const int rc = ba_create(dev, &c);
        if (rc != BA_OK) {
            m->err = std::string("device ") + std::to_string(dev) + ": " + (c ? ba_last_error(c) : "ba_create failed");
            if (c) ba_destroy(c);
            return rc;
        }
        m->ctx.push_back(c);
    }
#ifndef BA_EMULATE
    for (size_t d = 0; d < m->ctx.size(); d++) {
        ba_multi::Worker *w = new ba_multi::Worker();
        w->th = std::thread(multi_worker_main, w);
        m->workers.push_back(w);
    }
#endif
    return BA_OK;
}

void ba_destroy_multi(ba_multi *m)
{
    if (!m) return;
    for (ba_multi::Worker *w : m->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (ba_ctx *c : m->ctx) ba_destroy(c);
    delete m;
}

int ba_multi_devices(const ba_multi *m) { return m ? (int)m->ctx.size() : 0; }

const char *ba_multi_last_error(const ba_multi *m) { return m ? m->err.c_str() : "null multi-device context"; }

int ba_multi_set_scoring(ba_multi *m, int kind, int affine, const int64_t *la, int let, int alpha_len,
                         int64_t gap_open, const int32_t *index)
{
    if (!m) return BA_ERR_ARG;
    m->err.clear();
    for (size_t d = 0; d < m->ctx.size(); d++) {
        const int rc = ba_set_scoring(m->ctx[d], kind, affine, la, let, alpha_len, gap_open, index);
        if (rc != BA_OK) {
            multi_fail(m, (int)d, ba_last_error(m->ctx[d]));
            return rc;
        }
    }
    return BA_OK;
}

int ba_multi_set_option(ba_multi *m, const char *key, int64_t value)
{
    if (!m || !key) return BA_ERR_ARG;
    if (!strcmp(key, "contiguous_shares")) {   // 0: always deal by length buckets, even equal-shaped batches
        m->opt_contiguous = value;
        return BA_OK;
    }
    for (size_t d = 0; d < m->ctx.size(); d++) {
        const int rc = ba_set_option(m->ctx[d], key, value);
        if (rc != BA_OK) {
            multi_fail(m, (int)d, ba_last_error(m->ctx[d]));
            return rc;
        }
    }
    return BA_OK;
}

int ba_align_batch_multi(ba_multi *m, const uint8_t *letters, int stride, const int64_t *ref_off,
                         const int32_t *ref_len, const int64_t *qry_off, const int32_t *qry_len, int64_t n,
                         ba_result *out)
{
    if (!m || !out || n < 0 || (stride != 1 && stride != 2)) return BA_ERR_ARG;
    memset(out, 0, sizeof *out);
    m->err.clear();
    if (n > 0 && (!ref_off || !ref_len || !qry_off || !qry_len)) return BA_ERR_ARG;
    if (n > 0x7fffffff) return BA_ERR_ARG;
    const int D = (int)m->ctx.size();
    if (D < 1) return BA_ERR_ARG;
    ba_ctx *c0 = m->ctx[0];
    const double t0 = now_s();
    // argument check and "are all pairs of one shape?" in one pass, every device thread over its slice
    std::vector<int> bad((size_t)D, 0), differs((size_t)D, 0);
    multi_parallel(m, [&](int d) {
        const int64_t lo = n * d / D, hi = n * (d + 1) / D;
        int b = 0, df = 0;
        for (int64_t p = lo; p < hi; p++) {
            b |= (ref_len[p] < 0) | (qry_len[p] < 0) | (ref_off[p] < 0) | (qry_off[p] < 0);
            df |= (ref_len[p] != ref_len[0]) | (qry_len[p] != qry_len[0]);
        }
        bad[(size_t)d] = b;
        differs[(size_t)d] = df;
    });
    bool uniform = m->opt_contiguous != 0;
    for (int d = 0; d < D; d++) {
        if (bad[(size_t)d]) {
            m->err = "negative sequence length or offset";
            return BA_ERR_ARG;
        }
        if (differs[(size_t)d]) uniform = false;
    }
    std::vector<std::vector<int64_t>> shares;
    std::vector<int64_t> cut;
    // (one device: its share is the whole batch as the caller laid it out -- nothing to deal or gather)
    const bool ranges = multi_deal(ref_len, qry_len, n, D, uniform || D == 1, shares, cut, m->ctx[0]->opt_wide10 != 0);
    const double deal_s = now_s() - t0;
    BA_TSEC("multi deal", t0);

    // the one result block (pinned; owned by device 0's pool, freed by ba_multi_free_result)
    out->n = n;
    out->status = (int32_t *)pinned_get(c0, sizeof(int32_t) * (size_t)n);
    out->err_pos = (int64_t *)pinned_get(c0, sizeof(int64_t) * (size_t)n);
    out->seg_start = (int64_t *)pinned_get(c0, sizeof(int64_t) * (size_t)n);
    out->seg_count = (int32_t *)pinned_get(c0, sizeof(int32_t) * (size_t)n);
    auto release = [&]() {
        pinned_put(c0, out->status);
        pinned_put(c0, out->err_pos);
        pinned_put(c0, out->seg_start);
        pinned_put(c0, out->seg_count);
        pinned_put(c0, out->segs);
        memset(out, 0, sizeof *out);
    };
    if (!out->status || !out->err_pos || !out->seg_start || !out->seg_count) {
        release();
        m->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    if (n == 0) return BA_OK;

    struct Share {
        int64_t cnt = 0;
        std::vector<int64_t> loff_r, loff_q;
        std::vector<int32_t> lref, lqry;
        ba_result res;
        int rc = BA_OK;
        uint8_t *stage = nullptr;
        int64_t base = 0;
    };
    std::vector<Share> S((size_t)D);
    // ---- phase 1: every device aligns its share; segments stay on the device
    multi_parallel(m, [&](int d) {
        Share &s = S[(size_t)d];
        ba_ctx *c = m->ctx[(size_t)d];
        memset(&s.res, 0, sizeof s.res);
        if (ranges) {
            // contiguous range: the caller's arrays serve as they are; only the letters the range needs travel
            const int64_t lo = cut[(size_t)d], hi = cut[(size_t)d + 1];
            s.cnt = hi - lo;
            if (s.cnt == 0) return;
            int64_t bmin = INT64_MAX;
            for (int64_t p = lo; p < hi; p++) {
                if (ref_len[p]) bmin = std::min(bmin, ref_off[p]);
                if (qry_len[p]) bmin = std::min(bmin, qry_off[p]);
            }
            if (bmin == INT64_MAX) bmin = 0;
            s.loff_r.resize((size_t)s.cnt);
            s.loff_q.resize((size_t)s.cnt);
            for (int64_t k = 0; k < s.cnt; k++) {
                s.loff_r[(size_t)k] = ref_len[lo + k] ? ref_off[lo + k] - bmin : 0;
                s.loff_q[(size_t)k] = qry_len[lo + k] ? qry_off[lo + k] - bmin : 0;
            }
            s.rc = align_host(c, letters ? letters + bmin : nullptr, stride, s.loff_r.data(), ref_len + lo,
                              s.loff_q.data(), qry_len + lo, s.cnt, &s.res, nullptr, true);
        } else {
            const std::vector<int64_t> &ids = shares[(size_t)d];
            s.cnt = (int64_t)ids.size();
            if (s.cnt == 0) return;
            s.loff_r.resize((size_t)s.cnt);
            s.loff_q.resize((size_t)s.cnt);
            s.lref.resize((size_t)s.cnt);
            s.lqry.resize((size_t)s.cnt);
            int64_t o = 0;
            for (int64_t k = 0; k < s.cnt; k++) {
                const int64_t p = ids[(size_t)k];
                s.lref[(size_t)k] = ref_len[p];
                s.lqry[(size_t)k] = qry_len[p];
                s.loff_r[(size_t)k] = o;
                o += (int64_t)ref_len[p] * stride;
                s.loff_q[(size_t)k] = o;
                o += (int64_t)qry_len[p] * stride;
            }
            s.stage = (uint8_t *)pinned_get(c, (size_t)o + 16);
            if (!s.stage) {
                s.rc = BA_ERR_NOMEM;
                c->err = "pinned host allocation failed";
            } else {
                GatherSrc gs;
                gs.letters = letters;
                gs.ref_off = ref_off;
                gs.qry_off = qry_off;
                gs.ids = ids.data();
                gs.stage = s.stage;
                s.rc = align_host(c, nullptr, stride, s.loff_r.data(), s.lref.data(), s.loff_q.data(), s.lqry.data(),
                                  s.cnt, &s.res, &gs, true);
                pinned_put(c, s.stage);
                s.stage = nullptr;
            }
        }
        if (s.rc != BA_OK) multi_fail(m, d, ba_last_error(c));
    });
    BA_TSEC("multi phase 1 (align)", t0);
    int rc = BA_OK;
    int64_t total = 0;
    for (int d = 0; d < D; d++) {
        if (S[(size_t)d].rc != BA_OK && rc == BA_OK) rc = S[(size_t)d].rc;
        S[(size_t)d].base = total;
        total += S[(size_t)d].cnt ? S[(size_t)d].res.n_segs : 0;
    }
    if (rc == BA_OK) {
        out->segs = (int32_t *)pinned_get(c0, sizeof(int32_t) * 5 * (size_t)(total + 1));
        if (!out->segs) {
            rc = BA_ERR_NOMEM;
            m->err = "pinned host allocation failed";
        }
    }
    if (rc != BA_OK) {
        for (int d = 0; d < D; d++)
            if (S[(size_t)d].cnt && S[(size_t)d].rc == BA_OK) ba_free_result(m->ctx[(size_t)d], &S[(size_t)d].res);
        release();
        return rc;
    }
    out->n_segs = total;
    // ---- phase 2: segments straight to their place, per-pair arrays scattered by pair id
    multi_parallel(m, [&](int d) {
        Share &s = S[(size_t)d];
        if (s.cnt == 0) return;
        ba_ctx *c = m->ctx[(size_t)d];
        cudaSetDevice(c->device);
        const int64_t ndev = c->deferred_dev_segs, nall = s.res.n_segs;
        int32_t *dst = out->segs + 5 * s.base;
        cudaError_t e = cudaSuccess;
        if (ndev > 0)
            e = cudaMemcpyAsync(dst, c->segs.p, sizeof(int32_t) * 5 * (size_t)ndev, cudaMemcpyDeviceToHost, c->s_d2h);
        if (nall > ndev && s.res.segs)   // the exact path's segments are on the host already
            memcpy(dst + 5 * ndev, s.res.segs + 5 * ndev, sizeof(int32_t) * 5 * (size_t)(nall - ndev));
        if (ranges) {
            const int64_t lo = cut[(size_t)d];
            memcpy(out->status + lo, s.res.status, sizeof(int32_t) * (size_t)s.cnt);
            memcpy(out->err_pos + lo, s.res.err_pos, sizeof(int64_t) * (size_t)s.cnt);
            memcpy(out->seg_count + lo, s.res.seg_count, sizeof(int32_t) * (size_t)s.cnt);
            for (int64_t k = 0; k < s.cnt; k++) out->seg_start[lo + k] = s.res.seg_count[k] ? s.base + s.res.seg_start[k] : 0;
        } else {
            const std::vector<int64_t> &ids = shares[(size_t)d];
            for (int64_t k = 0; k < s.cnt; k++) {
                const int64_t p = ids[(size_t)k];
                out->status[p] = s.res.status[k];
                out->err_pos[p] = s.res.err_pos[k];
                out->seg_count[p] = s.res.seg_count[k];
                out->seg_start[p] = s.res.seg_count[k] ? s.base + s.res.seg_start[k] : 0;
            }
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_d2h);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            s.rc = BA_ERR_CUDA;
            multi_fail(m, d, cudaGetErrorString(e));
        }
    });
    BA_TSEC("multi phase 2 (gather)", t0);
    // ---- bookkeeping: spans are the slowest device's, counters add up
    for (int d = 0; d < D; d++) {
        Share &s = S[(size_t)d];
        if (s.cnt == 0) continue;
        if (s.rc != BA_OK) rc = s.rc;
        out->ms_h2d = std::max(out->ms_h2d, s.res.ms_h2d);
        out->ms_kernels = std::max(out->ms_kernels, s.res.ms_kernels);
        out->ms_d2h = std::max(out->ms_d2h, s.res.ms_d2h);
        out->ms_fill = std::max(out->ms_fill, s.res.ms_fill);
        out->ms_trace = std::max(out->ms_trace, s.res.ms_trace);
        out->ms_host = std::max(out->ms_host, s.res.ms_host);
        out->cells += s.res.cells;
        out->gpu_launches += s.res.gpu_launches;
        out->fill_launches += s.res.fill_launches;
        out->code_bytes += s.res.code_bytes;
        out->pairs_fast16 += s.res.pairs_fast16;
        out->wavefront_launches += s.res.wavefront_launches;
        out->warp_trace_launches += s.res.warp_trace_launches;
        out->waves = std::max(out->waves, s.res.waves);
        ba_free_result(m->ctx[(size_t)d], &s.res);
    }
    out->ms_host += (float)(deal_s * 1e3);
    if (rc != BA_OK) {
        release();
        return rc;
    }
    BA_TSEC("multi total", t0);
    return BA_OK;
}

void ba_multi_free_result(ba_multi *m, ba_result *res)
{
    if (!m || !res || m->ctx.empty()) return;
    ba_free_result(m->ctx[0], res);
}

}  // extern "C"
