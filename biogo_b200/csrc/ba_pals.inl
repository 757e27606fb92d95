// ba_pals.inl -- C ABI of the PALS banded DP (ba_pals.cuh): batch set-up, launch, and the host-side
// de-duplication of align/pals/dp/align.go:88-138.  Included at the end of ba_api.cu.

namespace {

// stable insertion sort on one key, then the reference's two sweeps (see the tie note in biogo_align.h)
int64_t pals_dedup_host(ba_pals_hit *segs, int64_t n)
{
    if (n <= 0) return 0;
    auto sort_by = [&](bool by_end) {
        for (int64_t a = 1; a < n; a++) {
            ba_pals_hit x = segs[a];
            int64_t b = a;
            while (b > 0 && (by_end ? segs[b - 1].aepos > x.aepos : segs[b - 1].abpos > x.abpos)) {
                segs[b] = segs[b - 1];
                b--;
            }
            segs[b] = x;
        }
    };
    int64_t i, j;
    sort_by(false);
    for (i = 0; i < n; i = j) {
        for (j = i + 1; j < n; j++) {
            if (segs[j].abpos != segs[i].abpos) break;
            if (segs[j].bbpos != segs[i].bbpos) break;
            if (segs[j].score > segs[i].score) {
                segs[i].score = -1;
                i = j;
            } else {
                segs[j].score = -1;
            }
        }
    }
    sort_by(true);
    for (i = 0; i < n; i = j) {
        for (j = i + 1; j < n; j++) {
            if (segs[j].aepos != segs[i].aepos) break;
            if (segs[j].bepos != segs[i].bepos) break;
            if (segs[j].score > segs[i].score) {
                segs[i].score = -1;
                i = j;
            } else {
                segs[j].score = -1;
            }
        }
    }
    int64_t found = 0;
    for (i = 0; i < n; i++)
        if (segs[i].score >= 0) segs[found++] = segs[i];
    return found;
}

}  // namespace

extern "C" {

int ba_pals_dp_batch(ba_ctx *ctx, const uint8_t *letters, const int64_t *target_off, const int32_t *target_len,
                     const int64_t *query_off, const int32_t *query_len, int64_t n, const int32_t *index,
                     const int64_t *trap_off, const int32_t *traps, int kmer, int min_hit_length, double min_id,
                     const ba_pals_costs *costs, ba_pals_result *out)
{
    if (!ctx || !out || n < 0 || !costs || !index) return BA_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (!ctx->d_sc_first) {
        ctx->err = "context has no device (ba_create failed)";
        return BA_ERR_CUDA;
    }
    if (n > 0 && (!target_off || !target_len || !query_off || !query_len || !trap_off)) return BA_ERR_ARG;
    if (n > 0x3fffffff) return BA_ERR_ARG;
    if (costs->diff_cost < 1 || costs->max_igap < 0 || costs->block_cost < 0 || costs->rmatch_cost == 0.0) {
        ctx->err = "pals: DiffCost must be >= 1, MaxIGap and BlockCost >= 0, RMatchCost != 0";
        return BA_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Lease lease(ctx);
    out->n = n;
    out->hit_off = (int64_t *)lease.get(sizeof(int64_t) * (size_t)(n + 1));
    out->status = (int32_t *)lease.get(sizeof(int32_t) * (size_t)(n + 1));
    if (!out->hit_off || !out->status) {
        out->hit_off = nullptr;
        out->status = nullptr;
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    out->hit_off[0] = 0;
    if (n == 0) {
        lease.keep(out->hit_off);
        lease.keep(out->status);
        return BA_OK;
    }
    // ---- sizes: letters, DP vectors, recursion stacks, hit slots
    const int64_t ntraps = trap_off[n];
    if (ntraps < 0 || (ntraps > 0 && !traps)) return BA_ERR_ARG;
    int64_t letters_bytes = 0, vec_total = 0, stack_total = 0, hit_total = 0;
    std::vector<int64_t> h_off((size_t)n * 3 + 3);   // vec_off | stack_off | hit_off (capacity)
    std::vector<int32_t> h_cap((size_t)n * 2);       // stack_cap | hit_cap
    const int64_t ml = std::max(min_hit_length, 1);
    const int64_t cmax = std::max<int64_t>(std::max<int64_t>(costs->match_cost < 0 ? -costs->match_cost : costs->match_cost,
                                                             costs->diff_cost), 1);
    for (int64_t p = 0; p < n; p++) {
        if (target_len[p] < 0 || query_len[p] < 0 || target_off[p] < 0 || query_off[p] < 0 || trap_off[p + 1] < trap_off[p]) {
            ctx->err = "pals: negative length/offset or decreasing trapezoid offsets";
            return BA_ERR_ARG;
        }
        // 32-bit device lanes: scores stay below (rows + columns) * max cost, band positions times DiffCost too
        if (((int64_t)target_len[p] + query_len[p] + 8) * cmax >= (1LL << 29)) {
            ctx->err = "pals: sequence lengths times costs exceed the 32-bit device lanes";
            return BA_ERR_RANGE;
        }
        letters_bytes = std::max(letters_bytes, std::max(target_off[p] + target_len[p], query_off[p] + query_len[p]));
        h_off[(size_t)p] = vec_total;
        vec_total += 2 * ((int64_t)target_len[p] + 2);
        int64_t hc = 0, sc = 8;
        for (int64_t x = trap_off[p]; x < trap_off[p + 1]; x++) {
            const int64_t h = std::max<int64_t>((int64_t)traps[4 * x] - traps[4 * x + 1], 0);
            hc += 2 * (h / ml) + 4;      // every recursion call emits at most one hit; children are disjoint and > ml high
            int64_t d = 8;
            for (int64_t y = h; y > 0; y >>= 1) d += 2;
            sc = std::max(sc, d);        // children are at most about half as high: depth ~ log2(h)
        }
        h_off[(size_t)n + 1 + (size_t)p] = stack_total;
        h_cap[(size_t)p] = (int32_t)sc;
        stack_total += sc;
        h_off[(size_t)2 * n + 2 + (size_t)p] = hit_total;
        h_cap[(size_t)n + (size_t)p] = (int32_t)std::min<int64_t>(hc, 0x7fffffff);
        hit_total += hc;
    }
    if (letters_bytes > 0 && !letters) return BA_ERR_ARG;
    int rc;
    // ---- device buffers
    const size_t b8 = sizeof(int64_t) * (size_t)n, b4 = sizeof(int32_t) * (size_t)n;
    if ((rc = ensure(ctx, ctx->letters, (size_t)letters_bytes + 16)) != BA_OK) return rc;
    ctx->last_n = -1;
    if ((rc = ensure(ctx, ctx->pals_traps, sizeof(int32_t) * 4 * (size_t)(ntraps + 1) + (size_t)ntraps + 64)) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->pals_vec, sizeof(int32_t) * (size_t)(vec_total + 4))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->pals_stack, sizeof(int32_t) * 4 * (size_t)(stack_total + 1))) != BA_OK) return rc;
    if ((rc = ensure(ctx, ctx->pals_hits, sizeof(PalsHitDev) * (size_t)(hit_total + 1))) != BA_OK) return rc;
    // one staging block: [t_off | q_off | trap_off(n+1) | vec_off | stack_off | hit_off][t_len | q_len | stack_cap | hit_cap][valid 256]
    const size_t meta_bytes = 6 * b8 + 8 + 4 * b4 + 256;
    uint8_t *stage = (uint8_t *)lease.get(meta_bytes + 64);
    if (!stage) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    int64_t *s8 = (int64_t *)stage;
    memcpy(s8, target_off, b8);
    memcpy(s8 + n, query_off, b8);
    memcpy(s8 + 2 * n, trap_off, b8 + 8);
    memcpy(s8 + 3 * n + 1, h_off.data(), b8);
    memcpy(s8 + 4 * n + 1, h_off.data() + n + 1, b8);
    memcpy(s8 + 5 * n + 1, h_off.data() + 2 * n + 2, b8);
    int32_t *s4 = (int32_t *)(s8 + 6 * n + 1);
    memcpy(s4, target_len, b4);
    memcpy(s4 + n, query_len, b4);
    memcpy(s4 + 2 * n, h_cap.data(), b4);
    memcpy(s4 + 3 * n, h_cap.data() + n, b4);
    uint8_t *sv = (uint8_t *)(s4 + 4 * n);
    for (int k = 0; k < 256; k++) sv[k] = index[k] >= 0 ? 1 : 0;
    if ((rc = ensure(ctx, ctx->pals_meta, meta_bytes + 4 * b4 + 256)) != BA_OK) return rc;
    uint8_t *dm = (uint8_t *)ctx->pals_meta.p;
    CK(cudaEventRecord(ctx->ev[0], st));
    CK(cudaMemcpyAsync(dm, stage, meta_bytes, cudaMemcpyHostToDevice, st));
    if (letters_bytes > 0) CK(cudaMemcpyAsync(ctx->letters.p, letters, (size_t)letters_bytes, cudaMemcpyHostToDevice, st));
    if (ntraps > 0)
        CK(cudaMemcpyAsync(ctx->pals_traps.p, traps, sizeof(int32_t) * 4 * (size_t)ntraps, cudaMemcpyHostToDevice, st));
    uint8_t *d_covered = (uint8_t *)ctx->pals_traps.p + sizeof(int32_t) * 4 * (size_t)(ntraps + 1);
    CK(cudaMemsetAsync(d_covered, 0, (size_t)ntraps + 1, st));
    // results + the work counter live behind the staged block
    int32_t *d_res = (int32_t *)(dm + ((meta_bytes + 15) & ~(size_t)15));
    CK(cudaMemsetAsync(d_res, 0, 2 * b4 + 16, st));

    PalsParams P;
    const int64_t *d8 = (const int64_t *)dm;
    const int32_t *d4 = (const int32_t *)(d8 + 6 * n + 1);
    P.letters = (const uint8_t *)ctx->letters.p;
    P.t_off = d8;
    P.q_off = d8 + n;
    P.trap_off = d8 + 2 * n;
    P.vec_off = d8 + 3 * n + 1;
    P.stack_off = d8 + 4 * n + 1;
    P.hit_off = d8 + 5 * n + 1;
    P.t_len = d4;
    P.q_len = d4 + n;
    P.stack_cap = d4 + 2 * n;
    P.hit_cap = d4 + 3 * n;
    P.valid = (const uint8_t *)(d4 + 4 * n);
    P.traps = (const int32_t *)ctx->pals_traps.p;
    P.covered = d_covered;
    P.n = (int)n;
    P.kmer = kmer;
    P.min_len = min_hit_length;
    P.max_diff = 1 - min_id;
    P.rmatch_cost = costs->rmatch_cost;
    P.max_igap = costs->max_igap;
    P.diff_cost = costs->diff_cost;
    P.match_cost = costs->match_cost;
    P.block_cost = costs->block_cost;
    P.vec = (int32_t *)ctx->pals_vec.p;
    P.stack = (int32_t *)ctx->pals_stack.p;
    P.hits = (PalsHitDev *)ctx->pals_hits.p;
    P.hit_count = d_res;
    P.status = d_res + n;
    P.counter = (int *)(d_res + 2 * n);
    int64_t blocks = (n + 3) / 4;
    blocks = std::min<int64_t>(blocks, (int64_t)ctx->sm_count * 8);
    CK(cudaEventRecord(ctx->ev[1], st));
    BA_LAUNCH(ba_pals_dp_kernel, (unsigned)blocks, 128, st, P);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[2], st));
    // ---- results
    int32_t *h_cnt = (int32_t *)lease.get(2 * b4 + 16);
    PalsHitDev *h_hits = (PalsHitDev *)lease.get(sizeof(PalsHitDev) * (size_t)(hit_total + 1));
    if (!h_cnt || !h_hits) {
        ctx->err = "pinned host allocation failed";
        cudaStreamSynchronize(st);
        return BA_ERR_NOMEM;
    }
    CK(cudaMemcpyAsync(h_cnt, d_res, 2 * b4, cudaMemcpyDeviceToHost, st));
    if (hit_total > 0)
        CK(cudaMemcpyAsync(h_hits, ctx->pals_hits.p, sizeof(PalsHitDev) * (size_t)hit_total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
        out->ms_kernels = ms;
    }
    int64_t total = 0;
    for (int64_t p = 0; p < n; p++) {
        out->status[p] = h_cnt[n + p];
        if (h_cnt[n + p] != 0) {
            ctx->err = "pals: internal capacity exhausted (hit slots or recursion stack)";
            return BA_ERR_RANGE;
        }
        total += h_cnt[p];
    }
    out->hits = (ba_pals_hit *)lease.get(sizeof(ba_pals_hit) * (size_t)(total + 1));
    if (!out->hits) {
        ctx->err = "pinned host allocation failed";
        return BA_ERR_NOMEM;
    }
    int64_t o = 0;
    for (int64_t p = 0; p < n; p++) {
        const PalsHitDev *src = h_hits + h_off[(size_t)2 * n + 2 + (size_t)p];
        ba_pals_hit *dst = out->hits + o;
        for (int k = 0; k < h_cnt[p]; k++) {
            dst[k].abpos = src[k].abpos;
            dst[k].bbpos = src[k].bbpos;
            dst[k].aepos = src[k].aepos;
            dst[k].bepos = src[k].bepos;
            dst[k].low_diagonal = src[k].low_diagonal;
            dst[k].high_diagonal = src[k].high_diagonal;
            dst[k].score = src[k].score;
            dst[k].error = src[k].error;
        }
        out->hit_off[p] = o;
        o += pals_dedup_host(dst, h_cnt[p]);
    }
    out->hit_off[n] = o;
    out->n_hits = o;
    lease.keep(out->hit_off);
    lease.keep(out->status);
    lease.keep(out->hits);
    return BA_OK;
}

void ba_pals_free_result(ba_ctx *ctx, ba_pals_result *res)
{
    if (!ctx || !res) return;
    pinned_put(ctx, res->hit_off);
    pinned_put(ctx, res->status);
    pinned_put(ctx, res->hits);
    memset(res, 0, sizeof *res);
}

}  // extern "C"
