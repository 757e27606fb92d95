/*
 * pals_oracle.c -- CPU restatement of biogo's align/pals/dp package: banded
 * x-drop dynamic programming over filter trapezoids.  TEST INFRASTRUCTURE: only tests/, bench.py's CPU
 * legs and __graft_entry__.smoke() may use it; the product path is biogo_b200/csrc/ba_pals.cuh.
 *
 * Literal restatement, int64 = Go int, float64 operations in the reference's order (compile with
 * -ffp-contract=off: Go on amd64 does not fuse multiply-add):
 *   kernel.alignRecursion   align/pals/dp/kernel.go:52-129
 *   kernel.traceForward     align/pals/dp/kernel.go:143-296
 *   kernel.traceReverse     align/pals/dp/kernel.go:298-454
 *   Aligner.AlignTraps      align/pals/dp/align.go:55-141  (incl. the start/end de-duplication)
 * The two DP vectors are indexed by absolute target position here; the reference's offsetSlice views
 * (kernel.go:33-46) address the same cells, and every cell a row reads was written by the row before.
 *
 * PINNED on the reference's own golden test (align/pals/dp/dp_test.go:22-84: de Bruijn sequences, six
 * trapezoids -> five hits with their Error values), tests/test_pals.py.  One thing is NOT pinned: the order
 * sort.Sort gives hits with EQUAL Abpos (resp. Aepos) in AlignTraps' de-duplication -- Go's sort is not
 * stable and its algorithm changed in Go 1.19; the golden vector has no such ties.  This restatement
 * sorts stably (what Go's insertion sort does for fewer than 12 hits).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t abpos, bbpos, aepos, bepos, low_diagonal, high_diagonal, score;
    double error;
} pals_hit;

typedef struct {
    int64_t max_igap, diff_cost, same_cost, match_cost, block_cost;
    double rmatch_cost;
} pals_costs;

typedef struct {
    int64_t top, bottom, left, right;
} pals_trap;

typedef struct {
    const uint8_t *target, *query;
    int64_t tlen, qlen;
    int64_t min_len;
    double max_diff;
    const int32_t *value_to_code;
    pals_costs c;
    pals_hit low_end, high_end;
    int64_t *vec[2];
    const pals_trap *traps;
    int64_t ntraps;
    uint8_t *covered;
    int64_t slot;
    pals_hit *out;
    int64_t nout, cap;
} kernel;

static void emit(kernel *k, pals_hit h)
{
    if (k->nout == k->cap) {
        k->cap = k->cap ? 2 * k->cap : 64;
        k->out = (pals_hit *)realloc(k->out, sizeof(pals_hit) * (size_t)k->cap);
    }
    k->out[k->nout++] = h;
}

static int same(const kernel *k, int64_t i, int64_t tj)
{
    return k->query[i] == k->target[tj] && k->value_to_code[k->query[i]] >= 0;
}

/* kernel.go:143-296 */
static void trace_forward(kernel *k, int64_t mid, int64_t low, int64_t high)
{
    int odd = 0;
    int64_t max_score, max_left, max_right, max_i, max_j, i, j;
    const int64_t tl = k->tlen, D = k->c.diff_cost;
    if (low < 0) low = 0;
    if (low > tl) low = tl;
    if (high > tl) high = tl;
    if (high < low) high = low;
    int64_t *this_v = k->vec[0], *that_v;
    for (j = low; j <= high; j++) this_v[j] = 0;
    high += k->c.max_igap;
    if (high > tl) high = tl;
    for (; j <= high; j++) this_v[j] = this_v[j - 1] - D;
    max_score = 0;
    max_right = mid - low;
    max_left = mid - high;
    max_i = mid;
    max_j = low;
    for (i = mid; low <= high && i < k->qlen; i++) {
        int64_t cost, score;
        that_v = this_v;
        this_v = odd ? k->vec[0] : k->vec[1];
        odd = !odd;
        score = that_v[low];
        this_v[low] = score - D;
        cost = this_v[low];
        for (j = low + 1; j <= high; j++) {
            int64_t ratchet, temp;
            temp = cost;
            cost = score;
            score = that_v[j];
            if (same(k, i, j - 1)) cost += k->c.match_cost;
            ratchet = cost;
            if (score > ratchet) ratchet = score;
            if (temp > ratchet) ratchet = temp;
            cost = ratchet - D;
            this_v[j] = cost;
            if (cost >= max_score) {
                max_score = cost;
                max_i = i + 1;
                max_j = j;
            }
        }
        if (j <= tl) {
            int64_t ratchet;
            if (same(k, i, j - 1)) score += k->c.match_cost;
            ratchet = score;
            if (cost > ratchet) ratchet = cost;
            score = ratchet - D;
            this_v[j] = score;
            if (score > max_score) {
                max_score = score;
                max_i = i + 1;
                max_j = j;
            }
            for (j++; j <= tl; j++) {
                score -= D;
                if (score < max_score - k->c.block_cost) break;
                this_v[j] = score;
            }
        }
        high = j - 1;
        while (low <= high && this_v[low] < max_score - k->c.block_cost) low++;
        while (low <= high && this_v[high] < max_score - k->c.block_cost) high--;
        if ((i + 1) - low > max_right) max_right = (i + 1) - low;
        if ((i + 1) - high < max_left) max_left = (i + 1) - high;
    }
    k->low_end.aepos = max_j;
    k->low_end.bepos = max_i;
    k->low_end.low_diagonal = max_left;
    k->low_end.high_diagonal = max_right;
    k->low_end.score = max_score;
}

/* kernel.go:298-454 */
static void trace_reverse(kernel *k, int64_t top, int64_t low, int64_t high, int64_t bottom, int64_t xfactor)
{
    int odd = 0;
    int64_t max_score, max_left, max_right, max_i, max_j, i, j;
    const int64_t tl = k->tlen, D = k->c.diff_cost;
    if (low < 0) low = 0;
    if (low > tl) low = tl;
    if (high > tl) high = tl;
    if (high < low) high = low;
    int64_t *this_v = k->vec[0], *that_v;
    for (j = high; j >= low; j--) this_v[j] = 0;
    low -= k->c.max_igap;
    if (low < 0) low = 0;
    for (; j >= low; j--) this_v[j] = this_v[j + 1] - D;
    max_score = 0;
    max_right = top - low;
    max_left = top - high;
    max_i = top;
    max_j = low;
    if (top - 1 <= bottom) xfactor = k->c.block_cost;
    for (i = top - 1; low <= high && i >= 0; i--) {
        int64_t cost, score;
        that_v = this_v;
        this_v = odd ? k->vec[0] : k->vec[1];
        odd = !odd;
        score = that_v[high];
        this_v[high] = score - D;
        cost = this_v[high];
        for (j = high - 1; j >= low; j--) {
            int64_t ratchet, temp;
            temp = cost;
            cost = score;
            score = that_v[j];
            if (same(k, i, j)) cost += k->c.match_cost;
            ratchet = cost;
            if (score > ratchet) ratchet = score;
            if (temp > ratchet) ratchet = temp;
            cost = ratchet - D;
            this_v[j] = cost;
            if (cost >= max_score) {
                max_score = cost;
                max_i = i;
                max_j = j;
            }
        }
        if (j >= 0) {
            int64_t ratchet;
            if (same(k, i, j)) score += k->c.match_cost;
            ratchet = score;
            if (cost > ratchet) ratchet = cost;
            score = ratchet - D;
            this_v[j] = score;
            if (score > max_score) {
                max_score = score;
                max_i = i;
                max_j = j;
            }
            for (j--; j >= 0; j--) {
                score -= D;
                if (score < max_score - xfactor) break;
                this_v[j] = score;
            }
        }
        low = j + 1;
        while (low <= high && this_v[low] < max_score - xfactor) low++;
        while (low <= high && this_v[high] < max_score - xfactor) high--;
        if (i == bottom) xfactor = k->c.block_cost;
        if (i - low > max_right) max_right = i - low;
        if (i - high < max_left) max_left = i - high;
    }
    k->high_end.abpos = max_j;
    k->high_end.bbpos = max_i;
    k->high_end.low_diagonal = max_left;
    k->high_end.high_diagonal = max_right;
    k->high_end.score = max_score;
}

/* kernel.go:52-129 */
static void align_recursion(kernel *k, pals_trap t)
{
    const int64_t mid = (t.bottom + t.top) / 2;
    trace_forward(k, mid, mid - t.right, mid - t.left);
    for (int64_t x = 1; x == 1 || (k->high_end.bbpos > mid + x * k->c.max_igap && k->high_end.score < k->low_end.score); x++)
        trace_reverse(k, k->low_end.bepos, k->low_end.aepos, k->low_end.aepos, mid + k->c.max_igap,
                      k->c.block_cost + 2 * x * k->c.diff_cost);
    k->high_end.aepos = k->low_end.aepos;
    k->high_end.bepos = k->low_end.bepos;
    pals_trap low_trap = t, high_trap = t;
    low_trap.top = k->high_end.bbpos - k->c.max_igap;
    high_trap.bottom = k->high_end.bepos + k->c.max_igap;
    if (k->high_end.bepos - k->high_end.bbpos >= k->min_len && k->high_end.aepos - k->high_end.abpos >= k->min_len) {
        int64_t indel = (k->high_end.abpos - k->high_end.bbpos) - (k->high_end.aepos - k->high_end.bepos);
        if (indel < 0) indel = -indel;
        const double identity = (1 / k->c.rmatch_cost) - (double)(k->high_end.score - indel) /
                                                             (k->c.rmatch_cost * (double)(k->high_end.bepos - k->high_end.bbpos));
        if (identity <= k->max_diff) {
            k->high_end.error = identity;
            /* kernel.go:81-118: `i` ranges over trapezoids[slot+1:] but indexes `covered` from 0 -- restated as is */
            for (int64_t i = 0; k->slot + 1 + i < k->ntraps; i++) {
                const pals_trap trap = k->traps[k->slot + 1 + i];
                int64_t trap_a, trap_b, cov_a, cov_b;
                if (trap.bottom >= k->high_end.bepos) break;
                trap_b = trap.top - trap.bottom + 1;
                trap_a = trap.right - trap.left + 1;
                cov_a = trap.left < k->high_end.low_diagonal ? k->high_end.low_diagonal : trap.left;
                cov_b = trap.right > k->high_end.high_diagonal ? k->high_end.high_diagonal : trap.right;
                if (cov_a > cov_b) continue;
                cov_a = cov_b - cov_a + 1;
                cov_b = trap.top > k->high_end.bepos ? k->high_end.bepos - trap.bottom + 1 : trap_b;
                if (((double)cov_a / (double)trap_a) * ((double)cov_b / (double)trap_b) > 0.99) k->covered[i] = 1;
            }
            /* Diagonals to this point are query-target, not target-query. */
            const int64_t lo = k->high_end.low_diagonal, hi = k->high_end.high_diagonal;
            k->high_end.low_diagonal = -hi;
            k->high_end.high_diagonal = -lo;
            emit(k, k->high_end);
        }
    }
    if (low_trap.top - low_trap.bottom > k->min_len && low_trap.top < t.top - k->c.max_igap) align_recursion(k, low_trap);
    if (high_trap.top - high_trap.bottom > k->min_len) align_recursion(k, high_trap);
}

/* stable insertion sorts on one key (see the header about ties) */
static void sort_by(pals_hit *h, int64_t n, int by_end)
{
    for (int64_t a = 1; a < n; a++) {
        pals_hit x = h[a];
        int64_t b = a;
        while (b > 0 && (by_end ? h[b - 1].aepos > x.aepos : h[b - 1].abpos > x.abpos)) {
            h[b] = h[b - 1];
            b--;
        }
        h[b] = x;
    }
}

/* align.go:88-138: remove lower scoring segments that begin or end at the same point as a higher scoring one */
int64_t pals_dedup(pals_hit *segs, int64_t n)
{
    int64_t i, j;
    if (n <= 0) return 0;
    sort_by(segs, n, 0);
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
    sort_by(segs, n, 1);
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

/* Aligner.AlignTraps (align.go:55-141).  traps: ntraps x {top, bottom, left, right}.  kmer = Aligner.k.
 * dedup = 0 returns the hits in emission order (what the device kernel produces before the host-side
 * de-duplication).  *out is malloc'ed (free with pals_free).  Returns the number of hits. */
int64_t pals_align_traps(const uint8_t *target, int64_t tlen, const uint8_t *query, int64_t qlen, const int32_t *value_to_code,
                         const int64_t *traps, int64_t ntraps, int64_t kmer, int64_t min_hit_length, double min_id,
                         const pals_costs *costs, int dedup, pals_hit **out)
{
    kernel k;
    memset(&k, 0, sizeof k);
    k.target = target;
    k.query = query;
    k.tlen = tlen;
    k.qlen = qlen;
    k.value_to_code = value_to_code;
    k.min_len = min_hit_length;
    k.max_diff = 1 - min_id;
    k.c = *costs;
    k.vec[0] = (int64_t *)calloc((size_t)tlen + 4, sizeof(int64_t));
    k.vec[1] = (int64_t *)calloc((size_t)tlen + 4, sizeof(int64_t));
    k.traps = (const pals_trap *)traps;
    k.ntraps = ntraps;
    k.covered = (uint8_t *)calloc((size_t)ntraps + 1, 1);
    for (int64_t i = 0; i < ntraps; i++) {
        const pals_trap t = k.traps[i];
        if (!k.covered[i] && t.top - t.bottom >= kmer) {
            k.slot = i;
            align_recursion(&k, t);
        }
    }
    free(k.vec[0]);
    free(k.vec[1]);
    free(k.covered);
    int64_t n = k.nout;
    if (dedup) n = pals_dedup(k.out, n);
    *out = k.out;
    return n;
}

void pals_free(void *p) { free(p); }

/* util.DeBruijn (util/debruijn.go:8-40): the FKM algorithm; k letters, words of length n.  buf needs k^n bytes. */
static void db_rec(uint8_t *a, uint8_t *s, int64_t *ns, int k, int n, int t, int p)
{
    if (t > n) {
        if (n % p == 0)
            for (int j = 1; j <= p; j++) s[(*ns)++] = a[j];
    } else {
        a[t] = a[t - p];
        db_rec(a, s, ns, k, n, t + 1, p);
        for (int j = a[t - p] + 1; j < k; j++) {
            a[t] = (uint8_t)j;
            db_rec(a, s, ns, k, n, t + 1, t);
        }
    }
}

int64_t pals_debruijn(int k, int n, uint8_t *buf)
{
    if (k == 0) return 0;
    if (k == 1) {
        memset(buf, 0, (size_t)n);
        return n;
    }
    uint8_t a[512];
    memset(a, 0, sizeof a);
    int64_t ns = 0;
    db_rec(a, buf, &ns, k, n, 1, 1);
    return ns;
}
