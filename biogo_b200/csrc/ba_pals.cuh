// ba_pals.cuh -- align/pals/dp on the device: the banded x-drop dynamic programming PALS
// runs over filter trapezoids (align/pals/dp/kernel.go:52-454, align.go:55-141).
//
// Parallel structure.  AlignTraps is sequential BETWEEN trapezoids of one (target, query) problem -- a hit
// marks later trapezoids as covered and the recursion's order decides the order of the hits -- so a PROBLEM
// is the unit that is independent: one warp per problem, many problems per batch (PALS aligns many contig
// pairs).  INSIDE a trace the rows are sequential too (row i needs row i-1 and the running maximum trims
// the band), but a row is not: with a_j = max(prev[j-1] + match, prev[j]) the recurrence
// c_j = max(a_j, c_{j-1}) - D becomes a prefix maximum of b_j = a_j + (x-1) D over the band position x
// (e_x = c_x + x D = max(b_x, e_{x-1})), which the 32 lanes compute 32 columns at a time with five shuffles.
// The running maximum's "last cell in scan order that is >= the maximum so far" is the last position of
// the row maximum, the x-drop extension past the band is an arithmetic sequence, and the band trimming is a
// ballot.  Integer arithmetic is 32-bit (the host checks the ranges); the identity and coverage tests are
// float64 with explicitly rounded operations in the reference's order (no fused multiply-add).
//
// Hits come out in the reference's emission order; the de-duplication by start / end point
// (align.go:88-138) runs on the host over the few hits of a problem (ba_api.cu).
#pragma once

#include "ba_common.cuh"

struct PalsHitDev {
    int32_t abpos, bbpos, aepos, bepos, low_diagonal, high_diagonal, score, pad;
    double error;
};

struct PalsParams {
    const uint8_t *letters;
    const int64_t *t_off, *q_off;
    const int32_t *t_len, *q_len;
    const int64_t *trap_off;     // n + 1: trapezoids of problem p are [trap_off[p], trap_off[p+1])
    const int32_t *traps;        // {top, bottom, left, right} each
    int n;
    int kmer, min_len;
    double max_diff, rmatch_cost;
    int max_igap, diff_cost, match_cost, block_cost;
    const uint8_t *valid;        // 256: letter is in the target's alphabet (valueToCode >= 0)
    int32_t *vec;                // DP vectors: 2 x vec_len[p] ints at vec_off[p]
    const int64_t *vec_off;
    int32_t *stack;              // pending trapezoids of the recursion: 4 ints each, stack_cap[p] entries at stack_off[p]
    const int64_t *stack_off;
    const int32_t *stack_cap;
    uint8_t *covered;            // per trapezoid (indexed like traps)
    PalsHitDev *hits;            // hit_cap[p] slots at hit_off[p]
    const int64_t *hit_off;
    const int32_t *hit_cap;
    int32_t *hit_count;          // per problem (may exceed the capacity: then status = 1)
    int32_t *status;             // 0 ok, 1 hit slots exhausted, 2 recursion stack exhausted
    int *counter;                // next problem
};

struct PalsEnd {
    int a, b, lo, hi, score;     // a: Aepos / Abpos, b: Bepos / Bbpos
};

struct PalsProb {
    const uint8_t *target, *query;
    int tl, ql;
    int32_t *v0, *v1;
};

#ifdef BA_EMULATE
static inline double ba_dmul(double a, double b) { volatile double r = a * b; return r; }
static inline double ba_ddiv(double a, double b) { volatile double r = a / b; return r; }
static inline double ba_dsub(double a, double b) { volatile double r = a - b; return r; }
#else
__device__ __forceinline__ double ba_dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ba_ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double ba_dsub(double a, double b) { return __dsub_rn(a, b); }
#endif

// One trace.  FWD: kernel.traceForward(mid = row0, low, high): rows row0, row0+1, ... of the query, columns
// scanned upwards, cell (i, j) compares query[i] with target[j-1].  !FWD: kernel.traceReverse(top = row0, low,
// high, bottom, xfactor): rows row0-1, row0-2, ..., columns scanned downwards, query[i] against target[j].
// Every lane returns the same result.
template <bool FWD>
__device__ __forceinline__ void pals_trace(const PalsParams &P, const PalsProb &pr, int row0, int low, int high, int bottom,
                                           int xfactor, PalsEnd &out)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int tl = pr.tl, D = P.diff_cost, MC = P.match_cost;
    if (low < 0) low = 0;
    if (low > tl) low = tl;
    if (high > tl) high = tl;
    if (high < low) high = low;
    int32_t *cur = pr.v0, *prev = pr.v1;
    // basis: zeros over [low, high], then a slope of -D over MaxIGap more columns in scan direction
    for (int j = low + lane; j <= high; j += 32) cur[j] = 0;
    if (FWD) {
        const int h2 = min(high + P.max_igap, tl);
        for (int j = high + 1 + lane; j <= h2; j += 32) cur[j] = -(j - high) * D;
        high = h2;
    } else {
        const int l2 = max(low - P.max_igap, 0);
        for (int j = low - 1 - lane; j >= l2; j -= 32) cur[j] = -(low - j) * D;
        low = l2;
    }
    __syncwarp();
    int max_score = 0;
    int max_right = row0 - low, max_left = row0 - high, max_i = row0, max_j = low;
    if (!FWD && row0 - 1 <= bottom) xfactor = P.block_cost;
    const int xf_fwd = P.block_cost;
    for (int i = FWD ? row0 : row0 - 1; low <= high && (FWD ? i < pr.ql : i >= 0); i += FWD ? 1 : -1) {
        {
            int32_t *t_ = prev;
            prev = cur;
            cur = t_;
        }
        const unsigned qi = pr.query[i];
        const bool qv = P.valid[qi] != 0;
        const int xf = FWD ? xf_fwd : xfactor;
        // scan order: position x = 0 is column `first`, x grows towards `last`
        const int first = FWD ? low : high, W = high - low + 1;
        const int step = FWD ? 1 : -1;
        // x = 0: cur = prev - D (not a candidate for the maximum)
        int carry = prev[first] - D;
        if (lane == 0) cur[first] = carry;
        int rowmax = INT32_MIN, rowarg = 0;   // maximum over x >= 1 and the LAST x attaining it
        for (int x0 = 1; x0 < W; x0 += 32) {
            const int x = x0 + lane;
            const bool on = x < W;
            const int j = first + step * x;
            int b = INT32_MIN / 2;
            if (on) {
                const int jp = j - step;                     // previous column in scan order
                const int tpos = FWD ? j - 1 : j;            // target letter of this cell
                const int diag = prev[jp] + ((qv && pr.target[tpos] == qi) ? MC : 0);
                const int a = max(diag, prev[j]);
                b = a + (x - 1) * D;
            }
            // inclusive prefix maximum over the lanes, seeded with the carry of the previous chunk
            int e = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(FULL, e, o);
                if (lane >= o) e = max(e, up);
            }
            e = max(e, carry + (x0 - 1) * D);
            const int c = e - x * D;
            if (on) cur[j] = c;
            // the row maximum and its last position
            int m = on ? c : INT32_MIN, arg = on ? x : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int om = __shfl_xor_sync(FULL, m, o), oa = __shfl_xor_sync(FULL, arg, o);
                if (om > m || (om == m && oa > arg)) {
                    m = om;
                    arg = oa;
                }
            }
            if (m >= rowmax) {
                rowmax = m;
                rowarg = arg;
            }
            const int lastx = min(x0 + 31, W - 1);
            carry = __shfl_sync(FULL, c, lastx - x0);
        }
        if (W > 1 && rowmax >= max_score) {
            max_score = rowmax;
            max_i = FWD ? i + 1 : i;
            max_j = first + step * rowarg;
        }
        // the cell just past the band, then the x-drop slope
        int edge = first + step * (W - 1);                   // last column of the band in scan order
        int jn = edge + step;
        if (FWD ? jn <= tl : jn >= 0) {
            const int tpos = FWD ? jn - 1 : jn;
            int score = prev[edge] + ((qv && pr.target[tpos] == qi) ? MC : 0);
            score = max(score, carry) - D;
            if (score > max_score) {
                max_score = score;
                max_i = FWD ? i + 1 : i;
                max_j = jn;
            }
            // cells jn +- k get score - k D while that stays >= max_score - xf (and inside the target)
            const int room = FWD ? tl - jn : jn;
            int more = 0;
            if (score - D >= max_score - xf) more = min(room, (score - (max_score - xf)) / D);
            __syncwarp();
            for (int k = lane; k <= more; k += 32) cur[jn + step * k] = score - k * D;
            edge = jn + step * more;
        }
        if (FWD)
            high = edge;
        else
            low = edge;
        __syncwarp();
        // trim both ends of the band below max_score - xf
        const int thr = max_score - xf;
        for (;;) {
            if (low > high) break;
            const int j = low + lane;
            const unsigned keep = __ballot_sync(FULL, j <= high && cur[j] >= thr);
            if (keep) {
                low += __ffs(keep) - 1;
                break;
            }
            low += 32;
        }
        if (low > high) low = high + 1;   // (the reference stops exactly one past)
        for (;;) {
            if (low > high) break;
            const int j = high - lane;
            const unsigned keep = __ballot_sync(FULL, j >= low && cur[j] >= thr);
            if (keep) {
                high -= __ffs(keep) - 1;
                break;
            }
            high -= 32;
        }
        if (high < low) high = low - 1;
        if (!FWD && i == bottom) xfactor = P.block_cost;
        const int ii = FWD ? i + 1 : i;
        if (ii - low > max_right) max_right = ii - low;
        if (ii - high < max_left) max_left = ii - high;
    }
    __syncwarp();
    out.a = max_j;
    out.b = max_i;
    out.lo = max_left;
    out.hi = max_right;
    out.score = max_score;
}

// One warp per problem: AlignTraps (align.go:62-86) with alignRecursion (kernel.go:52-129) on an explicit
// stack (the high sub-trapezoid is pushed first, so the low one is processed first, as the recursion does).
__global__ void __launch_bounds__(128) ba_pals_dp_kernel(PalsParams P)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(P.counter, 1);
        p = __shfl_sync(FULL, p, 0);
        if (p >= P.n) break;
        PalsProb pr;
        pr.target = P.letters + P.t_off[p];
        pr.query = P.letters + P.q_off[p];
        pr.tl = P.t_len[p];
        pr.ql = P.q_len[p];
        pr.v0 = P.vec + P.vec_off[p];
        pr.v1 = pr.v0 + (pr.tl + 2);
        const int64_t tr0 = P.trap_off[p];
        const int ntr = (int)(P.trap_off[p + 1] - tr0);
        const int32_t *traps = P.traps + 4 * tr0;
        uint8_t *covered = P.covered + tr0;
        int32_t *stack = P.stack + 4 * P.stack_off[p];
        const int scap = P.stack_cap[p];
        PalsHitDev *hits = P.hits + P.hit_off[p];
        const int hcap = P.hit_cap[p];
        int nhit = 0, status = 0;
        for (int slot = 0; slot < ntr; slot++) {
            __syncwarp();
            if (covered[slot]) continue;
            if (traps[4 * slot + 0] - traps[4 * slot + 1] < P.kmer) continue;
            int sp = 0;
            if (lane == 0) {
                stack[0] = traps[4 * slot + 0];
                stack[1] = traps[4 * slot + 1];
                stack[2] = traps[4 * slot + 2];
                stack[3] = traps[4 * slot + 3];
            }
            sp = 1;
            __syncwarp();
            while (sp > 0) {
                sp--;
                const int top = stack[4 * sp + 0], bot = stack[4 * sp + 1], left = stack[4 * sp + 2], right = stack[4 * sp + 3];
                __syncwarp();
                const int mid = (bot + top) / 2;
                PalsEnd lo_end, hi_end;
                pals_trace<true>(P, pr, mid, mid - right, mid - left, 0, 0, lo_end);
                for (int x = 1; x == 1 || (hi_end.b > mid + x * P.max_igap && hi_end.score < lo_end.score); x++)
                    pals_trace<false>(P, pr, lo_end.b, lo_end.a, lo_end.a, mid + P.max_igap, P.block_cost + 2 * x * P.diff_cost,
                                      hi_end);
                const int abpos = hi_end.a, bbpos = hi_end.b, aepos = lo_end.a, bepos = lo_end.b;
                const int low_top = bbpos - P.max_igap, high_bottom = bepos + P.max_igap;
                if (bepos - bbpos >= P.min_len && aepos - abpos >= P.min_len) {
                    int indel = (abpos - bbpos) - (aepos - bepos);
                    if (indel < 0) indel = -indel;
                    // identity := (1/RMatchCost) - float64(Score-indel)/(RMatchCost*float64(Bepos-Bbpos))
                    const double identity =
                        ba_dsub(ba_ddiv(1.0, P.rmatch_cost),
                                ba_ddiv((double)(hi_end.score - indel), ba_dmul(P.rmatch_cost, (double)(bepos - bbpos))));
                    if (identity <= P.max_diff) {
                        // kernel.go:81-118: the loop index runs over trapezoids[slot+1:] but indexes `covered` from 0
                        if (lane == 0) {
                            for (int i = 0; slot + 1 + i < ntr; i++) {
                                const int32_t *tq = traps + 4 * (slot + 1 + i);
                                const int ttop = tq[0], tbot = tq[1], tleft = tq[2], tright = tq[3];
                                if (tbot >= bepos) break;
                                const int trap_b = ttop - tbot + 1, trap_a = tright - tleft + 1;
                                int cov_a = tleft < hi_end.lo ? hi_end.lo : tleft;
                                int cov_b = tright > hi_end.hi ? hi_end.hi : tright;
                                if (cov_a > cov_b) continue;
                                cov_a = cov_b - cov_a + 1;
                                cov_b = ttop > bepos ? bepos - tbot + 1 : trap_b;
                                if (ba_dmul(ba_ddiv((double)cov_a, (double)trap_a), ba_ddiv((double)cov_b, (double)trap_b)) > 0.99)
                                    covered[i] = 1;
                            }
                            if (nhit < hcap) {
                                PalsHitDev h;
                                h.abpos = abpos;
                                h.bbpos = bbpos;
                                h.aepos = aepos;
                                h.bepos = bepos;
                                h.low_diagonal = -hi_end.hi;   // diagonals to this point are query-target, not target-query
                                h.high_diagonal = -hi_end.lo;
                                h.score = hi_end.score;
                                h.pad = 0;
                                h.error = identity;
                                hits[nhit] = h;
                            }
                        }
                        if (nhit >= hcap) status = 1;
                        nhit++;
                    }
                }
                // recursion: low part first, then high part  =>  push high, then low
                const bool go_low = (low_top - bot > P.min_len) && (low_top < top - P.max_igap);
                const bool go_high = (top - high_bottom > P.min_len);
                if (sp + (go_low ? 1 : 0) + (go_high ? 1 : 0) > scap) {
                    status = 2;
                    sp = 0;
                    break;
                }
                if (go_high) {
                    if (lane == 0) {
                        stack[4 * sp + 0] = top;
                        stack[4 * sp + 1] = high_bottom;
                        stack[4 * sp + 2] = left;
                        stack[4 * sp + 3] = right;
                    }
                    sp++;
                }
                if (go_low) {
                    if (lane == 0) {
                        stack[4 * sp + 0] = low_top;
                        stack[4 * sp + 1] = bot;
                        stack[4 * sp + 2] = left;
                        stack[4 * sp + 3] = right;
                    }
                    sp++;
                }
                __syncwarp();
            }
            if (status == 2) break;
        }
        if (lane == 0) {
            P.hit_count[p] = nhit;
            P.status[p] = status;
        }
        __syncwarp();
    }
}
