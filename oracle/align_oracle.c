/*
 * align_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see align_oracle.h).
 *
 * Literal CPU restatement of biogo's align package hot path: the full r*c
 * table is allocated and zeroed per call (8 B/cell linear, 24 B/cell affine),
 * filled row-major, then traced back with the reference's case order.
 * Every function cites the reference lines it follows (paths relative to the
 * reference root).  Go's `int` is int64 here; additions that may involve
 * minInt wrap exactly like Go's (two's complement).
 */
#include "align_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MININT INT64_MIN /* align/align.go:54 */

enum { DIAG = 0, UP = 1, LEFT = 2 }; /* align/align.go:47-51 */

/* Go integer addition wraps; C signed overflow is UB, so go through uint64. */
static inline int64_t wadd(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int64_t wsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }

/* align/align.go:75-83 */
static inline int64_t max3(int64_t a, int64_t b, int64_t c)
{
    if (b > a) a = b;
    if (c > a) return c;
    return a;
}
/* align/align.go:85-90 */
static inline int64_t max2(int64_t a, int64_t b) { return a > b ? a : b; }
/* align/align.go:92-97 */
static inline int64_t add(int64_t a, int64_t b)
{
    if (a == MININT || b == MININT) return MININT;
    return wadd(a, b);
}

typedef struct {
    orc_seg *v;
    int64_t n, cap;
} seglist;

static int seg_push(seglist *s, int64_t as, int64_t ae, int64_t bs, int64_t be, int64_t score)
{
    if (s->n == s->cap) {
        int64_t nc = s->cap ? s->cap * 2 : 8;
        orc_seg *nv = (orc_seg *)realloc(s->v, (size_t)nc * sizeof(orc_seg));
        if (!nv) return -1;
        s->v = nv;
        s->cap = nc;
    }
    s->v[s->n].a_start = as;
    s->v[s->n].a_end = ae;
    s->v[s->n].b_start = bs;
    s->v[s->n].b_end = be;
    s->v[s->n].score = score;
    s->n++;
    return 0;
}

static void seg_reverse(seglist *s)
{
    for (int64_t i = 0, j = s->n - 1; i < j; i++, j--) {
        orc_seg t = s->v[i];
        s->v[i] = s->v[j];
        s->v[j] = t;
    }
}

typedef struct {
    const orc_scoring *sc;
    const uint8_t *ref, *qry;
    int64_t nref, nqry;
    int stride;
    int64_t err_pos;
} ctx_t;

#define RL(k) (cx->ref[(int64_t)(k) * cx->stride]) /* rSeq[k]   (or rSeq[k].L) */
#define QL(k) (cx->qry[(int64_t)(k) * cx->stride]) /* qSeq[k]   (or qSeq[k].L) */
#define IDX(l) ((int64_t)cx->sc->index[(uint8_t)(l)])
/* la[x] with Go's bounds check -> runtime panic */
#define LA_CHECK(x)                                   \
    do {                                              \
        if ((x) < 0 || (x) >= let * let) goto panic_; \
    } while (0)

#define EMIT()                                                         \
    do {                                                               \
        if (seg_push(out, i, maxI, j, maxJ, score)) goto panic_;       \
        maxI = i;                                                      \
        maxJ = j;                                                      \
        score = 0;                                                     \
    } while (0)

/* ------------------------------------------------------------------ */
/* linear gaps: align/sw_letters.go:68-193, align/nw_letters.go:75-196,
 * align/fitted_letters.go:70-201 (and the *_qletters.go twins)         */
/* ------------------------------------------------------------------ */
static int align_linear(ctx_t *cx, seglist *out)
{
    const int kind = cx->sc->kind;
    const int64_t let = cx->sc->let;
    const int64_t *la = cx->sc->la;
    const int64_t r = cx->nref + 1, c = cx->nqry + 1;
    int status = ORC_OK;
    int64_t *table = (int64_t *)calloc((size_t)(r * c), sizeof(int64_t)); /* make([]int, r*c) */
    if (!table) return ORC_PANIC;

    int64_t maxS = 0, maxI = 0, maxJ = 0, score = 0;
    int64_t i, j;

    if (kind != ORC_SW) {
        /* nw_letters.go:91-93, fitted_letters.go:83-85: row 0 */
        for (j = 0; j + 1 < c; j++) {
            int64_t x = IDX(QL(j));
            LA_CHECK(x);
            table[j + 1] = wadd(table[j], la[x]);
        }
    }
    if (kind == ORC_NW) {
        /* nw_letters.go:94-96: column 0 */
        for (i = 1; i < r; i++) {
            int64_t x = IDX(RL(i - 1)) * let;
            LA_CHECK(x);
            table[i * c] = wadd(table[(i - 1) * c], la[x]);
        }
    }

    /* fill: sw_letters.go:92-120, nw_letters.go:98-118, fitted_letters.go:87-107 */
    for (i = 1; i < r; i++) {
        for (j = 1; j < c; j++) {
            int64_t rVal = IDX(RL(i - 1)), qVal = IDX(QL(j - 1));
            if (rVal < 0) {
                cx->err_pos = i - 1;
                status = ORC_ILLEGAL_REF;
                goto done;
            }
            if (qVal < 0) {
                cx->err_pos = j - 1;
                status = ORC_ILLEGAL_QRY;
                goto done;
            }
            int64_t p = i * c + j;
            LA_CHECK(rVal * let + qVal);
            LA_CHECK(rVal * let);
            LA_CHECK(qVal);
            int64_t diagScore = wadd(table[p - c - 1], la[rVal * let + qVal]);
            int64_t upScore = wadd(table[p - c], la[rVal * let]);
            int64_t leftScore = wadd(table[p - 1], la[qVal]);
            score = max3(diagScore, upScore, leftScore);
            if (kind == ORC_SW) {
                /* sw_letters.go:110-118 */
                if (score > 0) {
                    if (score >= maxS && score == diagScore) {
                        maxS = score;
                        maxI = i;
                        maxJ = j;
                    }
                } else {
                    score = 0;
                }
            }
            table[p] = score;
        }
    }

    /* traceback start */
    score = 0;
    int last = DIAG;
    if (kind == ORC_SW) {
        i = maxI; /* sw_letters.go:127 */
        j = maxJ;
    } else if (kind == ORC_NW) {
        i = r - 1; /* nw_letters.go:125-126 */
        j = c - 1;
        maxI = i;
        maxJ = j;
    } else {
        /* fitted_letters.go:112-139 */
        int64_t max = MININT, qVal;
        i = 0;
        j = c - 1;
        for (;;) {
            if (j - 1 < 0 || j - 1 >= cx->nqry) goto panic_; /* qSeq[j-1] out of range */
            qVal = IDX(QL(j - 1));
            if (qVal >= 0) break;
            j--;
        }
        for (int64_t y = 1; y < r; y++) {
            int64_t v = table[y * c + c - 1];
            int64_t rVal = IDX(RL(y - 1));
            if (rVal < 0) continue;
            LA_CHECK(rVal * let + qVal);
            if (v >= max && la[rVal * let + qVal] >= 0) {
                i = y;
                max = v;
            }
        }
        maxI = i;
        maxJ = j;
    }

    /* sw_letters.go:128-183, nw_letters.go:127-176, fitted_letters.go:140-188 */
    while (i > 0 && j > 0) {
        int64_t rVal = IDX(RL(i - 1)), qVal = IDX(QL(j - 1));
        int64_t p = i * c + j;
        int guard = (kind == ORC_SW) || (p != r * c - 1);
        LA_CHECK(rVal * let + qVal);
        LA_CHECK(rVal * let);
        LA_CHECK(qVal);
        if (kind == ORC_SW && table[p] == 0) {
            break;
        } else if (table[p] == wadd(table[p - c - 1], la[rVal * let + qVal])) {
            if (last != DIAG) EMIT();
            score = wadd(score, wsub(table[p], table[p - c - 1]));
            i--;
            j--;
            last = DIAG;
        } else if (table[p] == wadd(table[p - c], la[rVal * let])) {
            if (last != UP && guard) EMIT();
            score = wadd(score, wsub(table[p], table[p - c]));
            i--;
            last = UP;
        } else if (table[p] == wadd(table[p - 1], la[qVal])) {
            if (last != LEFT && guard) EMIT();
            score = wadd(score, wsub(table[p], table[p - 1]));
            j--;
            last = LEFT;
        } else {
            status = ORC_NO_PATH;
            goto done;
        }
    }
    if (seg_push(out, i, maxI, j, maxJ, score)) goto panic_;
    if (kind == ORC_NW && i != j) {
        /* nw_letters.go:183-189 */
        if (seg_push(out, 0, i, 0, j, table[i * c + j])) goto panic_;
    }
    seg_reverse(out);
    goto done;

panic_:
    status = ORC_PANIC;
done:
    free(table);
    return status;
}

/* ------------------------------------------------------------------ */
/* affine gaps: align/sw_affine_letters.go:85-292,
 * align/nw_affine_letters.go:98-332, align/fitted_affine_letters.go:93-308 */
/* ------------------------------------------------------------------ */
typedef int64_t cell3[3];

static int align_affine(ctx_t *cx, seglist *out)
{
    const int kind = cx->sc->kind;
    const int64_t let = cx->sc->let;
    const int64_t *la = cx->sc->la;
    const int64_t go = cx->sc->gap_open;
    const int64_t r = cx->nref + 1, c = cx->nqry + 1;
    const int64_t len = r * c;
    int status = ORC_OK;
    cell3 *table = (cell3 *)calloc((size_t)len, sizeof(cell3)); /* make([][3]int, r*c) */
    if (!table) return ORC_PANIC;

    int64_t maxS = 0, maxI = 0, maxJ = 0, score = 0;
    int64_t i, j;

#define TBL_CHECK(x)                        \
    do {                                    \
        if ((x) < 0 || (x) >= len) goto panic_; \
    } while (0)

    if (kind != ORC_SW) {
        /* nw_affine_letters.go:114-142, fitted_affine_letters.go:106-132 */
        table[0][DIAG] = 0;
        table[0][UP] = MININT;
        table[0][LEFT] = MININT;
        TBL_CHECK(1);
        if (cx->nqry < 1) goto panic_; /* qSeq[0] */
        {
            int64_t x = IDX(QL(0));
            LA_CHECK(x);
            table[1][DIAG] = MININT;
            table[1][UP] = MININT;
            table[1][LEFT] = wadd(go, la[x]);
        }
        if (c < 2) goto panic_; /* table[2:c] with c < 2 */
        for (j = 0; j + 2 < c; j++) {
            int64_t x = IDX(QL(j + 1));
            LA_CHECK(x);
            table[j + 2][DIAG] = MININT;
            table[j + 2][UP] = MININT;
            table[j + 2][LEFT] = wadd(table[j + 1][LEFT], la[x]);
        }
        TBL_CHECK(c); /* table[c]: panics when the reference is empty */
        if (kind == ORC_NW) {
            if (cx->nref < 1) goto panic_; /* rSeq[0] */
            int64_t x = IDX(RL(0)) * let;
            LA_CHECK(x);
            table[c][DIAG] = MININT;
            table[c][UP] = wadd(go, la[x]);
            table[c][LEFT] = MININT;
            for (i = 2; i < r; i++) {
                int64_t y = IDX(RL(i - 1)) * let;
                LA_CHECK(y);
                table[i * c][DIAG] = MININT;
                table[i * c][UP] = wadd(table[(i - 1) * c][UP], la[y]);
                table[i * c][LEFT] = MININT;
            }
        } else {
            /* fitted_affine_letters.go:123-132: `up` is omitted, so it is 0 */
            table[c][DIAG] = MININT;
            table[c][UP] = 0;
            table[c][LEFT] = MININT;
            for (i = 2; i < r; i++) {
                table[i * c][DIAG] = MININT;
                table[i * c][UP] = 0;
                table[i * c][LEFT] = MININT;
            }
        }
    }

    /* fill: sw_affine_letters.go:108-157, nw_affine_letters.go:144-174,
     * fitted_affine_letters.go:134-164 */
    for (i = 1; i < r; i++) {
        for (j = 1; j < c; j++) {
            int64_t rVal = IDX(RL(i - 1)), qVal = IDX(QL(j - 1));
            if (rVal < 0) {
                cx->err_pos = i - 1;
                status = ORC_ILLEGAL_REF;
                goto done;
            }
            if (qVal < 0) {
                cx->err_pos = j - 1;
                status = ORC_ILLEGAL_QRY;
                goto done;
            }
            int64_t p = i * c + j;
            LA_CHECK(rVal * let + qVal);
            LA_CHECK(rVal * let);
            LA_CHECK(qVal);
            int64_t diagScore = table[p - c - 1][DIAG];
            int64_t upScore = table[p - c - 1][UP];
            int64_t leftScore = table[p - c - 1][LEFT];
            if (kind == ORC_SW) {
                score = max3(diagScore, upScore, leftScore);
                int matched = score == diagScore;
                score = wadd(score, la[rVal * let + qVal]);
                if (score > 0) {
                    if (score >= maxS && matched) {
                        maxS = score;
                        maxI = i;
                        maxJ = j;
                    }
                } else {
                    score = 0;
                }
                table[p][DIAG] = score;

                score = max2(wadd(wadd(table[p - c][DIAG], go), la[rVal * let]),
                             wadd(table[p - c][UP], la[rVal * let]));
                if (score < 0) score = 0;
                table[p][UP] = score;

                score = max2(wadd(wadd(table[p - 1][DIAG], go), la[qVal]),
                             wadd(table[p - 1][LEFT], la[qVal]));
                if (score < 0) score = 0;
                table[p][LEFT] = score;
            } else {
                table[p][DIAG] = wadd(max3(diagScore, upScore, leftScore), la[rVal * let + qVal]);
                table[p][UP] = max2(add(table[p - c][DIAG], wadd(go, la[rVal * let])),
                                    add(table[p - c][UP], la[rVal * let]));
                table[p][LEFT] = max2(add(table[p - 1][DIAG], wadd(go, la[qVal])),
                                      add(table[p - 1][LEFT], la[qVal]));
            }
        }
    }

    /* traceback start */
    int layer = DIAG, last = DIAG;
    score = 0;
    if (kind == ORC_SW) {
        i = maxI; /* sw_affine_letters.go:163-164 */
        j = maxJ;
    } else if (kind == ORC_NW) {
        /* nw_affine_letters.go:183-192 */
        i = r - 1;
        j = c - 1;
        int64_t best = table[i * c + j][0];
        for (int k = 1; k < 3; k++) {
            if (table[i * c + j][k] > best) {
                best = table[i * c + j][k];
                layer = k;
            }
        }
        maxI = i;
        maxJ = j;
    } else {
        /* fitted_affine_letters.go:170-183 */
        int64_t max = MININT;
        i = 0;
        j = c - 1;
        for (int64_t y = 1; y < r; y++) {
            int64_t v = table[y * c + c - 1][DIAG];
            if (v >= max) {
                i = y;
                max = v;
            }
        }
        maxI = i;
        maxJ = j;
    }

    /* sw_affine_letters.go:166-279, nw_affine_letters.go:193-304,
     * fitted_affine_letters.go:184-295 */
    while (i > 0 && j > 0) {
        int64_t rVal = IDX(RL(i - 1)), qVal = IDX(QL(j - 1));
        int64_t p = i * c + j;
        int64_t cur = table[p][layer];
        int corner = (p == len - 1);
        LA_CHECK(rVal * let + qVal);
        LA_CHECK(rVal * let);
        LA_CHECK(qVal);
        int64_t er = la[rVal * let], eq = la[qVal], sub = la[rVal * let + qVal];
        int move, nl;
        int64_t pred;
        if (kind == ORC_SW && cur == 0) {
            break;
        } else if (cur == wadd(table[p - c][UP], er)) {
            move = UP; nl = UP; pred = table[p - c][UP];
        } else if (cur == wadd(table[p - 1][LEFT], eq)) {
            move = LEFT; nl = LEFT; pred = table[p - 1][LEFT];
        } else if (cur == wadd(wadd(table[p - c][DIAG], go), er)) {
            move = UP; nl = DIAG; pred = table[p - c][DIAG];
        } else if (cur == wadd(wadd(table[p - 1][DIAG], go), eq)) {
            move = LEFT; nl = DIAG; pred = table[p - 1][DIAG];
        } else if (kind == ORC_SW) {
            /* sw_affine_letters.go:227-271: diag layer first */
            if (cur == wadd(table[p - c - 1][DIAG], sub)) {
                move = DIAG; nl = DIAG; pred = table[p - c - 1][DIAG];
            } else if (cur == wadd(table[p - c - 1][UP], sub)) {
                move = DIAG; nl = UP; pred = table[p - c - 1][UP];
            } else if (cur == wadd(table[p - c - 1][LEFT], sub)) {
                move = DIAG; nl = LEFT; pred = table[p - c - 1][LEFT];
            } else {
                status = ORC_NO_PATH;
                goto done;
            }
        } else {
            /* nw_affine_letters.go:255-299, fitted_affine_letters.go:246-290: up, left, diag */
            if (cur == wadd(table[p - c - 1][UP], sub)) {
                move = DIAG; nl = UP; pred = table[p - c - 1][UP];
            } else if (cur == wadd(table[p - c - 1][LEFT], sub)) {
                move = DIAG; nl = LEFT; pred = table[p - c - 1][LEFT];
            } else if (cur == wadd(table[p - c - 1][DIAG], sub)) {
                move = DIAG; nl = DIAG; pred = table[p - c - 1][DIAG];
            } else {
                status = ORC_NO_PATH;
                goto done;
            }
        }
        if (last != move && (move == DIAG || !corner)) EMIT();
        score = wadd(score, wsub(cur, pred));
        if (move != LEFT) i--;
        if (move != UP) j--;
        layer = nl;
        last = move;
    }
    if (seg_push(out, i, maxI, j, maxJ, score)) goto panic_;
    if (kind == ORC_NW && i != j) {
        /* nw_affine_letters.go:311-325 */
        int l2 = (i == 0) ? LEFT : UP;
        if (seg_push(out, 0, i, 0, j, table[i * c + j][l2])) goto panic_;
    }
    seg_reverse(out);
    goto done;

panic_:
    status = ORC_PANIC;
done:
    free(table);
    return status;
#undef TBL_CHECK
}

int orc_align(const orc_scoring *sc, const uint8_t *ref, int64_t nref, const uint8_t *qry,
              int64_t nqry, int stride, orc_seg **segs, int64_t *nsegs, int64_t *err_pos)
{
    ctx_t cx;
    seglist out = {0, 0, 0};
    cx.sc = sc;
    cx.ref = ref;
    cx.qry = qry;
    cx.nref = nref;
    cx.nqry = nqry;
    cx.stride = stride;
    cx.err_pos = -1;
    int st = sc->affine ? align_affine(&cx, &out) : align_linear(&cx, &out);
    if (err_pos) *err_pos = cx.err_pos;
    if (st != ORC_OK) {
        free(out.v);
        out.v = 0;
        out.n = 0;
    }
    *segs = out.v;
    *nsegs = out.n;
    return st;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------ */
/* batch driver (CPU baseline shape: a worker pool over independent pairs) */
/* ------------------------------------------------------------------ */
typedef struct {
    const orc_scoring *sc;
    const uint8_t *letters;
    int stride;
    const int64_t *ref_off, *qry_off;
    const int32_t *ref_len, *qry_len;
    int64_t n;
    int64_t next;
    pthread_mutex_t mu;
    int32_t *status;
    int64_t *err_pos;
    orc_seg **per_segs;
    int64_t *per_n;
} batch_t;

static void *batch_worker(void *arg)
{
    batch_t *b = (batch_t *)arg;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        int64_t lo = b->next;
        b->next += 16;
        pthread_mutex_unlock(&b->mu);
        if (lo >= b->n) break;
        int64_t hi = lo + 16 < b->n ? lo + 16 : b->n;
        for (int64_t k = lo; k < hi; k++) {
            int64_t ep = -1;
            b->status[k] = orc_align(b->sc, b->letters + b->ref_off[k], b->ref_len[k],
                                     b->letters + b->qry_off[k], b->qry_len[k], b->stride,
                                     &b->per_segs[k], &b->per_n[k], &ep);
            b->err_pos[k] = ep;
        }
    }
    return 0;
}

int orc_align_batch(const orc_scoring *sc, const uint8_t *letters, int stride,
                    const int64_t *ref_off, const int32_t *ref_len, const int64_t *qry_off,
                    const int32_t *qry_len, int64_t n, int nthreads, int32_t *status,
                    int64_t *err_pos, int64_t *seg_off, int64_t **segs_out)
{
    batch_t b;
    memset(&b, 0, sizeof(b));
    b.sc = sc;
    b.letters = letters;
    b.stride = stride;
    b.ref_off = ref_off;
    b.ref_len = ref_len;
    b.qry_off = qry_off;
    b.qry_len = qry_len;
    b.n = n;
    b.status = status;
    b.err_pos = err_pos;
    b.per_segs = (orc_seg **)calloc((size_t)(n > 0 ? n : 1), sizeof(orc_seg *));
    b.per_n = (int64_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
    if (!b.per_segs || !b.per_n) return -1;
    pthread_mutex_init(&b.mu, 0);
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], 0, batch_worker, &b);
    batch_worker(&b);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], 0);
    free(th);
    pthread_mutex_destroy(&b.mu);

    int64_t total = 0;
    for (int64_t k = 0; k < n; k++) {
        seg_off[k] = total;
        total += b.per_n[k];
    }
    seg_off[n] = total;
    int64_t *flat = (int64_t *)malloc((size_t)(total > 0 ? total : 1) * 5 * sizeof(int64_t));
    if (!flat) return -1;
    for (int64_t k = 0; k < n; k++) {
        for (int64_t s = 0; s < b.per_n[k]; s++) {
            int64_t *d = flat + (seg_off[k] + s) * 5;
            d[0] = b.per_segs[k][s].a_start;
            d[1] = b.per_segs[k][s].a_end;
            d[2] = b.per_segs[k][s].b_start;
            d[3] = b.per_segs[k][s].b_end;
            d[4] = b.per_segs[k][s].score;
        }
        free(b.per_segs[k]);
    }
    free(b.per_segs);
    free(b.per_n);
    *segs_out = flat;
    return 0;
}

/* ------------------------------------------------------------------ */
/* helpers on the boundary: alphabet index, featPair.String, align.Format */
/* ------------------------------------------------------------------ */
/* alphabet/alphabet.go:183-219 */
void orc_build_index(const char *letters, int case_sensitive, int32_t index[256])
{
    int n = (int)strlen(letters);
    for (int i = 0; i < 256; i++) index[i] = -1;
    if (case_sensitive) {
        for (int i = 0; i < n; i++) index[(uint8_t)letters[i]] = i;
        return;
    }
    /* a.letters = ToLower(letters) + ToUpper(letters) */
    for (int i = 0; i < n; i++) {
        char l = letters[i];
        if (l >= 'A' && l <= 'Z') l = (char)(l - 'A' + 'a');
        index[(uint8_t)l] = i;
    }
    for (int i = 0; i < n; i++) {
        char lo = letters[i], up = letters[i];
        if (lo >= 'A' && lo <= 'Z') lo = (char)(lo - 'A' + 'a');
        if (up >= 'a' && up <= 'z') up = (char)(up - 'a' + 'A');
        index[(uint8_t)up] = index[(uint8_t)lo];
    }
}

/* align/align.go:129-144, rendered the way fmt.Sprint renders a []feat.Pair */
int orc_format_pairs(const orc_seg *s, int64_t n, char *buf, int64_t buflen)
{
    int64_t o = 0;
#define PUT(...)                                                         \
    do {                                                                 \
        int w_ = snprintf(buf + o, (size_t)(buflen - o), __VA_ARGS__);   \
        if (w_ < 0 || w_ >= buflen - o) return -1;                       \
        o += w_;                                                         \
    } while (0)
    PUT("[");
    for (int64_t k = 0; k < n; k++) {
        if (k) PUT(" ");
        if (s[k].a_start == s[k].a_end)
            PUT("-/[%lld,%lld)=%lld", (long long)s[k].b_start, (long long)s[k].b_end,
                (long long)s[k].score);
        else if (s[k].b_start == s[k].b_end)
            PUT("[%lld,%lld)/-=%lld", (long long)s[k].a_start, (long long)s[k].a_end,
                (long long)s[k].score);
        else
            PUT("[%lld,%lld)/[%lld,%lld)=%lld", (long long)s[k].a_start, (long long)s[k].a_end,
                (long long)s[k].b_start, (long long)s[k].b_end, (long long)s[k].score);
    }
    PUT("]");
#undef PUT
    return (int)o;
}

/* align/align.go:148-170 */
int orc_format_alignment(const uint8_t *ref, const uint8_t *qry, int stride, const orc_seg *s,
                         int64_t n, uint8_t gap, uint8_t *out_a, uint8_t *out_b, int64_t cap,
                         int64_t *out_len)
{
    int64_t la_ = 0, lb_ = 0;
    for (int64_t k = 0; k < n; k++) {
        int64_t alen = s[k].a_end - s[k].a_start, blen = s[k].b_end - s[k].b_start;
        if (alen == 0) {
            if (la_ + blen > cap) return -1;
            for (int64_t x = 0; x < blen; x++) out_a[la_++] = gap;
        } else {
            if (la_ + alen > cap) return -1;
            for (int64_t x = s[k].a_start; x < s[k].a_end; x++) out_a[la_++] = ref[x * stride];
        }
        if (blen == 0) {
            if (lb_ + alen > cap) return -1;
            for (int64_t x = 0; x < alen; x++) out_b[lb_++] = gap;
        } else {
            if (lb_ + blen > cap) return -1;
            for (int64_t x = s[k].b_start; x < s[k].b_end; x++) out_b[lb_++] = qry[x * stride];
        }
    }
    out_len[0] = la_;
    out_len[1] = lb_;
    return 0;
}
