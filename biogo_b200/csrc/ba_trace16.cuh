// ba_trace16.cuh -- traceback for the checkpointed fill (ba_fill16.cuh).
//
// One thread per pair.  The walk needs bíogo's layer-independent candidate order at every
// visited cell (align/sw_affine_letters.go:171-275), i.e. exact VALUES around the path.  The
// fill stored only sparse checkpoints, so this kernel re-derives the table tile by tile:
// a tile is the part of one column strip between two of its row checkpoints (16 rows x
// COLS columns).  For the tile under the walk it rebuilds the boundary (top row from the row
// checkpoint, left column from the column checkpoints, the missing third layer by its own
// one-dimensional recurrence), recomputes the needed cells with the very same
// affine_cell<KIND>() that ba_fill32 uses, keeps their 8-bit codes in shared memory and walks
// them with the same rules as ba_walk_kernel until the path leaves the tile.
#pragma once

#include "ba_common.cuh"
#include "ba_fill16.cuh"
#include "ba_fill32.cuh"
#include "ba_walk.cuh"

#define BA_TRACE_THREADS 128
#ifndef BA_TRACE_MINB
#define BA_TRACE_MINB 4   // resident CTAs per SM the register allocation is tuned for
#endif
#define BA_TILE_ROWS BA16_PERIOD

struct ba_true_c { static constexpr bool value = true; };
struct ba_false_c { static constexpr bool value = false; };

struct Trace16Params {
    const uint8_t *enc;
    int stride;
    const int64_t *ref_off;
    const int32_t *ref_len;
    const int64_t *qry_off;
    const int32_t *qry_len;
    int32_t *status;
    const uint8_t *acgt_only;  // per pair letter flags (see ba_fill16)
    int flag_mask;
    const WorkItem *items;
    int n_items;
    const uint8_t *arena;
    const DeviceScoring *sc;
    int32_t *seg_count;        // per item: segments emitted (may exceed the slot capacity); -1 = not eligible
    const int64_t *slot_off;   // per item: first slot, in segments
    const int32_t *slot_cap;   // per item
    int32_t *slots;            // 5 ints per segment, in EMISSION order (reverse of final)
    int *counter;              // ba_trace16p_kernel: next work item
    int row_bias;              // go + e when the fill ran the SYM kernels (their row checkpoints hold M + go + e), else 0
    int val_bias;              // vb of the biased SYM SW kernels (every checkpointed value carries + vb), else 0
};

// One pair's view of its couple's checkpoint region (layout: ba_fill16.cuh "Checkpoint geometry").
struct CkView {
    const uint8_t *base;
    int n, m, cols, npass, jshift;   // the PAIR's own shape: rows, columns, passes, column shift (SW right-aligns)
    int n_ck, np_ck;                 // the region's: rows and passes it is laid out for (max over the couple)
    int half;                        // 0: my values are the low 16-bit halves of the stored words, 1: the high halves
    int rs;                          // double-elements per lane row of a row block
    int bias;                        // SYM kernels store G = M + go + e in the row blocks: M = G - bias
    int vb;                          // biased SYM SW kernels: every stored value carries + vb
    bool lin;    // linear kinds: one table T (elements of one word)
    bool clamp;  // affine SW: the fill defers the clamp of U and L at zero and stores them raw
    unsigned inv_cols;         // ceil(2^32 / cols): strip_of without a division
    int64_t pass_bytes, colck_bytes;
    __device__ __forceinline__ void init(const uint8_t *arena, const WorkItem &item, int n_, int m_, bool lin_, bool clamp_,
                                         int bias_, int vb_)
    {
        vb = vb_;
        base = arena + item.code_off;
        n = n_;
        m = m_;
        cols = item.cols;
        lin = lin_;
        clamp = clamp_;
        bias = bias_;
        n_ck = item.n_ck;
        np_ck = item.ck & 0xffff;
        half = (item.ck >> 16) & 1;
        rs = ba16_row_stride(cols);
        inv_cols = cols > 1 ? (unsigned)(0x100000000ull / (unsigned)cols) + 1u : 0u;
        pass_bytes = ba16_pass_bytes(cols, n_ck, lin);
        colck_bytes = ba16_colck_bytes(n_ck, lin);
    }
    // Strip g (global index: pass*32 + strip) belongs to lane (g & 31) >> 1 of the fill, which
    // handles row i at step i-1+lane.
    __device__ __forceinline__ static int lane_of(int g) { return (g & 31) >> 1; }
    __device__ __forceinline__ int pick(unsigned w) const { return (int)(short)(half ? (w >> 16) : (w & 0xffffu)); }
    // Raw checkpoint words and their decoding are separate steps: a tile requests everything its boundary needs
    // in one batch and decodes afterwards, so the loads' latencies overlap instead of adding up (ncu r02: 45 % of
    // the traceback's stall samples sat on the first use of a checkpoint word).
    // {M, L} words of the last column of strip g at row r
    __device__ __forceinline__ uint2 col_raw(int g, int r) const
    {
        const int pass = g >> 5, t = lane_of(g), s = g & 1;
        const int st = r - 1 + t;
        const int e = st >> BA16_PSHIFT, j = st & (BA16_PERIOD - 1);
        const uint8_t *blk = base + (int64_t)pass * pass_bytes + (int64_t)e * ba16_colblk(lin);
        // linear: [lane][step][strip] words; affine: [lane][strip][step] {M, L}
        if (lin) return make_uint2(*(const unsigned *)(blk + t * 64 + j * 8 + s * 4), 0u);
        return *(const uint2 *)(blk + t * 128 + s * 64 + j * 8);
    }
    __device__ __forceinline__ void col_pick(const uint2 &w, int &M, int &L) const
    {
        if (lin) {
            M = pick(w.x);
            L = 0;
            return;
        }
        M = pick(w.x) - vb;
        L = pick(w.y) - vb;
        if (clamp) L = max(L, 0);
    }
    // (M, clamped L) of the last column of strip g at row r
    __device__ __forceinline__ void col_ml(int g, int r, int &M, int &L) const { col_pick(col_raw(g, r), M, L); }
    // {M or G, U} words of column k of strip g at its checkpoint row of period e
    __device__ __forceinline__ uint2 row_raw(int g, int e, int k) const
    {
        const int pass = g >> 5, t = lane_of(g), s = g & 1;
        const int kk = s * cols + k;
        const uint8_t *blk = base + (int64_t)pass * pass_bytes + colck_bytes + (int64_t)e * ba16_rowblk(cols, lin);
        if (lin) return make_uint2(*(const unsigned *)(blk + (t * rs * 2 + kk) * 4), 0u);
        return *(const uint2 *)(blk + (t * rs * 2 + kk) * 8);
    }
    __device__ __forceinline__ void row_pick(const uint2 &w, int &M, int &U) const
    {
        if (lin) {
            M = pick(w.x);
            U = 0;
            return;
        }
        M = pick(w.x) - bias - vb;
        U = pick(w.y) - vb;
        if (clamp) U = max(U, 0);
    }
    // (M, clamped U) of column k of strip g at its checkpoint row of period e
    __device__ __forceinline__ void row_mu(int g, int e, int k, int &M, int &U) const { row_pick(row_raw(g, e, k), M, U); }
    // start-cell candidate record x = pass*16 + lane of this pair: {best score, last period attaining it}
    __device__ __forceinline__ int2 cand(int x) const
    {
        return ((const int2 *)(base + (int64_t)np_ck * pass_bytes))[2 * x + half];
    }
    // (j - jshift - 1) / cols for j >= 1 (the numerator is non-negative and below 2^20: exact with a 32-bit reciprocal)
    __device__ __forceinline__ int strip_of(int j) const
    {
        const unsigned x = (unsigned)max(j - jshift - 1, 0);
        return cols > 1 ? (int)__umulhi(x, inv_cols) : (int)x;
    }
    __device__ __forceinline__ int left_col(int g) const { return jshift + g * cols; }
};

// What the caller wants to learn from a tile besides its codes.
struct TileProbe {
    int want = -1;              // SW: report the last cell (row-major) whose M equals this scaled value
    int fi = 0, fj = 0;
    int col = -1;               // Fitted: track the last row maximising M in this table column
    int best = INT32_MIN, besti = 0;
    int ci = -1, cj = -1;       // NW: capture the three layer values of this cell
    int cM = 0, cU = 0, cL = 0;
};

// DP border values in the scaled domain (uniform gap entries: condition C1 of ba_fill16.cuh).
template <int KIND>
__device__ __forceinline__ Cell3 border_row0(int j, int go, int eq)
{
    Cell3 c;
    if (KIND == KIND_SW) {
        c.M = c.U = c.L = 0;
    } else if (j == 0) {
        c.M = 0;                // table[0] = {0, minInt, minInt}
        c.U = c.L = BA_NEG;
    } else {
        c.M = c.U = BA_NEG;     // nw_affine_letters.go:119-131
        c.L = go + j * eq;
    }
    return c;
}
template <int KIND>
__device__ __forceinline__ Cell3 border_col0(int i, int go, int er)
{
    Cell3 c;
    if (KIND == KIND_SW) {
        c.M = c.U = c.L = 0;
    } else {
        c.M = c.L = BA_NEG;     // nw_affine_letters.go:132-142 / fitted_affine_letters.go:123-132
        c.U = (KIND == KIND_NW) ? go + i * er : 0;
    }
    return c;
}

// Recompute the tile of strip g, period e, for rows rt+1 .. imax and columns up to jmax, and
// store the cell codes in `tile` (byte (row*8 + k) * BA_TRACE_THREADS apart).
// S is the scoring matrix scaled by BA_SCALE in shared memory (row stride `let`).
template <int KIND, int TS, int MC>
__device__ __forceinline__ void trace16_tile_aff(const CkView &ck, const int *S, int let, int go, const uint8_t *rs,
                                             const uint8_t *qs, int stride, int g, int e, int imax, int jmax,
                                             uint8_t *tile, TileProbe &pb)
{
    const int v = CkView::lane_of(g);
    const int rt = max(BA16_PERIOD * e - v, 0); // top boundary row (0 = DP border)
    const int c0 = ck.left_col(g);
    const int cL = max(c0, 0);                  // effective left boundary column (0 = DP border)
    const int kfirst = cL - c0;                 // first real column of the strip
    const int klast = jmax - c0 - 1;            // last needed column (inclusive)
    const int er_u = S[let], eq_u = S[1];       // uniform gap entries (letters >= 1)

    const int gl = g - 1;                       // strip whose last column is cL
    const bool hasL = cL > 0;                   // the left boundary is a checkpointed column (else the DP border)

    // ---- request everything the boundary needs in ONE batch (decoded below): query letters, the top row's
    // checkpoints, the left column at the corner and at the first row, and the rows U's recurrence starts from
    unsigned qraw[MC];
    uint2 rowraw[MC];
#pragma unroll
    for (int k = 0; k < MC; k++) {
        const bool on = k >= kfirst && k <= klast;
        qraw[k] = on ? (unsigned)qs[(c0 + k) * stride] : 0u;
        rowraw[k] = (on && rt >= 1) ? ck.row_raw(g, e - 1, k) : make_uint2(0u, 0u);
    }
    const uint2 cornraw = (hasL && rt >= 1) ? ck.col_raw(gl, rt) : make_uint2(0u, 0u);
    uint2 nraw = (hasL && rt + 1 <= imax) ? ck.col_raw(gl, rt + 1) : make_uint2(0u, 0u);   // row loop: one row ahead
    int nR = (rt + 1 <= imax) ? (int)rs[(int64_t)rt * stride] : 0;
    // U of column cL is checkpointed (with M) only at strip gl's own rows; start from the closest one at or
    // above rt (r0) and run U's recurrence down the column on the stored M of rows r0 .. rt-1 (at most
    // BA16_PERIOD - 1 of them).
    constexpr int NCH = BA16_PERIOD - 1;
    int r0 = 0;
    uint2 u0raw = make_uint2(0u, 0u), chraw[NCH];
#pragma unroll
    for (int x = 0; x < NCH; x++) chraw[x] = make_uint2(0u, 0u);
    if (hasL) {
        const int vl = CkView::lane_of(gl);
        r0 = rt - ((rt + vl) & (BA16_PERIOD - 1));  // largest row <= rt with (r0 + vl) % PERIOD == 0
        if (r0 >= 1)
            u0raw = ck.row_raw(gl, (r0 + vl) / BA16_PERIOD - 1, ck.cols - 1);
        else
            r0 = 0;                             // DP border row
#pragma unroll
        for (int x = 0; x < NCH; x++) {
            const int rr = r0 + x;
            if (rr < rt && rr >= 1) chraw[x] = ck.col_raw(gl, rr);
        }
    }

    // ---- per-column constants of this tile
    int qo[MC], eqv[MC];
#pragma unroll
    for (int k = 0; k < MC; k++) {
        qo[k] = 0;
        eqv[k] = 0;
        if (k >= kfirst && k <= klast) {
            qo[k] = (int)qraw[k];
            eqv[k] = S[qo[k]];
        }
    }

    // ---- corner (rt, cL) and the U layer of the left boundary column
    Cell3 corner = (rt == 0) ? border_row0<KIND>(cL, go, eq_u) : border_col0<KIND>(rt, go, er_u);
    int leftU = corner.U;                       // U at (row, cL), advanced row by row
    if (hasL) {
        int u;
        if (r0 >= 1) {
            int mm;
            ck.row_pick(u0raw, mm, u);
            u *= BA_SCALE;
        } else {
            u = border_row0<KIND>(cL, go, eq_u).U;
        }
        // rows r0+1 .. rt: U[r] = max(M[r-1] + go + er, U[r-1] + er [, 0]); the gap entry is the same for every
        // letter this path admits (condition C1 of ba_fill16.cuh), so no letter is looked up here
#pragma unroll
        for (int x = 0; x < NCH; x++) {
            const int rr = r0 + x;              // the row whose M feeds row rr + 1
            if (rr < rt) {
                int mp, lp;
                if (rr >= 1) {
                    ck.col_pick(chraw[x], mp, lp);
                    mp *= BA_SCALE;
                } else {
                    mp = border_row0<KIND>(cL, go, eq_u).M;
                }
                u = max(mp + go + er_u, u + er_u);
                if (KIND == KIND_SW) u = max(u, 0);
            }
        }
        leftU = u;
        if (rt >= 1) {
            int mm, ll;
            ck.col_pick(cornraw, mm, ll);
            corner.M = mm * BA_SCALE;
            corner.L = ll * BA_SCALE;
            corner.U = u;
        }
    }

    // ---- top boundary row rt for my columns
    Cell3 Pv[MC];
    {
        int lM = corner.M, lL = corner.L;
#pragma unroll
        for (int k = 0; k < MC; k++) {
            Pv[k].M = Pv[k].U = Pv[k].L = 0;
            if (k >= kfirst && k <= klast) {
                if (rt >= 1) {
                    int mm, uu;
                    ck.row_pick(rowraw[k], mm, uu);
                    Pv[k].M = mm * BA_SCALE;
                    Pv[k].U = uu * BA_SCALE;
                    Pv[k].L = max(lM + go + eqv[k], lL + eqv[k]);
                    if (KIND == KIND_SW) Pv[k].L = max(Pv[k].L, 0);
                    lM = Pv[k].M;
                    lL = Pv[k].L;
                } else {
                    Pv[k] = border_row0<KIND>(c0 + k + 1, go, eq_u);
                }
            }
        }
    }

    // ---- rows rt+1 .. imax
    // Two forms of the row loop.  ALL: the strip is as wide as the tile arrays and starts at a real column, so
    // every cell of a row is computed without a test -- cells right of jmax are never read (values flow left to
    // right, the walk stays up-left of (imax, jmax)), and leaving out the per-cell branch region saves a fifth of
    // a cell's instructions.  Otherwise (narrower strips, SW's padding columns) each cell is tested.
    auto rows = [&](auto all_c) {
        constexpr bool ALL = decltype(all_c)::value;
        Cell3 dprev = corner;  // (row-1, cL)
        uint8_t *trow = tile;
        // the next row's letter and left-boundary checkpoint are requested one row ahead and decoded when their
        // row comes, so their latency hides under the current row's cells
        for (int r = rt + 1; r <= imax; r++, trow += MC * TS) {
            const int *Srow = S + nR * let;
            int mm = 0, ll = 0;
            if (hasL) ck.col_pick(nraw, mm, ll);
            if (r + 1 <= imax) {
                nR = (int)rs[(int64_t)r * stride];
                if (hasL) nraw = ck.col_raw(gl, r + 1);
            }
            const int er = Srow[0];
            const int goer = go + er;
            Cell3 left;
            if (hasL) {
                // U[r][cL] = max(M[r-1][cL] + go + er, U[r-1][cL] + er [, 0])
                leftU = max(dprev.M + goer, leftU + er);
                if (KIND == KIND_SW) leftU = max(leftU, 0);
                left.M = mm * BA_SCALE;
                left.L = ll * BA_SCALE;
                left.U = leftU;
            } else {
                left = border_col0<KIND>(r, go, er_u);
            }
            Cell3 d = dprev, l = left;
#pragma unroll
            for (int k = 0; k < MC; k++) {
                if (ALL || (k >= kfirst && k <= klast)) {
                    unsigned code;
                    int dm;
                    const Cell3 old = Pv[k];
                    const Cell3 nw = affine_cell<KIND>(d, old, l, Srow[qo[k]], er, goer, eqv[k], go + eqv[k], code, dm);
                    trow[k * TS] = (uint8_t)code;
                    const int col = c0 + k + 1;
                    const bool in = !ALL || k <= klast;   // the probes look at real cells only
                    if (KIND == KIND_SW && pb.want >= 0 && nw.M == pb.want && in) {
                        pb.fi = r;
                        pb.fj = col;
                    }
                    if (KIND == KIND_FIT && col == pb.col && nw.M >= pb.best && in) {
                        pb.best = nw.M;
                        pb.besti = r;
                    }
                    if (KIND == KIND_NW && r == pb.ci && col == pb.cj && in) {
                        pb.cM = nw.M;
                        pb.cU = nw.U;
                        pb.cL = nw.L;
                    }
                    Pv[k] = nw;
                    d = old;
                    l = nw;
                }
            }
            dprev = left;
        }
    };
    if (kfirst == 0 && ck.cols == MC)
        rows(ba_true_c{});
    else
        rows(ba_false_c{});
}

// Linear-gap tile (ba_fill16 LIN): one table T; the checkpoints hold it in their M field.  Codes
// are linear_cell<KIND>'s: 3 diag, 2 up, 1 left, 0 = SW cell clamped to zero (stop).
template <int KIND, int TS, int MC>
__device__ __forceinline__ void trace16_tile_lin(const CkView &ck, const int *S, int let, const uint8_t *rs,
                                                 const uint8_t *qs, int stride, int g, int e, int imax, int jmax,
                                                 uint8_t *tile, TileProbe &pb)
{
    const int v = CkView::lane_of(g);
    const int rt = max(BA16_PERIOD * e - v, 0);
    const int c0 = ck.left_col(g);
    const int cL = max(c0, 0);
    const int kfirst = cL - c0;
    const int klast = jmax - c0 - 1;
    const int er_u = S[let], eq_u = S[1];
    const int gl = g - 1;

    const bool hasL = cL > 0;
    // request the boundary in one batch, decode afterwards (see trace16_tile_aff)
    unsigned qraw[MC];
    uint2 rowraw[MC];
#pragma unroll
    for (int k = 0; k < MC; k++) {
        const bool on = k >= kfirst && k <= klast;
        qraw[k] = on ? (unsigned)qs[(c0 + k) * stride] : 0u;
        rowraw[k] = (on && rt >= 1) ? ck.row_raw(g, e - 1, k) : make_uint2(0u, 0u);
    }
    const uint2 cornraw = (hasL && rt >= 1) ? ck.col_raw(gl, rt) : make_uint2(0u, 0u);
    uint2 nraw = (hasL && rt + 1 <= imax) ? ck.col_raw(gl, rt + 1) : make_uint2(0u, 0u);
    int nR = (rt + 1 <= imax) ? (int)rs[(int64_t)rt * stride] : 0;

    int qo[MC], eqv[MC], Tv[MC];
#pragma unroll
    for (int k = 0; k < MC; k++) {
        qo[k] = 0;
        eqv[k] = 0;
        Tv[k] = 0;
        if (k >= kfirst && k <= klast) {
            qo[k] = (int)qraw[k];
            eqv[k] = S[qo[k]];
            if (rt >= 1) {
                int mm, uu;
                ck.row_pick(rowraw[k], mm, uu);
                Tv[k] = mm * BA_SCALE;
            } else if (KIND != KIND_SW) {
                Tv[k] = (c0 + k + 1) * eq_u;    // row 0: nw_letters.go:91-93, fitted_letters.go:83-85
            }
        }
    }
    // T at (row, cL): DP border column (nw_letters.go:94-96; zero for SW and Fitted) or strip gl's
    // checkpointed last column (its raw word w was requested earlier)
    auto left_of = [&](int r, const uint2 &w) -> int {
        if (hasL) {
            if (r == 0) return (KIND == KIND_SW) ? 0 : cL * eq_u;
            int mm, ll;
            ck.col_pick(w, mm, ll);
            return mm * BA_SCALE;
        }
        return (KIND == KIND_NW) ? r * er_u : 0;
    };
    int dprev = left_of(rt, cornraw);
    uint8_t *trow = tile;
    // the next row's letter and left boundary word are requested one row ahead
    for (int r = rt + 1; r <= imax; r++, trow += MC * TS) {
        const int *Srow = S + nR * let;
        const int left = left_of(r, nraw);
        if (r + 1 <= imax) {
            nR = (int)rs[(int64_t)r * stride];
            if (hasL) nraw = ck.col_raw(gl, r + 1);
        }
        const int er = Srow[0];
        int d = dprev, l = left;
#pragma unroll
        for (int k = 0; k < MC; k++) {
            if (k >= kfirst && k <= klast) {
                unsigned code;
                const int old = Tv[k];
                const int nw = linear_cell<KIND>(d, old, l, Srow[qo[k]], er, eqv[k], code);
                trow[k * TS] = (uint8_t)code;
                const int col = c0 + k + 1;
                if (KIND == KIND_SW && pb.want >= 0 && nw == pb.want) {
                    pb.fi = r;
                    pb.fj = col;
                }
                // fitted_letters.go:128-135: last row with T >= max whose letter scores >= 0
                // against the last query letter
                if (KIND == KIND_FIT && col == pb.col && nw >= pb.best && Srow[qo[k]] >= 0) {
                    pb.best = nw;
                    pb.besti = r;
                }
                Tv[k] = nw;
                d = old;
                l = nw;
            }
        }
        dprev = left;
    }
}

// MC: widest strip (columns) the tile arrays are laid out for: BA_MAX_COLS, or BA16_MAX_COLS for waves that
// contain the 10-column single-pass class (ba_plan.h cols16_for)
template <int KIND, bool LIN, int TS = BA_TRACE_THREADS, int MC = BA_MAX_COLS>
__device__ __forceinline__ void trace16_tile(const CkView &ck, const int *S, int let, int go, const uint8_t *rs,
                                             const uint8_t *qs, int stride, int g, int e, int imax, int jmax,
                                             uint8_t *tile, TileProbe &pb)
{
    if (LIN)
        trace16_tile_lin<KIND, TS, MC>(ck, S, let, rs, qs, stride, g, e, imax, jmax, tile, pb);
    else
        trace16_tile_aff<KIND, TS, MC>(ck, S, let, go, rs, qs, stride, g, e, imax, jmax, tile, pb);
}

// One step of the walk: decode the cell's code for the current layer into (move, next layer,
// score delta).  Affine: the seven candidates of align/sw_affine_letters.go:171-275 (diag
// sub-order differs for NW/Fitted, nw_affine_letters.go:255-299).  Linear: diag, up, left
// (align/sw_letters.go:137-171).  Returns false when no candidate matched.
template <int KIND, bool LIN>
__device__ __forceinline__ bool trace16_decode(unsigned code, int layer, int er, int eq, int sub, int go, int &move,
                                               int &nl, int &delta)
{
    nl = 0;
    if (LIN) {
        switch (code & 3u) {
        case 3: move = MV_DIAG; delta = sub; return true;
        case 2: move = MV_UP;   delta = er;  return true;
        case 1: move = MV_LEFT; delta = eq;  return true;
        default: move = MV_DIAG; delta = 0;  return false;
        }
    }
    const unsigned f = (layer == 0) ? (code & 7u) : (layer == 1) ? ((code >> 3) & 3u) : (code >> 5);
    switch (f) {
    case 1: move = MV_UP;   nl = 1; delta = er;      return true;
    case 2: move = MV_LEFT; nl = 2; delta = eq;      return true;
    case 3: move = MV_UP;   nl = 0; delta = go + er; return true;
    case 4: move = MV_LEFT; nl = 0; delta = go + eq; return true;
    case 5: move = MV_DIAG; nl = (KIND == KIND_SW) ? 0 : 1; delta = sub; return true;
    case 6: move = MV_DIAG; nl = (KIND == KIND_SW) ? 1 : 2; delta = sub; return true;
    case 7: move = MV_DIAG; nl = (KIND == KIND_SW) ? 2 : 0; delta = sub; return true;
    default: move = MV_DIAG; delta = 0; return false;
    }
}

// The walk inside one cached tile, shared by the three kernels.  State of a walk:
struct Walk16 {
    int i, j, layer, last, score, maxI, maxJ, cur, nseg;
    int opens;   // NW / Fitted: gap-open steps of the current run (its score is summed when the run ends)
    bool bad;
};

// Score of the run of `last` moves that led from (maxI, maxJ) to (i, j): the sum of the walk's per-step
// deltas (trace16_decode), taken in one tight loop when the run ends instead of step by step inside the
// latency-bound walk.  Only SW needs the running score (it stops when the score is used up).
__device__ __forceinline__ int trace16_run_score(int last, int i, int maxI, int j, int maxJ, int opens, const uint8_t *rs,
                                                 const uint8_t *qs, int stride, const int *S, int let, int go)
{
    int sc = opens * go;
    if (last == MV_DIAG) {
        const uint8_t *pr = rs + (ptrdiff_t)i * stride, *pq = qs + (ptrdiff_t)j * stride;
        for (int x = i; x < maxI; x++, pr += stride, pq += stride) sc += S[(int)*pr * let + (int)*pq] >> BA_SCALE_SHIFT;
    } else if (last == MV_UP) {
        const uint8_t *pr = rs + (ptrdiff_t)i * stride;
        for (int x = i; x < maxI; x++, pr += stride) sc += S[(int)*pr * let] >> BA_SCALE_SHIFT;
    } else {
        const uint8_t *pq = qs + (ptrdiff_t)j * stride;
        for (int y = j; y < maxJ; y++, pq += stride) sc += S[(int)*pq] >> BA_SCALE_SHIFT;
    }
    return sc;
}

// Walks from (w.i, w.j) until the path leaves the tile (rows rt+1.., columns c0+1..), the SW score is
// used up, or no candidate matches.  Same decisions as trace16_decode, arranged for a short instruction
// chain (a team walks one path on one warp per scheduler, so every instruction of a step is exposed
// latency -- measured with -DBA_TRACE_STATS: 230 cycles per row-or-column step before):
//   * the move and the next layer come out of the code byte through constant tables (no branches);
//   * positions are pointers that step by constants; rows and columns left in the tile are counters;
//   * the code byte and the letters of the NEXT cell are requested as soon as the move is known, and
//     only then is the score of the step looked up -- one table read, chosen by the move.
// (Requesting the three possible next code bytes one step ahead and selecting by the move was measured
// and dropped: the extra predicated loads cost more issue slots than the shorter chain saves, 120 -> 150.)
// tcodes: the tile's code bytes, cell (r, k) at ((r * MCW) + k) * TSW.  emitter: this thread writes segments.
// RUNS: sum a run's score when it ends (trace16_run_score) -- for teams, whose threads all walk the same path; a
// thread per pair keeps the per-step score (the run loops of 32 different paths would serialise).  SW always does.
template <int KIND, bool LIN, int MCW, int TSW, bool RUNS>
__device__ __forceinline__ void trace16_walk_tile(Walk16 &w, const uint8_t *tcodes, int rt, int cbase, int c0,
                                                  const uint8_t *rs, const uint8_t *qs, int stride, const int *S, int let,
                                                  int go, int n, int m, int32_t *out, int cap, bool emitter)
{
    // tables indexed by the layer's field x = (code >> shift) & mask of the code byte (trace16_tile_aff packs
    // f0 | f1 << 3 | f2 << 5: three bits for layers M and L, two for U), four bits per entry.
    // MV: the move, 3 = no candidate.  SH / MK: shift and mask of the NEXT layer's field.  OP: the step opens a gap.
    constexpr unsigned SH_D0 = (KIND == KIND_SW) ? 0u : 3u, SH_D1 = (KIND == KIND_SW) ? 3u : 5u,
                       SH_D2 = (KIND == KIND_SW) ? 5u : 0u;   // next layer after the three diagonal candidates
    constexpr unsigned MV_A = 3u | (unsigned)MV_UP << 4 | (unsigned)MV_LEFT << 8 | (unsigned)MV_UP << 12 |
                              (unsigned)MV_LEFT << 16 | (unsigned)MV_DIAG << 20 | (unsigned)MV_DIAG << 24 |
                              (unsigned)MV_DIAG << 28;
    constexpr unsigned SH_A = 0u | 3u << 4 | 5u << 8 | 0u << 12 | 0u << 16 | SH_D0 << 20 | SH_D1 << 24 | SH_D2 << 28;
#define BA_MK_OF(sh_) ((sh_) == 3u ? 3u : 7u)
    constexpr unsigned MK_A = 7u | 3u << 4 | 7u << 8 | 7u << 12 | 7u << 16 | BA_MK_OF(SH_D0) << 20 | BA_MK_OF(SH_D1) << 24 |
                              BA_MK_OF(SH_D2) << 28;
#undef BA_MK_OF
    constexpr unsigned MV_LIN = 3u | (unsigned)MV_LEFT << 4 | (unsigned)MV_UP << 8 | (unsigned)MV_DIAG << 12;
    constexpr unsigned OP_A = 1u << 12 | 1u << 16;   // candidates 3 and 4 open a gap
    constexpr bool STEPS = (KIND == KIND_SW) || !RUNS;   // per-step score
    int ri = w.i - rt, cj = w.j - c0;   // rows / columns of the tile still ahead (including the current cell)
    int cur = w.cur;
    if (ri <= 0 || cj <= 0 || (KIND == KIND_SW && cur == 0)) return;
    const uint8_t *pc = tcodes + ((ri - 1) * MCW + (w.j - cbase - 1)) * TSW;
    const uint8_t *pr = rs + (ptrdiff_t)(w.i - 1) * stride, *pq = qs + (ptrdiff_t)(w.j - 1) * stride;
    unsigned code = *pc;
    int R = 0, Q = 0;
    if (STEPS) {
        R = *pr;
        Q = *pq;
    }
    unsigned sh = LIN ? 0u : (w.layer == 0 ? 0u : w.layer == 1 ? 3u : 5u);
    unsigned mk = LIN ? 3u : (w.layer == 1 ? 3u : 7u);
    int last = w.last, score = w.score, opens = w.opens;
    bool corner = (w.i == n && w.j == m);   // only the first cell of a walk can be the corner
    for (;;) {
        const unsigned t = ((code >> sh) & mk) << 2;
        const int move = (int)(((LIN ? MV_LIN : MV_A) >> t) & 3u);
        if (move == 3) {
            w.bad = true;
            break;
        }
        const bool open = !LIN && ((OP_A >> t) & 1u) != 0u;
        // a new segment when the move changes
        if (last != move) {
            // linear SW has no corner guard (sw_letters.go:147-165); every other walk has
            if (move == MV_DIAG || !corner || (LIN && KIND == KIND_SW)) {
                const int i = rt + ri, j = c0 + cj;
                const int sc = STEPS ? score
                                     : trace16_run_score(last, i, w.maxI, j, w.maxJ, opens, rs, qs, stride, S, let, go);
                if (emitter && w.nseg < cap) {
                    int32_t *o_ = out + 5 * w.nseg;
                    o_[0] = i;
                    o_[1] = w.maxI;
                    o_[2] = j;
                    o_[3] = w.maxJ;
                    o_[4] = sc;
                }
                w.nseg++;
                w.maxI = i;
                w.maxJ = j;
                score = 0;
                opens = 0;
            }
        }
        corner = false;
        if (move != MV_LEFT) {
            ri--;
            pc -= MCW * TSW;
            if (STEPS) pr -= stride;
        }
        if (move != MV_UP) {
            cj--;
            pc -= TSW;
            if (STEPS) pq -= stride;
        }
        const bool more = ri > 0 && cj > 0;
        unsigned ncode = 0;
        int nR = R, nQ = Q;
        if (more) {
            ncode = *pc;
            if (STEPS) {
                nR = *pr;
                nQ = *pq;
            }
        }
        if (STEPS) {
            // the score of the step: one table read, chosen by the move
            const int idx = (move == MV_DIAG) ? R * let + Q : (move == MV_UP) ? R * let : Q;
            const int delta = (S[idx] >> BA_SCALE_SHIFT) + (open ? go : 0);
            score += delta;
            if (KIND == KIND_SW) cur -= delta;
        } else {
            opens += open ? 1 : 0;
        }
        if (!LIN) {
            sh = (SH_A >> t) & 7u;
            mk = (MK_A >> t) & 7u;
        }
        last = move;
        if (!more || (KIND == KIND_SW && cur == 0)) break;
        code = ncode;
        R = nR;
        Q = nQ;
    }
    w.i = rt + ri;
    w.j = c0 + cj;
    w.cur = cur;
    w.last = last;
    w.score = score;
    w.opens = opens;
    w.layer = LIN ? 0 : (sh == 0u ? 0 : sh == 3u ? 1 : 2);
}

// (linear tiles are cheap, the walk is bound by checkpoint-load latency: more resident CTAs)
// TS = threads per CTA: 128, or 64 -- a CTA that fits into the SM slot one ba_fill16 CTA frees
// (registers and shared memory), so a traceback running on a high-priority stream beside the next
// wave's fill is scheduled as fill CTAs retire.
template <int KIND, bool LIN, int TS = BA_TRACE_THREADS, int MC = BA_MAX_COLS>
__global__ void __launch_bounds__(TS, (LIN ? 6 : BA_TRACE_MINB) * (BA_TRACE_THREADS / TS)) ba_trace16_kernel(Trace16Params P)
{
    __shared__ uint8_t tiles[BA_TILE_ROWS * MC * TS];
    __shared__ int S[BA_MAX_LET * BA_MAX_LET];
    {
        const int ll = P.sc->let;
        for (int x = threadIdx.x; x < ll * ll; x += blockDim.x) S[x] = P.sc->la[x] * BA_SCALE;
    }
    __syncthreads();
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= P.n_items) return;
    uint8_t *tile = tiles + threadIdx.x;
    const WorkItem item = P.items[it];
    const int p = item.pair;
    if (P.status[p] != 0) {
        P.seg_count[it] = 0;
        return;
    }
    if ((P.acgt_only[p] & P.flag_mask) != P.flag_mask) {
        P.seg_count[it] = -1;  // the host re-runs this pair on the exact path
        return;
    }
    const int n = P.ref_len[p], m = P.qry_len[p];
    const uint8_t *rs = P.enc + P.ref_off[p];
    const uint8_t *qs = P.enc + P.qry_off[p];
    const int stride = P.stride;
    const DeviceScoring *sc = P.sc;
    const int let = sc->let;
    const int go = LIN ? 0 : sc->gap_open;
    const int go8 = go * BA_SCALE;

    CkView ck;
    ck.init(P.arena, item, n, m, LIN, (KIND == KIND_SW) && !LIN, P.row_bias, P.val_bias);
    ck.npass = (n > 0 && m > 0) ? (int)ba16_npass(item.cols, m) : 0;
    ck.jshift = (KIND == KIND_SW) ? m - ck.npass * 32 * item.cols : 0;

    int32_t *out = P.slots + 5 * P.slot_off[it];
    const int cap = P.slot_cap[it];
    int nseg = 0;
#define T16_EMIT(a0, a1, b0, b1, sc_)     \
    do {                                  \
        if (nseg < cap) {                 \
            int32_t *o_ = out + 5 * nseg; \
            o_[0] = (a0);                 \
            o_[1] = (a1);                 \
            o_[2] = (b0);                 \
            o_[3] = (b1);                 \
            o_[4] = (sc_);                \
        }                                 \
        nseg++;                           \
    } while (0)

    // ---- start cell
    int i = 0, j = 0, cur = 0, layer = 0;
    if (KIND == KIND_SW && n > 0 && m > 0) {
        // the last cell in row-major order holding the table maximum (sw_affine_letters.go:126-134
        // under condition C2 of ba_fill16.cuh)
        int Smax = 0;
#pragma unroll 8
        for (int x = 0; x < ck.npass * 16; x++) {
            const int2 c = ck.cand(x);
            if (c.y >= 0) Smax = max(Smax, c.x);
        }
        if (Smax > 0) {
            for (int x = 0; x < ck.npass * 16; x++) {
                const int2 c = ck.cand(x);
                if (c.y < 0 || c.x != Smax) continue;
                const int e = c.y;
                for (int h = 0; h < 2; h++) {  // the lane's two strips
                    const int g = (x >> 4) * 32 + 2 * (x & 15) + h;
                    const int v = CkView::lane_of(g);
                    const int rlast = min(n, BA16_PERIOD * e + BA16_PERIOD - v);
                    const int clast = ck.left_col(g) + ck.cols;
                    if (rlast < 1 || clast < 1) continue;
                    if (rlast < i || (rlast == i && clast < j)) continue;  // cannot come later
                    TileProbe pb;
                    pb.want = Smax * BA_SCALE;
                    trace16_tile<KIND, LIN, TS, MC>(ck, S, let, go8, rs, qs, stride, g, e, rlast, clast, tile, pb);
                    if (pb.fi > i || (pb.fi == i && pb.fj > j)) {
                        i = pb.fi;
                        j = pb.fj;
                    }
                }
            }
            cur = Smax;
        }
    }
    if (KIND == KIND_NW) {
        // start at (r-1, c-1) in the layer with the largest value, strict '>' scanning M, U, L
        // (nw_affine_letters.go:183-192)
        i = n;
        j = m;
        const int g = ck.strip_of(j);
        TileProbe pb;
        pb.ci = i;
        pb.cj = j;
        trace16_tile<KIND, LIN, TS, MC>(ck, S, let, go8, rs, qs, stride, g, (i + CkView::lane_of(g) - 1) >> BA16_PSHIFT, i, j,
                           tile, pb);
        int best = pb.cM;
        if (pb.cU > best) {
            best = pb.cU;
            layer = 1;
        }
        if (pb.cL > best) layer = 2;
    }
    if (KIND == KIND_FIT) {
        // j = c-1; i = last row maximising M in the last column (fitted_affine_letters.go:170-181)
        j = m;
        const int g = ck.strip_of(j);
        const int v = CkView::lane_of(g);
        TileProbe pb;
        pb.col = m;
        for (int e = 0; BA16_PERIOD * e - v < n; e++) {
            const int rlast = min(n, BA16_PERIOD * e + BA16_PERIOD - v);
            if (rlast < 1) continue;
            trace16_tile<KIND, LIN, TS, MC>(ck, S, let, go8, rs, qs, stride, g, e, rlast, j, tile, pb);
        }
        i = pb.besti;
    }

    Walk16 w;
    w.i = i, w.j = j, w.layer = layer, w.last = MV_DIAG, w.score = 0, w.maxI = i, w.maxJ = j, w.cur = cur, w.nseg = nseg;
    w.opens = 0, w.bad = false;
    while (w.i > 0 && w.j > 0 && !w.bad) {
        if (KIND == KIND_SW && w.cur == 0) break;
        const int g = ck.strip_of(w.j);
        const int v = CkView::lane_of(g);
        const int e = (w.i + v - 1) >> BA16_PSHIFT;
        const int rt = max(BA16_PERIOD * e - v, 0);
        const int c0 = max(ck.left_col(g), 0);
        TileProbe pb0;
        trace16_tile<KIND, LIN, TS, MC>(ck, S, let, go8, rs, qs, stride, g, e, w.i, w.j, tile, pb0);
        trace16_walk_tile<KIND, LIN, MC, TS, false>(w, tile, rt, ck.left_col(g), c0, rs, qs, stride, S, let, go, n, m, out, cap, true);
    }
    i = w.i, j = w.j, nseg = w.nseg;
    const int maxI = w.maxI, maxJ = w.maxJ, score = w.score;
    const bool bad = w.bad;
    T16_EMIT(i, maxI, j, maxJ, score);
    if (KIND == KIND_NW && i != j) {
        // nw_affine_letters.go:311-325: leading gap, its score is the border value
        int b = go;
        if (i == 0)
            for (int x = 0; x < j; x++) b += S[qs[x * stride]] >> BA_SCALE_SHIFT;
        else
            for (int x = 0; x < i; x++) b += S[(int)rs[x * stride] * let] >> BA_SCALE_SHIFT;
        T16_EMIT(0, i, 0, j, b);
    }
#undef T16_EMIT
    if (bad) {
        P.status[p] = 4;  // BA_PAIR_NO_PATH
        nseg = 0;
    }
    P.seg_count[it] = nseg;
}

// ---------------------------------------------------------------------------------------
// Persistent thread-per-pair variant for large batches.  Every thread claims pairs from a global
// counter (the work list is sorted largest first), so a lane that finishes early takes the
// next pair instead of idling until the slowest of its warp is done, and there is no
// under-filled last round of CTAs.  The pair is processed as a small state machine around ONE
// tile computation site -- start-cell search tiles (SW candidates, Fitted last column, NW start
// layer) and walk tiles all go through the same trace16_tile call -- so lanes in different
// phases of different pairs still execute the expensive code together.
enum { T16P_NEW = 0, T16P_SW_SEARCH, T16P_FIT_SEARCH, T16P_NW_PROBE, T16P_WALK, T16P_FINISH };

template <int KIND, bool LIN>
__global__ void __launch_bounds__(BA_TRACE_THREADS, BA_TRACE_MINB) ba_trace16p_kernel(Trace16Params P)
{
    __shared__ uint8_t tiles[BA_TILE_ROWS * BA_MAX_COLS * BA_TRACE_THREADS];
    __shared__ int S[BA_MAX_LET * BA_MAX_LET];
    {
        const int ll = P.sc->let;
        for (int x = threadIdx.x; x < ll * ll; x += blockDim.x) S[x] = P.sc->la[x] * BA_SCALE;
    }
    __syncthreads();
    uint8_t *tile = tiles + threadIdx.x;
    const int stride = P.stride;
    const int let = P.sc->let;
    const int go = LIN ? 0 : P.sc->gap_open;
    const int go8 = go * BA_SCALE;

    // ---- per-pair state
    int phase = T16P_NEW;
    int it = 0, p = 0, n = 0, m = 0, cap = 0, nseg = 0;
    const uint8_t *rs = nullptr, *qs = nullptr;
    int32_t *out = nullptr;
    CkView ck;
    int i = 0, j = 0, cur = 0, layer = 0, maxI = 0, maxJ = 0, score = 0, last = MV_DIAG, opens = 0;
    bool bad = false;
    int sx = 0, sh = 0, Smax = 0;   // SW search: candidate index, strip half, table maximum
    int fe = 0;                     // Fitted search: period
    TileProbe pb;

    for (;;) {
        if (phase == T16P_NEW) {
            it = atomicAdd(P.counter, 1);
            if (it >= P.n_items) break;
            const WorkItem item = P.items[it];
            p = item.pair;
            if (P.status[p] != 0) {
                P.seg_count[it] = 0;
                continue;
            }
            if ((P.acgt_only[p] & P.flag_mask) != P.flag_mask) {
                P.seg_count[it] = -1;  // the host re-runs this pair on the exact path
                continue;
            }
            n = P.ref_len[p];
            m = P.qry_len[p];
            rs = P.enc + P.ref_off[p];
            qs = P.enc + P.qry_off[p];
            ck.init(P.arena, item, n, m, LIN, (KIND == KIND_SW) && !LIN, P.row_bias, P.val_bias);
            ck.npass = (n > 0 && m > 0) ? (int)ba16_npass(item.cols, m) : 0;
            ck.jshift = (KIND == KIND_SW) ? m - ck.npass * 32 * item.cols : 0;
            out = P.slots + 5 * P.slot_off[it];
            cap = P.slot_cap[it];
            nseg = 0;
            i = j = cur = layer = score = opens = 0;
            last = MV_DIAG;
            bad = false;
            pb = TileProbe();
            if (KIND == KIND_SW) {
                Smax = 0;
                if (n > 0 && m > 0) {
                                for (int x = 0; x < ck.npass * 16; x++) {
                        const int2 c = ck.cand(x);
                        if (c.y >= 0) Smax = max(Smax, c.x);
                    }
                }
                sx = 0;
                sh = 0;
                pb.want = Smax * BA_SCALE;
                phase = Smax > 0 ? T16P_SW_SEARCH : T16P_WALK;
            } else if (KIND == KIND_NW) {
                i = n;
                j = m;
                pb.ci = i;
                pb.cj = j;
                phase = T16P_NW_PROBE;
            } else {
                j = m;
                pb.col = m;
                fe = 0;
                phase = T16P_FIT_SEARCH;
            }
            if (phase == T16P_WALK) {
                maxI = i;
                maxJ = j;
            }
        }

        // ---- which tile does this pair need next?
        int tg = 0, te = 0, ti = 0, tj = 0;
        bool need = false;
        if (KIND == KIND_SW && phase == T16P_SW_SEARCH) {
            // candidates: per (pass, lane) {best score, last period attaining it}; a lane has two strips
                while (sx < ck.npass * 16) {
                const int2 c = ck.cand(sx);
                if (c.y >= 0 && c.x == Smax) {
                    const int g = (sx >> 4) * 32 + 2 * (sx & 15) + sh;
                    const int v = CkView::lane_of(g);
                    const int rlast = min(n, BA16_PERIOD * c.y + BA16_PERIOD - v);
                    const int clast = ck.left_col(g) + ck.cols;
                    const bool useful = rlast >= 1 && clast >= 1 && !(rlast < i || (rlast == i && clast < j));
                    if (++sh == 2) {
                        sh = 0;
                        sx++;
                    }
                    if (useful) {
                        tg = g;
                        te = c.y;
                        ti = rlast;
                        tj = clast;
                        need = true;
                        break;
                    }
                } else {
                    sh = 0;
                    sx++;
                }
            }
            if (!need) {
                cur = Smax;
                maxI = i;
                maxJ = j;
                pb.want = -1;
                phase = T16P_WALK;
            }
        }
        if (KIND == KIND_FIT && phase == T16P_FIT_SEARCH) {
            // last row maximising the last column (fitted_affine_letters.go:170-181, fitted_letters.go:112-139)
            const int g = ck.strip_of(j);
            const int v = CkView::lane_of(g);
            while (BA16_PERIOD * fe - v < n) {
                const int rlast = min(n, BA16_PERIOD * fe + BA16_PERIOD - v);
                const int e = fe++;
                if (rlast < 1) continue;
                tg = g;
                te = e;
                ti = rlast;
                tj = j;
                need = true;
                break;
            }
            if (!need) {
                i = pb.besti;
                maxI = i;
                maxJ = j;
                pb.col = -1;
                phase = T16P_WALK;
            }
        }
        if (KIND == KIND_NW && phase == T16P_NW_PROBE) {
            tg = ck.strip_of(j);
            te = (i + CkView::lane_of(tg) - 1) >> BA16_PSHIFT;
            ti = i;
            tj = j;
            need = n > 0 && m > 0;
            if (!need) {
                maxI = i;
                maxJ = j;
                phase = T16P_WALK;
            }
        }
        if (phase == T16P_WALK) {
            if (i > 0 && j > 0 && !bad && !(KIND == KIND_SW && cur == 0)) {
                tg = ck.strip_of(j);
                te = (i + CkView::lane_of(tg) - 1) >> BA16_PSHIFT;
                ti = i;
                tj = j;
                need = true;
            } else {
                phase = T16P_FINISH;
            }
        }
        if (phase == T16P_FINISH) {
#define T16P_EMIT(a0, a1, b0, b1, sc_)    \
    do {                                  \
        if (nseg < cap) {                 \
            int32_t *o_ = out + 5 * nseg; \
            o_[0] = (a0);                 \
            o_[1] = (a1);                 \
            o_[2] = (b0);                 \
            o_[3] = (b1);                 \
            o_[4] = (sc_);                \
        }                                 \
        nseg++;                           \
    } while (0)
            T16P_EMIT(i, maxI, j, maxJ, score);
            if (KIND == KIND_NW && i != j) {
                // nw_affine_letters.go:311-325 / nw_letters.go:183-189: leading gap, its score is the border value
                int b = go;
                if (i == 0)
                    for (int x = 0; x < j; x++) b += S[qs[x * stride]] >> BA_SCALE_SHIFT;
                else
                    for (int x = 0; x < i; x++) b += S[(int)rs[x * stride] * let] >> BA_SCALE_SHIFT;
                T16P_EMIT(0, i, 0, j, b);
            }
            if (bad) {
                P.status[p] = 4;  // BA_PAIR_NO_PATH
                nseg = 0;
            }
            P.seg_count[it] = nseg;
            phase = T16P_NEW;
            continue;
        }
        if (!need) continue;  // the phase changed: decide again

        // ---- the one tile computation all phases share
        trace16_tile<KIND, LIN>(ck, S, let, go8, rs, qs, stride, tg, te, ti, tj, tile, pb);

        // ---- what the tile was for
        if (KIND == KIND_SW && phase == T16P_SW_SEARCH) {
            if (pb.fi > i || (pb.fi == i && pb.fj > j)) {
                i = pb.fi;
                j = pb.fj;
            }
        } else if (KIND == KIND_NW && phase == T16P_NW_PROBE) {
            // start layer: strict '>' scanning M, U, L (nw_affine_letters.go:183-192); linear has one table
            int best = pb.cM;
            if (pb.cU > best) {
                best = pb.cU;
                layer = 1;
            }
            if (pb.cL > best) layer = 2;
            pb.ci = pb.cj = -1;
            maxI = i;
            maxJ = j;
            phase = T16P_WALK;
        } else if (phase == T16P_WALK) {
            const int v = CkView::lane_of(tg);
            const int rt = max(BA16_PERIOD * te - v, 0);
            const int c0 = max(ck.left_col(tg), 0);
            const int cbase = ck.left_col(tg);
            Walk16 w;
            w.i = i, w.j = j, w.layer = layer, w.last = last, w.score = score, w.maxI = maxI, w.maxJ = maxJ, w.cur = cur;
            w.nseg = nseg, w.bad = bad, w.opens = opens;
            trace16_walk_tile<KIND, LIN, BA_MAX_COLS, BA_TRACE_THREADS, false>(w, tile, rt, cbase, c0, rs, qs, stride, S, let, go, n, m,
                                                                        out, cap, true);
            i = w.i, j = w.j, layer = w.layer, last = w.last, score = w.score, maxI = w.maxI, maxJ = w.maxJ, cur = w.cur;
            nseg = w.nseg, bad = w.bad, opens = w.opens;
        }
#undef T16P_EMIT
    }
}

// ---------------------------------------------------------------------------------------
// Team-per-pair variant for batches too small to fill the GPU with one thread per pair
// (kb-scale pairs, BASELINE config 4).  Same tiles, same walk; a team of NT threads
//   * splits the start-cell search (SW candidates / Fitted last-column periods), and
//   * rebuilds a WINDOW of tiles up-left of the walk position concurrently, one full tile per
//     thread; then every thread walks the identical path through the cached tiles (all read
//     the same code byte: a shared-memory broadcast) until it leaves the window.
// NT = 32 (a warp; four pairs per CTA): window = 4 strips x 8 periods.  NT = 128 (the whole CTA,
// for a handful of very long pairs): a diagonal BAND of 42 strips x 3 periods -- strip s of the
// window is centred on the period the main diagonal through the anchor reaches there -- so one
// round covers ~330 columns of a near-diagonal path instead of 32.
// Thread 0 emits.  A window always contains the tile under the walk, so every round advances.
#ifdef BA_TRACE_STATS   // kernel experiments (tools/build_variant.sh): where the team traceback spends its time
__device__ unsigned long long ba_trace_stats[8];   // rounds, steps, cycles: search, rebuild, walk
#define BA_TSTAT_ADD(k, v) do { if (threadIdx.x == 0) atomicAdd(&ba_trace_stats[k], (unsigned long long)(v)); } while (0)
#define BA_TSTAT_CLK() clock64()
#else
#define BA_TSTAT_ADD(k, v) do { } while (0)
#define BA_TSTAT_CLK() 0ll
#endif
#define BA_TW_STRIPS 4
#define BA_TW_PERIODS 8
#define BA_TB_PERIODS 3
#define BA_TB_STRIPS (BA_TRACE_THREADS / BA_TB_PERIODS)
template <int KIND, bool LIN, int NT, int MC = BA_MAX_COLS>
__global__ void __launch_bounds__(BA_TRACE_THREADS, BA_TRACE_MINB) ba_trace16w_kernel(Trace16Params P)
{
    static_assert(NT == 32 || NT == BA_TRACE_THREADS, "a team is a warp or the CTA");
    __shared__ int red[BA_TRACE_THREADS / 32][2];
    __shared__ uint8_t tiles[BA_TILE_ROWS * MC * BA_TRACE_THREADS];
    __shared__ int S[BA_MAX_LET * BA_MAX_LET];
    {
        const int ll = P.sc->let;
        for (int x = threadIdx.x; x < ll * ll; x += blockDim.x) S[x] = P.sc->la[x] * BA_SCALE;
    }
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & (NT - 1);    // my index inside the team
    const int wbase = threadIdx.x & ~(NT - 1);  // first thread of my team: its NT tile columns are the window
    const int it = blockIdx.x * (BA_TRACE_THREADS / NT) + threadIdx.x / NT;
    if (it >= P.n_items) return;                // team-uniform
#define T16W_SYNC()                  \
    do {                             \
        if (NT == 32)                \
            __syncwarp();            \
        else                         \
            __syncthreads();         \
    } while (0)
    // lexicographic maximum of (a, b) over the team, result in every thread
#define T16W_MAX2(a, b)                                                                          \
    do {                                                                                         \
        for (int o_ = 16; o_; o_ >>= 1) {                                                        \
            const int oa_ = __shfl_xor_sync(FULL, (a), o_), ob_ = __shfl_xor_sync(FULL, (b), o_); \
            if (oa_ > (a) || (oa_ == (a) && ob_ > (b))) {                                        \
                (a) = oa_;                                                                       \
                (b) = ob_;                                                                       \
            }                                                                                    \
        }                                                                                        \
        if (NT > 32) {                                                                           \
            __syncthreads();                                                                     \
            if ((threadIdx.x & 31) == 0) {                                                       \
                red[threadIdx.x >> 5][0] = (a);                                                  \
                red[threadIdx.x >> 5][1] = (b);                                                  \
            }                                                                                    \
            __syncthreads();                                                                     \
            for (int w_ = 0; w_ < BA_TRACE_THREADS / 32; w_++) {                                 \
                const int oa_ = red[w_][0], ob_ = red[w_][1];                                    \
                if (oa_ > (a) || (oa_ == (a) && ob_ > (b))) {                                    \
                    (a) = oa_;                                                                   \
                    (b) = ob_;                                                                   \
                }                                                                                \
            }                                                                                    \
        }                                                                                        \
    } while (0)
    uint8_t *tile = tiles + threadIdx.x;
    const WorkItem item = P.items[it];
    const int p = item.pair;
    if (P.status[p] != 0) {
        if (lane == 0) P.seg_count[it] = 0;
        return;
    }
    if ((P.acgt_only[p] & P.flag_mask) != P.flag_mask) {
        if (lane == 0) P.seg_count[it] = -1;
        return;
    }
    const int n = P.ref_len[p], m = P.qry_len[p];
    const uint8_t *rs = P.enc + P.ref_off[p];
    const uint8_t *qs = P.enc + P.qry_off[p];
    const int stride = P.stride;
    const int let = P.sc->let;
    const int go = LIN ? 0 : P.sc->gap_open;
    const int go8 = go * BA_SCALE;

    CkView ck;
    ck.init(P.arena, item, n, m, LIN, (KIND == KIND_SW) && !LIN, P.row_bias, P.val_bias);
    ck.npass = (n > 0 && m > 0) ? (int)ba16_npass(item.cols, m) : 0;
    ck.jshift = (KIND == KIND_SW) ? m - ck.npass * 32 * item.cols : 0;

    int32_t *out = P.slots + 5 * P.slot_off[it];
    const int cap = P.slot_cap[it];
    int nseg = 0;
#define T16W_EMIT(a0, a1, b0, b1, sc_)    \
    do {                                  \
        if (lane == 0 && nseg < cap) {    \
            int32_t *o_ = out + 5 * nseg; \
            o_[0] = (a0);                 \
            o_[1] = (a1);                 \
            o_[2] = (b0);                 \
            o_[3] = (b1);                 \
            o_[4] = (sc_);                \
        }                                 \
        nseg++;                           \
    } while (0)

    // ---- start cell (rules as in ba_trace16_kernel), searched by all lanes
    int i = 0, j = 0, cur = 0, layer = 0;
    long long tclk = BA_TSTAT_CLK();
    (void)tclk;
    if (KIND == KIND_SW && n > 0 && m > 0) {
        int Smax = 0;
        for (int x = lane; x < ck.npass * 16; x += NT) {
            const int2 c = ck.cand(x);
            if (c.y >= 0) Smax = max(Smax, c.x);
        }
        {
            int z = 0;
            T16W_MAX2(Smax, z);
        }
        if (Smax > 0) {
            for (int x = lane; x < ck.npass * 16; x += NT) {
                const int2 c = ck.cand(x);
                if (c.y < 0 || c.x != Smax) continue;
                const int e = c.y;
                for (int h = 0; h < 2; h++) {
                    const int g = (x >> 4) * 32 + 2 * (x & 15) + h;
                    const int v = CkView::lane_of(g);
                    const int rlast = min(n, BA16_PERIOD * e + BA16_PERIOD - v);
                    const int clast = ck.left_col(g) + ck.cols;
                    if (rlast < 1 || clast < 1) continue;
                    if (rlast < i || (rlast == i && clast < j)) continue;
                    TileProbe pb;
                    pb.want = Smax * BA_SCALE;
                    trace16_tile<KIND, LIN, BA_TRACE_THREADS, MC>(ck, S, let, go8, rs, qs, stride, g, e, rlast, clast, tile, pb);
                    if (pb.fi > i || (pb.fi == i && pb.fj > j)) {
                        i = pb.fi;
                        j = pb.fj;
                    }
                }
            }
            // last cell in row-major order over the team
            T16W_MAX2(i, j);
            cur = Smax;
        }
    }
    if (KIND == KIND_NW) {
        i = n;
        j = m;
        const int g = ck.strip_of(j);
        TileProbe pb;
        pb.ci = i;
        pb.cj = j;
        trace16_tile<KIND, LIN, BA_TRACE_THREADS, MC>(ck, S, let, go8, rs, qs, stride, g, (i + CkView::lane_of(g) - 1) >> BA16_PSHIFT, i, j,
                           tile, pb);
        int best = pb.cM;
        if (pb.cU > best) {
            best = pb.cU;
            layer = 1;
        }
        if (pb.cL > best) layer = 2;
    }
    if (KIND == KIND_FIT) {
        j = m;
        const int g = ck.strip_of(j);
        const int v = CkView::lane_of(g);
        TileProbe pb;
        pb.col = m;
        for (int e = lane; BA16_PERIOD * e - v < n; e += NT) {
            const int rlast = min(n, BA16_PERIOD * e + BA16_PERIOD - v);
            if (rlast < 1) continue;
            trace16_tile<KIND, LIN, BA_TRACE_THREADS, MC>(ck, S, let, go8, rs, qs, stride, g, e, rlast, j, tile, pb);
        }
        // last row attaining the maximum: larger value wins, ties go to the larger row
        int best = pb.best, besti = pb.besti;
        T16W_MAX2(best, besti);
        i = besti;
    }
    T16W_SYNC();
    BA_TSTAT_ADD(2, BA_TSTAT_CLK() - tclk);

    int maxI = i, maxJ = j, score = 0, last = MV_DIAG, opens = 0;
    bool bad = false;
    while (i > 0 && j > 0 && !bad) {
        if (KIND == KIND_SW && cur == 0) break;
        tclk = BA_TSTAT_CLK();
        BA_TSTAT_ADD(0, 1);
        BA_TSTAT_ADD(1, i + j);
        // ---- window anchored at (i, j).  Warp team: my tile is strip g_hi - lane / 8, period (period
        // of row i0 there) - lane % 8.  CTA team: strip g_hi - lane / 3, period (period of the
        // diagonal's row i0 - ws*cols there) + 1 - lane % 3.
        const int g_hi = ck.strip_of(j);
        const int i0 = i;
        {
            const int ws = (NT == 32) ? (lane >> 3) : (lane / BA_TB_PERIODS);
            const int g = g_hi - ws;
            const int v = CkView::lane_of(g);  // (negative g is skipped below)
            const int e = (NT == 32) ? ((i0 + v - 1) >> BA16_PSHIFT) - (lane & 7)
                                     : ((max(i0 - ws * ck.cols, 1) + v - 1) >> BA16_PSHIFT) + 1 - (lane % BA_TB_PERIODS);
            const int clast = min(ck.left_col(g) + ck.cols, m);
            const int rlast = min(min(n, BA16_PERIOD * e + BA16_PERIOD - v), i0);
            const int rtop = max(BA16_PERIOD * e - v, 0);
            if (g >= 0 && e >= 0 && clast >= 1 && rlast > rtop) {
                TileProbe pb0;
                trace16_tile<KIND, LIN, BA_TRACE_THREADS, MC>(ck, S, let, go8, rs, qs, stride, g, e, rlast, clast, tile, pb0);
            }
        }
        T16W_SYNC();
        BA_TSTAT_ADD(3, BA_TSTAT_CLK() - tclk);
        tclk = BA_TSTAT_CLK();
        for (;;) {
            if (i <= 0 || j <= 0) break;
            if (KIND == KIND_SW && cur == 0) break;
            const int g = ck.strip_of(j);
            const int v = CkView::lane_of(g);
            const int ws = g_hi - g;
            int we, slot;
            if (NT == 32) {
                we = ((i0 + v - 1) >> BA16_PSHIFT) - ((i + v - 1) >> BA16_PSHIFT);
                if (ws >= BA_TW_STRIPS || we >= BA_TW_PERIODS) break;  // left the window: rebuild
                slot = ws * BA_TW_PERIODS + we;
            } else {
                we = ((max(i0 - ws * ck.cols, 1) + v - 1) >> BA16_PSHIFT) + 1 - ((i + v - 1) >> BA16_PSHIFT);
                if (ws >= BA_TB_STRIPS || we < 0 || we >= BA_TB_PERIODS) break;
                slot = ws * BA_TB_PERIODS + we;
            }
            // inside this tile the walk only needs tile-local arithmetic
            const int rt = max(BA16_PERIOD * ((i + v - 1) >> BA16_PSHIFT) - v, 0);
            const int cbase = ck.left_col(g);
            const int c0 = max(cbase, 0);
            const uint8_t *tcodes = tiles + wbase + slot;
            Walk16 w;
            w.i = i, w.j = j, w.layer = layer, w.last = last, w.score = score, w.maxI = maxI, w.maxJ = maxJ, w.cur = cur;
            w.nseg = nseg, w.bad = bad, w.opens = opens;
            trace16_walk_tile<KIND, LIN, MC, BA_TRACE_THREADS, true>(w, tcodes, rt, cbase, c0, rs, qs, stride, S, let, go, n, m, out, cap,
                                                               lane == 0);
            i = w.i, j = w.j, layer = w.layer, last = w.last, score = w.score, maxI = w.maxI, maxJ = w.maxJ, cur = w.cur;
            nseg = w.nseg, bad = w.bad, opens = w.opens;
            if (bad) break;
        }
        BA_TSTAT_ADD(4, BA_TSTAT_CLK() - tclk);
        BA_TSTAT_ADD(1, -(long long)(i + j));
        T16W_SYNC();
    }
    if (KIND != KIND_SW) score = trace16_run_score(last, i, maxI, j, maxJ, opens, rs, qs, stride, S, let, go);
    T16W_EMIT(i, maxI, j, maxJ, score);
    if (KIND == KIND_NW && i != j) {
        int b = go;
        if (i == 0)
            for (int x = 0; x < j; x++) b += S[qs[x * stride]] >> BA_SCALE_SHIFT;
        else
            for (int x = 0; x < i; x++) b += S[(int)rs[x * stride] * let] >> BA_SCALE_SHIFT;
        T16W_EMIT(0, i, 0, j, b);
    }
#undef T16W_EMIT
#undef T16W_SYNC
#undef T16W_MAX2
    if (lane == 0) {
        if (bad) {
            P.status[p] = 4;  // BA_PAIR_NO_PATH
            nseg = 0;
        }
        P.seg_count[it] = nseg;
    }
}

// ---- per-wave finish: clip + overflow list, scan (cub), compact + pair-order scatter, totals ----
// Device-side bookkeeping of one batch (an int64 array): [0] = segments compacted so far (the next wave's
// base), [1] = pairs handed to the exact path, then three words per wave {base, segments, overflow so far}
// that travel to the host after every wave.
#define BA_META_HEAD 2   // int64 words before the per-wave records
#define BA_META_WAVE 3   // int64 words per wave

// seg_count clipped to what actually gets compacted (overflowed / ineligible items contribute 0
// and are appended to the overflow list: the host re-runs them on the exact two-pass path)
__global__ void ba_clip_counts_kernel(const int32_t *__restrict__ seg_count, const int32_t *__restrict__ slot_cap,
                                      const WorkItem *__restrict__ items, int32_t *__restrict__ clipped,
                                      long long *__restrict__ meta, int32_t *__restrict__ overflow, int n_items)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    const int c = seg_count[it];
    const bool ovf = (c < 0 || c > slot_cap[it]);
    clipped[it] = ovf ? 0 : c;
    if (ovf) {
        const long long k = (long long)atomicAdd((unsigned long long *)(meta + 1), 1ull);
        overflow[k] = items[it].pair;
    }
}

// Reverse the emission-order slots of every item into its final, compact place (after the segments
// of the earlier waves) and publish the pair's (seg_start, seg_count) in PAIR order, so the host
// neither scans nor scatters.
__global__ void ba_compact_kernel(const int32_t *__restrict__ slots, const int64_t *__restrict__ slot_off,
                                  const int32_t *__restrict__ slot_cap, const int32_t *__restrict__ seg_count,
                                  const int64_t *__restrict__ seg_base, const WorkItem *__restrict__ items,
                                  const long long *__restrict__ meta, int32_t *__restrict__ segs,
                                  int64_t *__restrict__ pair_seg_start, int32_t *__restrict__ pair_seg_count,
                                  int n_items)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    const int cnt = seg_count[it];
    if (cnt < 0 || cnt > slot_cap[it]) return;  // not eligible / overflow: re-run on the exact two-pass path
    const int64_t base = (int64_t)meta[0] + seg_base[it];
    const int32_t p = items[it].pair;
    pair_seg_start[p] = base;
    pair_seg_count[p] = cnt;
    const int32_t *src = slots + 5 * slot_off[it];
    int32_t *dst = segs + 5 * base;
    for (int k = 0; k < cnt; k++) {
        const int32_t *s = src + 5 * (cnt - 1 - k);
#pragma unroll
        for (int x = 0; x < 5; x++) dst[5 * k + x] = s[x];
    }
}

// after the compaction of wave `w`: record {base, segments, overflow so far} and advance the total
__global__ void ba_wave_total_kernel(const int64_t *__restrict__ seg_base, const int32_t *__restrict__ clipped,
                                     int n_items, long long *__restrict__ meta, int w)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const long long wsegs = n_items > 0 ? (long long)seg_base[n_items - 1] + clipped[n_items - 1] : 0;
    long long *rec = meta + BA_META_HEAD + (long long)BA_META_WAVE * w;
    rec[0] = meta[0];
    rec[1] = wsegs;
    rec[2] = meta[1];
    meta[0] += wsegs;
}
