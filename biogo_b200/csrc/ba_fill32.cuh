// ba_fill32.cuh -- 32-bit "exact code" DP fill kernels (all six aligner kinds).
//
// One warp per pair, systolic over the query: lane t owns COLS consecutive columns of a
// 32*COLS wide pass and processes row i at step s = i-1+t, so every step the warp
// computes one anti-diagonal band of cells.  Boundary values move between neighbouring
// lanes with warp shuffles; between passes (queries longer than 32*COLS) the right-most
// column is staged through a small global scratch row.  Every cell's traceback code is
// derived from VALUE EQUALITY in the reference's switch order (never from "which operand
// won the max"), which is what makes the result bit-exact with bíogo's layer-independent
// traceback (align/sw_affine_letters.go:171-275 etc.).
//
// Recurrences restated from: align/sw_letters.go:92-120, align/nw_letters.go:88-118,
// align/fitted_letters.go:80-107, align/sw_affine_letters.go:108-157,
// align/nw_affine_letters.go:111-174, align/fitted_affine_letters.go:103-164.
#pragma once

#include "ba_common.cuh"

struct FillParams {
    const uint8_t *enc;        // letters mapped through alphabet.Index (0xFF illegal)
    int stride;                // 1 Letters, 2 QLetters
    const int64_t *ref_off;
    const int32_t *ref_len;
    const int64_t *qry_off;
    const int32_t *qry_len;
    const int32_t *status;     // per pair, BA_PAIR_*
    const WorkItem *items;     // this launch's work list
    int n_items;
    int *counter;              // dynamic work counter (one per launch)
    uint8_t *arena;            // traceback codes
    int32_t *bnd;              // inter-pass boundary scratch, bnd_stride ints per warp
    int64_t bnd_stride;
    PairStart *start;          // per pair
    int32_t *aux;              // Fitted: last-column values, per item aux_off
    const int64_t *aux_off;    // per item (parallel to items)
    const DeviceScoring *sc;
};

struct Cell3 {
    int M, U, L;
};

__device__ __forceinline__ unsigned umin3(unsigned a, unsigned b, unsigned c)
{
    return __vimin3_u32(a, b, c);
}

// One affine cell.  Returns the new (M,U,L) and the packed 8-bit code
//   bits 0-2: first candidate equal to M  (1..7, numbering below)
//   bits 3-4: first candidate equal to U  (1..3)
//   bits 5-7: first candidate equal to L  (1..4)
// Candidate order (align/sw_affine_letters.go:171-275 / nw_affine_letters.go:198-299):
//   1: U[p-c]+er  2: L[p-1]+eq  3: M[p-c]+go+er  4: M[p-1]+go+eq
//   SW: 5: M[p-c-1]+sub 6: U[p-c-1]+sub 7: L[p-c-1]+sub
//   NW/Fitted: 5: U[p-c-1]+sub 6: L[p-c-1]+sub 7: M[p-c-1]+sub
template <int KIND>
__device__ __forceinline__ Cell3 affine_cell(const Cell3 &d, const Cell3 &u, const Cell3 &l,
                                             int sub, int er, int goer, int eq, int goeq,
                                             unsigned &code, int &dm_out)
{
    const int c1 = u.U + er;
    const int c3 = u.M + goer;
    const int c2 = l.L + eq;
    const int c4 = l.M + goeq;
    const int dm = __vimax3_s32(d.M, d.U, d.L);
    Cell3 o;
    if (KIND == KIND_SW) {
        o.M = __viaddmax_s32_relu(dm, sub, 0);
        o.U = __vimax3_s32_relu(c1, c3, 0);
        o.L = __vimax3_s32_relu(c2, c4, 0);
    } else {
        o.M = dm + sub;
        o.U = max(c1, c3);
        o.L = max(c2, c4);
    }
    const int tM = (KIND == KIND_SW) ? 5 : 7;
    const int tU = (KIND == KIND_SW) ? 6 : 5;
    const int tL = (KIND == KIND_SW) ? 7 : 6;
    // M layer: seven candidates.  The diagonal three are compared against dm (M = dm+sub
    // whenever M is not clamped; a clamped M is 0 and its code is never consulted).
    unsigned m1 = umin3((unsigned)(c1 + 1 - o.M), (unsigned)(c2 + 2 - o.M), (unsigned)(c3 + 3 - o.M));
    unsigned m2 = umin3((unsigned)(c4 + 4 - o.M), (unsigned)(d.M - dm + tM), (unsigned)(d.U - dm + tU));
    unsigned cm = umin3(m1, m2, (unsigned)(d.L - dm + tL));
    // U layer: candidates 1..3; L layer: candidates 1..4.  For SW a clamped layer has no
    // matching candidate; bound the key so it cannot spill into the neighbouring fields.
    unsigned u3 = (unsigned)(c3 + 3 - o.U);
    unsigned l4 = (unsigned)(c4 + 4 - o.L);
    if (KIND == KIND_SW) {
        u3 = min(u3, 3u);
        l4 = min(l4, 4u);
    }
    unsigned cu = umin3((unsigned)(c1 + 1 - o.U), (unsigned)(c2 + 2 - o.U), u3);
    unsigned cl = umin3((unsigned)(c1 + 1 - o.L), (unsigned)(c2 + 2 - o.L), (unsigned)(c3 + 3 - o.L));
    cl = min(cl, l4);
    code = cm | (cu << 3) | (cl << 5);
    dm_out = dm;
    return o;
}

// One linear cell (scaled values, low three bits clear).  The three candidates carry the
// tags 3 (diag), 2 (up), 1 (left) so a single signed 3-way max returns both the value
// and the FIRST equal candidate in the reference's order diag, up, left
// (align/sw_letters.go:137-171, align/nw_letters.go:133-171).
template <int KIND>
__device__ __forceinline__ int linear_cell(int dT, int uT, int lT, int sub, int er, int eq,
                                           unsigned &code)
{
    const int cd = dT + sub + 3;
    const int cu = uT + er + 2;
    const int cl = lT + eq + 1;
    int t = (KIND == KIND_SW) ? __vimax3_s32_relu(cd, cu, cl) : __vimax3_s32(cd, cu, cl);
    code = (unsigned)t & 3u;
    return t & ~7;
}

template <int COLS, typename T>
__device__ __forceinline__ T select_col(const T (&a)[COLS], int k)
{
    T v = a[0];
#pragma unroll
    for (int x = 1; x < COLS; x++)
        if (k == x) v = a[x];
    return v;
}

template <int KIND, bool AFFINE, int COLS>
__global__ void __launch_bounds__(BA_FILL_WARPS * 32) ba_fill32_kernel(FillParams P)
{
    __shared__ int S[BA_MAX_LET * BA_MAX_LET];
    const int let = P.sc->let;
    for (int x = threadIdx.x; x < let * let; x += blockDim.x) S[x] = P.sc->la[x] * BA_SCALE;
    __syncthreads();
    const int go = P.sc->gap_open * BA_SCALE;

    const int t = threadIdx.x & 31;
    const int warp_global = blockIdx.x * BA_FILL_WARPS + (threadIdx.x >> 5);
    int32_t *bnd = P.bnd + (int64_t)warp_global * P.bnd_stride;
    constexpr int W = 32 * COLS;
    constexpr int CB = AFFINE ? (COLS <= 4 ? 4 : 8) : (COLS <= 4 ? 1 : 2);
    const unsigned FULL = 0xffffffffu;

    for (;;) {
        int it = 0;
        if (t == 0) it = atomicAdd(P.counter, 1);
        it = __shfl_sync(FULL, it, 0);
        if (it >= P.n_items) break;
        const WorkItem item = P.items[it];
        const int p = item.pair;
        const int n = P.ref_len[p], m = P.qry_len[p];
        if (P.status[p] != 0 || n <= 0 || m <= 0) continue;
        const uint8_t *rptr = P.enc + P.ref_off[p];
        const uint8_t *qptr = P.enc + P.qry_off[p];
        uint8_t *codes = P.arena + item.code_off;
        const int stride = P.stride;
        const int npass = (m + W - 1) / W;
        const int nsteps = n + 31;
        const int tm = ((m - 1) % W) / COLS, km = (m - 1) % COLS;  // owner of the last column

        int carry = 0;                       // sum of eq over columns 1..jb (scaled)
        int best = BA_SCALE, bi = 0, bj = 0; // SW best cell (score > 0 required)

        for (int pass = 0; pass < npass; pass++) {
            const bool lastpass = (pass == npass - 1);
            const int j0 = pass * W + t * COLS;  // my columns are j0+1 .. j0+COLS
            int qoff[COLS], eqv[COLS];
#pragma unroll
            for (int k = 0; k < COLS; k++) {
                int j = j0 + 1 + k;
                int q = (j <= m) ? (int)qptr[(int64_t)(j - 1) * stride] : 0;
                qoff[k] = q;
                eqv[k] = S[q];  // la[gap*let + q]
            }
            // ---- row-0 border for my columns, and the diagonal seed (row 0, column j0)
            Cell3 Pv[COLS];   // previous-row values of my columns (affine)
            int Tv[COLS];     // previous-row values (linear)
            Cell3 sdiag;      // (row i-1, column j0): diagonal input of my first column
            int sdiagT = 0;
            sdiag.M = sdiag.U = sdiag.L = 0;
            if (KIND != KIND_SW) {
                int pre[COLS];
                int acc = 0;
#pragma unroll
                for (int k = 0; k < COLS; k++) {
                    acc += eqv[k];
                    pre[k] = acc;
                }
                int incl = acc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(FULL, incl, o);
                    if (t >= o) incl += v;
                }
                const int base = carry + incl - acc;  // border sum up to column j0
                const int total = __shfl_sync(FULL, incl, 31);
#pragma unroll
                for (int k = 0; k < COLS; k++) {
                    // nw_affine_letters.go:119-131 / nw_letters.go:91-93
                    Pv[k].M = BA_NEG;
                    Pv[k].U = BA_NEG;
                    Pv[k].L = go + base + pre[k];
                    Tv[k] = base + pre[k];
                }
                if (j0 == 0) {
                    sdiag.M = 0;  // table[0] = {0, minInt, minInt}
                    sdiag.U = BA_NEG;
                    sdiag.L = BA_NEG;
                    sdiagT = 0;
                } else {
                    sdiag.M = BA_NEG;
                    sdiag.U = BA_NEG;
                    sdiag.L = go + base;
                    sdiagT = base;
                }
                carry += total;
            } else {
#pragma unroll
                for (int k = 0; k < COLS; k++) {
                    Pv[k].M = Pv[k].U = Pv[k].L = 0;
                    Tv[k] = 0;
                }
            }
#pragma unroll
            for (int k = 0; k < COLS; k++) qoff[k] *= 4;  // byte offset into a matrix row
            int goeq[COLS];
#pragma unroll
            for (int k = 0; k < COLS; k++) goeq[k] = go + eqv[k];

            Cell3 outl;       // my last column's newest values, handed to lane t+1
            outl.M = outl.U = outl.L = 0;
            int outT = 0;
            int cum_er = 0;   // lane 0, pass 0: running column-0 border (NW kinds)

            uint8_t *cptr = codes + ((int64_t)pass * nsteps * 32 + t) * CB;

            for (int s = 0; s < nsteps; s++, cptr += 32 * CB) {
                const int i = s - t + 1;
                Cell3 recv;
                int recvT = 0;
                if (AFFINE) {
                    recv.M = __shfl_up_sync(FULL, outl.M, 1);
                    recv.U = __shfl_up_sync(FULL, outl.U, 1);
                    recv.L = __shfl_up_sync(FULL, outl.L, 1);
                } else {
                    recvT = __shfl_up_sync(FULL, outT, 1);
                    recv.M = recv.U = recv.L = 0;
                }
                if (i < 1 || i > n) continue;  // ramp-up / ramp-down of the systolic array

                const int R = rptr[(int64_t)(i - 1) * stride];
                const int *Srow = S + R * let;
                const int er = Srow[0];  // la[rVal*let + gap]
                const int goer = go + er;
                if (t == 0) {
                    // left boundary: DP column 0 (pass 0) or the staged column of the previous pass
                    if (pass == 0) {
                        if (KIND == KIND_NW) cum_er += er;
                        if (AFFINE) {
                            if (KIND == KIND_SW) {
                                recv.M = recv.U = recv.L = 0;
                            } else if (KIND == KIND_NW) {
                                recv.M = BA_NEG;  // nw_affine_letters.go:132-142
                                recv.U = go + cum_er;
                                recv.L = BA_NEG;
                            } else {
                                recv.M = BA_NEG;  // fitted_affine_letters.go:123-132 (up omitted = 0)
                                recv.U = 0;
                                recv.L = BA_NEG;
                            }
                        } else {
                            recvT = (KIND == KIND_NW) ? cum_er : 0;
                        }
                    } else {
                        if (AFFINE) {
                            recv.M = bnd[3 * i + 0];
                            recv.U = bnd[3 * i + 1];
                            recv.L = bnd[3 * i + 2];
                        } else {
                            recvT = bnd[i];
                        }
                    }
                }

                unsigned cw0 = 0, cw1 = 0;  // packed codes of this lane-step
                if (AFFINE) {
                    Cell3 d = sdiag, l = recv;
                    int rowmax = 0;
                    int dmv[COLS];
#pragma unroll
                    for (int k = 0; k < COLS; k++) {
                        const int sub = *(const int *)((const char *)Srow + qoff[k]);
                        unsigned code;
                        const Cell3 old = Pv[k];
                        const Cell3 nw = affine_cell<KIND>(d, old, l, sub, er, goer, eqv[k], goeq[k],
                                                           code, dmv[k]);
                        if (KIND == KIND_SW) {
                            dmv[k] = (dmv[k] == d.M);  // matched := score == diagScore
                            rowmax = max(rowmax, nw.M);
                        }
                        Pv[k] = nw;
                        d = old;
                        l = nw;
                        if (k < 4)
                            cw0 |= code << (8 * k);
                        else
                            cw1 |= code << (8 * (k - 4));
                    }
                    if (KIND == KIND_SW && rowmax >= best) {
                        // sw_affine_letters.go:126-134, row-major order inside my strip
#pragma unroll
                        for (int k = 0; k < COLS; k++) {
                            // "last in row-major order": rows only grow inside a pass, but a later pass
                            // revisits small rows, so ties are settled on (i, j), not on arrival order
                            const int v = Pv[k].M, jj = j0 + 1 + k;
                            if (dmv[k] && jj <= m && (v > best || (v == best && (i > bi || (i == bi && jj > bj))))) {
                                best = v;
                                bi = i;
                                bj = jj;
                            }
                        }
                    }
                    sdiag = recv;
                    outl = Pv[COLS - 1];
                    if (CB == 4)
                        *(unsigned *)cptr = cw0;
                    else
                        *(uint2 *)cptr = make_uint2(cw0, cw1);
                    if (!lastpass && t == 31) {
                        bnd[3 * i + 0] = outl.M;
                        bnd[3 * i + 1] = outl.U;
                        bnd[3 * i + 2] = outl.L;
                    }
                    if (lastpass && t == tm) {
                        if (KIND == KIND_FIT) {
                            int mm[COLS];
#pragma unroll
                            for (int k = 0; k < COLS; k++) mm[k] = Pv[k].M;
                            P.aux[P.aux_off[it] + i] = select_col<COLS>(mm, km);
                        }
                        if (KIND == KIND_NW && i == n) {
                            int a[COLS], b[COLS], c[COLS];
#pragma unroll
                            for (int k = 0; k < COLS; k++) {
                                a[k] = Pv[k].M;
                                b[k] = Pv[k].U;
                                c[k] = Pv[k].L;
                            }
                            // nw_affine_letters.go:185-191: strict '>' scanning M, U, L
                            int bm = select_col<COLS>(a, km), layer = 0;
                            int bu = select_col<COLS>(b, km), bl = select_col<COLS>(c, km);
                            if (bu > bm) {
                                bm = bu;
                                layer = 1;
                            }
                            if (bl > bm) layer = 2;
                            PairStart st;
                            st.score = 0;
                            st.i = n;
                            st.j = m;
                            st.layer = layer;
                            P.start[p] = st;
                        }
                    }
                } else {
                    int d = sdiagT, l = recvT;
                    int rowmax = 0;
                    unsigned cbits = 0;
                    unsigned cdv[COLS];
#pragma unroll
                    for (int k = 0; k < COLS; k++) {
                        const int sub = *(const int *)((const char *)Srow + qoff[k]);
                        const int old = Tv[k];
                        const int nw = linear_cell<KIND>(d, old, l, sub, er, eqv[k], cdv[k]);
                        if (KIND == KIND_SW) rowmax = max(rowmax, nw);
                        Tv[k] = nw;
                        d = old;
                        l = nw;
                        cbits |= cdv[k] << (2 * k);
                    }
                    if (KIND == KIND_SW && rowmax >= best) {
                        // sw_letters.go:110-114: score >= maxS && score == diagScore
#pragma unroll
                        for (int k = 0; k < COLS; k++) {
                            const int v = Tv[k], jj = j0 + 1 + k;  // ties: see the affine branch
                            if (cdv[k] == 3u && jj <= m && (v > best || (v == best && (i > bi || (i == bi && jj > bj))))) {
                                best = v;
                                bi = i;
                                bj = jj;
                            }
                        }
                    }
                    sdiagT = recvT;
                    outT = Tv[COLS - 1];
                    if (CB == 1)
                        *cptr = (uint8_t)cbits;
                    else
                        *(uint16_t *)cptr = (uint16_t)cbits;
                    if (!lastpass && t == 31) bnd[i] = outT;
                    if (KIND == KIND_FIT && lastpass && t == tm)
                        P.aux[P.aux_off[it] + i] = select_col<COLS>(Tv, km);
                }
            }
            __syncwarp();
        }

        if (KIND == KIND_SW) {
            // lexicographic max of (score, i, j) == "last in row-major order among the
            // qualifying cells with the largest score" (sw_letters.go:112)
            if (bi == 0) best = 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                int ob = __shfl_xor_sync(FULL, best, o);
                int oi = __shfl_xor_sync(FULL, bi, o);
                int oj = __shfl_xor_sync(FULL, bj, o);
                bool take = (ob > best) || (ob == best && (oi > bi || (oi == bi && oj > bj)));
                if (take) {
                    best = ob;
                    bi = oi;
                    bj = oj;
                }
            }
            if (t == 0) {
                PairStart st;
                st.score = best >> BA_SCALE_SHIFT;
                st.i = bi;
                st.j = bj;
                st.layer = 0;
                P.start[p] = st;
            }
        }
    }
}
