// ba_walk.cuh -- letter encoding, input validation and the on-device traceback walk.
//
// The walk follows the packed direction codes written by the fill kernels and emits
// feat.Pair segments with the reference's exact emission rules:
//   align/sw_letters.go:125-190, align/nw_letters.go:123-193, align/fitted_letters.go:112-199,
//   align/sw_affine_letters.go:162-289, align/nw_affine_letters.go:179-329,
//   align/fitted_affine_letters.go:169-305.
// Segment scores telescope (every step adds cur - pred, which equals the matched
// candidate's constant), so the walk needs only codes, letters and the matrix.
#pragma once

#include "ba_common.cuh"

// ---------------------------------------------------------------------------
// enc[x] = alphabet.Index[letters[x]] for the whole packed buffer (0xFF = not in the
// alphabet, 0xFE = index beyond the scoring matrix).  Pure function of the position, so
// sequences may share or overlap storage.
__global__ void ba_encode_kernel(const uint8_t *__restrict__ letters, uint8_t *__restrict__ enc,
                                 int64_t nbytes, const DeviceScoring *sc)
{
    __shared__ uint8_t map[256];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) {
        uint8_t v = sc->index[x];
        if (v != 0xFF && v >= sc->let) v = 0xFE;
        map[x] = v;
    }
    __syncthreads();
    const int64_t nvec = nbytes / 16;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((((uintptr_t)letters) | ((uintptr_t)enc)) & 15) == 0;
    if (aligned) {
        for (int64_t v = gid; v < nvec; v += gsz) {
            uint4 in = ((const uint4 *)letters)[v];
            uint32_t w[4] = {in.x, in.y, in.z, in.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t o = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) o |= (uint32_t)map[(w[q] >> (8 * b)) & 0xFF] << (8 * b);
                w[q] = o;
            }
            ((uint4 *)enc)[v] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        for (int64_t x = nvec * 16 + gid; x < nbytes; x += gsz) enc[x] = map[letters[x]];
    } else {
        for (int64_t x = gid; x < nbytes; x += gsz) enc[x] = map[letters[x]];
    }
}

// ---------------------------------------------------------------------------
// Per-pair validation: reproduces which error (or panic) the reference raises first.
//  SW/SWAffine   sw_letters.go:97-102: checked per cell, ref first => ref[0], then the first
//                bad query letter, then the first bad ref letter; nothing if a side is empty.
//  NW            nw_letters.go:91-96: border set-up indexes la[index[..]] => Go panic.
//  Fitted        fitted_letters.go:83-85 panic on a bad query letter; :120-126 panic on an
//                empty query; bad ref letter => Errorf (:94-96).
//  NWAffine      nw_affine_letters.go:119-142: panics on empty input or any bad letter.
//  FittedAffine  fitted_affine_letters.go:108-132: panics on empty input / bad query letter;
//                bad ref letter => Errorf (:141-143).
__global__ void ba_validate_kernel(const uint8_t *__restrict__ enc, int stride,
                                   const int64_t *__restrict__ ref_off, const int32_t *__restrict__ ref_len,
                                   const int64_t *__restrict__ qry_off, const int32_t *__restrict__ qry_len,
                                   int64_t npairs, int kind, int affine,
                                   int32_t *__restrict__ status, int64_t *__restrict__ err_pos,
                                   uint8_t *__restrict__ acgt_only)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < npairs; p += nwarps) {
        const int n = ref_len[p], m = qry_len[p];
        const uint8_t *r = enc + ref_off[p];
        const uint8_t *q = enc + qry_off[p];
        // Per lane: first illegal position and the letter-class flags of its share; one reduction
        // per pair at the end (no warp collective inside the scan, so the byte loads pipeline).
        int bad_r = 0x7fffffff, bad_q = 0x7fffffff;  // first illegal position
        unsigned flags = 0;                          // 1: index beyond the matrix, 2: outside 1..4, 4: gap letter
        // A sequence is scanned in aligned 32-bit words (the encoded buffer has slack behind its last byte), four
        // (Letters) or two (QLetters) letters per load, and every test is made on the whole word: bytes outside
        // [0, len*stride) and quality bytes are first replaced by the harmless index 1, then "some byte is 0 / 0xFF /
        // 0xFE / above 4" are the classic carry tricks (exact as ANY-byte tests; a per-byte loop cost ten
        // instructions per letter and made this kernel instruction-bound).  Only a word that holds an illegal
        // letter (0xFF) is looked at byte by byte, for the position.
        auto scan = [&](const uint8_t *sq, int len, int &bad) {
            if (stride <= 2) {
                const int head = (int)((uintptr_t)sq & 3u);
                const uint32_t *wp = (const uint32_t *)(sq - head);
                const int total = len * stride, nwords = (head + total + 3) >> 2;
                // letter positions inside a word: all four bytes, or every second one starting at the head's parity
                const uint32_t lmask = stride == 1 ? 0xFFFFFFFFu : ((head & 1) ? 0xFF00FF00u : 0x00FF00FFu);
#pragma unroll 4
                for (int w = lane; w < nwords; w += 32) {
                    const uint32_t word = wp[w];
                    const int o0 = 4 * w - head;            // sequence offset of the word's byte 0
                    uint32_t vm = lmask;
                    if (o0 < 0) vm &= 0xFFFFFFFFu << (8 * (-o0));                       // bytes before the sequence
                    if (o0 + 4 > total) vm &= 0xFFFFFFFFu >> (8 * (o0 + 4 - total));    // bytes behind it
                    const uint32_t x = (word & vm) | (0x01010101u & ~vm);
                    const uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;            // some byte == 0
                    const uint32_t nx = ~x, ff = (nx - 0x01010101u) & ~nx & 0x80808080u;   // some byte == 0xFF
                    const uint32_t y = x ^ 0xFEFEFEFEu, fe = (y - 0x01010101u) & ~y & 0x80808080u;   // some byte == 0xFE
                    const uint32_t gt4 = ((x & 0x7F7F7F7Fu) + 0x7B7B7B7Bu | x) & 0x80808080u;   // some byte > 4
                    flags |= (fe ? 1u : 0u) | ((z | gt4) ? 2u : 0u) | (z ? 4u : 0u);
                    if (ff) {
#pragma unroll
                        for (int b = 3; b >= 0; b--)
                            if (((x >> (8 * b)) & 0xFFu) == 0xFFu) bad = min(bad, (o0 + b) / stride);
                    }
                }
            } else {
#pragma unroll 4
                for (int x = lane; x < len; x += 32) {
                    const unsigned v = sq[(int64_t)x * stride];
                    if (v == 0xFFu) bad = min(bad, x);
                    flags |= (v == 0xFEu ? 1u : 0u) | ((v < 1u || v > 4u) ? 2u : 0u) | (v == 0u ? 4u : 0u);
                }
            }
        };
        scan(r, n, bad_r);
        scan(q, m, bad_q);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bad_r = min(bad_r, __shfl_xor_sync(0xffffffffu, bad_r, o));
            bad_q = min(bad_q, __shfl_xor_sync(0xffffffffu, bad_q, o));
            flags |= __shfl_xor_sync(0xffffffffu, flags, o);
        }
        const int oob = (int)(flags & 1u), other = (int)(flags & 2u), zero = (int)(flags & 4u);
        const bool hr = bad_r != 0x7fffffff, hq = bad_q != 0x7fffffff;
        int st = 0;
        int64_t ep = -1;
        if (kind == KIND_SW) {
            if (n > 0 && m > 0) {
                if (hr && bad_r == 0) {
                    st = 1; ep = 0;
                } else if (hq) {
                    st = 2; ep = bad_q;
                } else if (hr) {
                    st = 1; ep = bad_r;
                } else if (oob) {
                    st = 3;
                }
            }
        } else if (kind == KIND_NW) {
            if (affine && (n == 0 || m == 0)) st = 3;
            else if (hr || hq || oob) st = 3;
        } else {
            if (m == 0 || hq) st = 3;
            else if (affine && n == 0) st = 3;
            else if (hr) {
                st = 1; ep = bad_r;
            } else if (oob) st = 3;
        }
        if (lane == 0) {
            status[p] = st;
            err_pos[p] = ep;
            // the scan above stops at the first illegal letter, which makes st != 0 anyway
            // bit0: only alphabet indices 1..4 (PRMT profiles); bit1: no gap letter (uniform gap entries)
            if (acgt_only) acgt_only[p] = (st == 0) ? ((other ? 0 : 1) | (zero ? 0 : 2)) : 0;
        }
    }
}

// ---------------------------------------------------------------------------
struct WalkParams {
    const uint8_t *enc;
    int stride;
    const int64_t *ref_off;
    const int32_t *ref_len;
    const int64_t *qry_off;
    const int32_t *qry_len;
    int32_t *status;
    const WorkItem *items;
    int n_items;
    const uint8_t *arena;
    const PairStart *start;
    const int32_t *aux;
    const int64_t *aux_off;
    const DeviceScoring *sc;
    int32_t *seg_count;        // per item
    const int64_t *seg_base;   // per item: exclusive prefix of seg_count (write pass)
    int32_t *segs;             // 5 ints per segment (write pass)
};

enum { MV_DIAG = 0, MV_UP = 1, MV_LEFT = 2 };

struct CodeView {
    const uint8_t *base;
    int cols, cb, W;
    int64_t nsteps;
    bool affine;
    __device__ __forceinline__ unsigned fetch(int i, int j) const
    {
        const int jj = j - 1;
        const int pass = jj / W;
        const int w = jj - pass * W;
        const int t = w / cols, k = w - t * cols;
        const int64_t g = ((int64_t)pass * nsteps + (i - 1 + t)) * 32 + t;
        if (affine) return base[g * cb + k];
        return (base[g * cb + (k >> 2)] >> (2 * (k & 3))) & 3u;
    }
};

template <int KIND, bool AFFINE, bool WRITE>
__global__ void ba_walk_kernel(WalkParams P)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= P.n_items) return;
    const WorkItem item = P.items[it];
    const int p = item.pair;
    if (P.status[p] != 0) {
        if (!WRITE) P.seg_count[it] = 0;
        return;
    }
    const int n = P.ref_len[p], m = P.qry_len[p];
    const uint8_t *rs = P.enc + P.ref_off[p];
    const uint8_t *qs = P.enc + P.qry_off[p];
    const int stride = P.stride;
    const int let = P.sc->let;
    const int32_t *la = P.sc->la;
    const int go = P.sc->gap_open;

    CodeView cv;
    cv.base = P.arena + item.code_off;
    cv.cols = item.cols;
    cv.cb = ba_code_group_bytes(AFFINE, item.cols);
    cv.W = 32 * item.cols;
    cv.nsteps = (int64_t)n + 31;
    cv.affine = AFFINE;

    // ---- start cell
    int i, j, layer = 0;
    int cur = 0;  // SW only: value of the current cell/layer, tracked by telescoping
    if (KIND == KIND_SW) {
        if (n > 0 && m > 0) {
            const PairStart st = P.start[p];
            i = st.i;
            j = st.j;
            cur = st.score;
        } else {
            i = j = 0;
        }
    } else if (KIND == KIND_NW) {
        i = n;
        j = m;
        if (AFFINE) layer = P.start[p].layer;
    } else {
        // fitted_letters.go:127-137 / fitted_affine_letters.go:170-181
        j = m;
        i = 0;
        if (n > 0) {
            const int32_t *lc = P.aux + P.aux_off[it];
            const int qv = qs[(int64_t)(m - 1) * stride];
            int best = INT32_MIN;
            for (int y = 1; y <= n; y++) {
                const int v = lc[y];
                bool ok = v >= best;
                if (!AFFINE) ok = ok && la[(int)rs[(int64_t)(y - 1) * stride] * let + qv] >= 0;
                if (ok) {
                    i = y;
                    best = v;
                }
            }
        }
    }

    int maxI = i, maxJ = j, score = 0, last = MV_DIAG;
    int nseg = 0;
    int32_t *out = WRITE ? P.segs + 5 * (P.seg_base[it] + P.seg_count[it] - 1) : nullptr;
    bool bad = false;

#define BA_EMIT(a0, a1, b0, b1, sc_)  \
    do {                              \
        if (WRITE) {                  \
            out[0] = (a0);            \
            out[1] = (a1);            \
            out[2] = (b0);            \
            out[3] = (b1);            \
            out[4] = (sc_);           \
            out -= 5;                 \
        }                             \
        nseg++;                       \
    } while (0)

    while (i > 0 && j > 0) {
        if (KIND == KIND_SW && cur == 0) break;
        const int R = rs[(int64_t)(i - 1) * stride], Q = qs[(int64_t)(j - 1) * stride];
        const unsigned code = cv.fetch(i, j);
        const bool corner = (i == n && j == m);
        int move, nl = 0, delta;
        if (AFFINE) {
            unsigned f = (layer == 0) ? (code & 7u) : (layer == 1) ? ((code >> 3) & 3u) : (code >> 5);
            const int er = la[R * let], eq = la[Q];
            switch (f) {
            case 1: move = MV_UP;   nl = 1; delta = er;      break;
            case 2: move = MV_LEFT; nl = 2; delta = eq;      break;
            case 3: move = MV_UP;   nl = 0; delta = go + er; break;
            case 4: move = MV_LEFT; nl = 0; delta = go + eq; break;
            case 5: move = MV_DIAG; nl = (KIND == KIND_SW) ? 0 : 1; delta = la[R * let + Q]; break;
            case 6: move = MV_DIAG; nl = (KIND == KIND_SW) ? 1 : 2; delta = la[R * let + Q]; break;
            case 7: move = MV_DIAG; nl = (KIND == KIND_SW) ? 2 : 0; delta = la[R * let + Q]; break;
            default: move = MV_DIAG; delta = 0; bad = true; break;
            }
        } else {
            switch (code) {
            case 3: move = MV_DIAG; delta = la[R * let + Q]; break;
            case 2: move = MV_UP;   delta = la[R * let];     break;
            case 1: move = MV_LEFT; delta = la[Q];           break;
            default: move = MV_DIAG; delta = 0; bad = true;  break;
            }
        }
        if (bad) break;
        // sw_letters.go has no corner guard; every other kernel guards up/left moves.
        const bool guard = (KIND == KIND_SW && !AFFINE) || move == MV_DIAG || !corner;
        if (last != move && guard) {
            BA_EMIT(i, maxI, j, maxJ, score);
            maxI = i;
            maxJ = j;
            score = 0;
        }
        score += delta;
        if (KIND == KIND_SW) cur -= delta;
        if (move != MV_LEFT) i--;
        if (move != MV_UP) j--;
        layer = nl;
        last = move;
    }
    BA_EMIT(i, maxI, j, maxJ, score);
    if (KIND == KIND_NW && i != j) {
        // nw_letters.go:183-189 / nw_affine_letters.go:311-325: leading gap, border value
        int b = AFFINE ? go : 0;
        if (i == 0)
            for (int x = 0; x < j; x++) b += la[qs[(int64_t)x * stride]];
        else
            for (int x = 0; x < i; x++) b += la[(int)rs[(int64_t)x * stride] * let];
        BA_EMIT(0, i, 0, j, b);
    }
#undef BA_EMIT
    if (bad) P.status[p] = 4;  // BA_PAIR_NO_PATH
    if (!WRITE) P.seg_count[it] = bad ? 0 : nseg;
}
