// ba_fill16.cuh -- packed 16-bit DPX fill for SWAffine with checkpointed traceback.
//
// Why a second fill kernel: ba_fill32 derives every cell's exact tie-break code, which costs
// ~50 issue slots per cell and makes it ALU-issue bound (profiles/r01_fill32_ncu_summary.txt).
// The walk only ever visits O(r+c) cells, so this kernel computes VALUES ONLY -- seven DPX /
// packed-integer instructions per TWO cells -- and stores sparse checkpoints of the table:
//   * the right-most column of every 32*COLS/32-wide strip, every row        (M, L)
//   * one row of every strip every BA16_PERIOD steps                          (M, U)
// ba_trace16 (ba_trace16.cuh) then re-derives the exact codes only for the 16-row x COLS-column
// tiles the traceback actually crosses, with the very same affine_cell<> as ba_fill32.
//
// Layout: a warp holds two pairs, one per 16-lane group.  Each lane carries TWO virtual
// threads in the halves of its 32-bit registers: low half = virtual thread t (strip t of the
// pass), high half = virtual thread t+16.  Virtual thread v handles row i at step s = i-1+v, so
// one warp shuffle (rotate within the group) hands the boundary column from strip v to v+1
// for both halves at once; lane 0 re-routes lane 15's low half into its own high half.
//
// Exactness conditions checked by the host before this path is chosen (else ba_fill32):
//   C1  every gap entry la[R][0] (resp. la[0][Q]) is one value er <= 0 (eq <= 0) for letters 1..4
//   C2  go + er + max(sub) < 0 and go + eq + max(sub) < 0   => every cell that attains the table
//       maximum has max3(diag) == diag M, i.e. qualifies for bíogo's best-cell rule
//       (align/sw_affine_letters.go:126-134), so "last cell in row-major order with the maximum
//       score" is the reference's start cell; U and L may skip the clamp at 0 inside the
//       recurrence (it is re-applied when checkpoints are written)
//   C3  letters restricted to alphabet indices 1..4 and |la| < 128: the substitution score of
//       a packed cell pair is ONE PRMT (byte permute with sign replication) of two 4-byte row
//       profiles
//   C4  min(n, m) * max(sub) < 2^15 - 256
#pragma once

#include "ba_common.cuh"

#define BA16_PROF_RING 512           // ring of row profiles per group (entries)
#define BA16_PROF_CHUNK 256          // rows loaded per refill
#define BA16_WARPS 4                 // warps per CTA
#define BA16_PERIOD 8                // steps between two row checkpoints (= tile rows of ba_trace16)
#define BA16_PSHIFT 3
#define BA16_GROUPS (2 * BA16_WARPS) // pairs in flight per CTA

struct Fill16Params {
    const uint8_t *enc;
    int stride;
    const int64_t *ref_off;
    const int32_t *ref_len;
    const int64_t *qry_off;
    const int32_t *qry_len;
    const int32_t *status;
    const uint8_t *acgt_only;  // per pair: letters are all alphabet indices 1..4 (else: exact path)
    const WorkItem *items;
    int n_items;
    int *counter;
    uint8_t *arena;          // checkpoints, WorkItem.code_off bytes into it
    int32_t *bnd;            // inter-pass boundary, bnd_stride ints per group
    int64_t bnd_stride;
    const DeviceScoring *sc;
};

// ---- packed helpers ---------------------------------------------------------------------
__device__ __forceinline__ unsigned ba_prmt(unsigned a, unsigned b, unsigned s)
{
#ifdef BA_EMULATE
    return emu_prmt(a, b, s);
#else
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
#endif
}
__device__ __forceinline__ unsigned pack16(int lo, int hi)
{
    return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16);
}
__device__ __forceinline__ int lo16(unsigned v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int hi16(unsigned v) { return (int)(short)(v >> 16); }

// Checkpoint geometry shared with ba_trace16 ------------------------------------------------
// Everything is organised in 32-byte SECTORS, one per (period e, virtual thread v), so that
// the traceback thread of a pair fetches a whole tile boundary with one or two DRAM sectors:
//   colck[e][v] = 8 steps   x {M, L} int16 : last column of strip v at steps 8e .. 8e+7
//   rowck[e][v] = 8 columns x {M, U} int16 : row of strip v processed at step 8e+7
// A period is BA16_PERIOD = 8 steps; one pass of one pair holds nper = ceil((n+31)/8) periods of
// 32 sectors for each of the two arrays (about 1 byte per DP cell for 8-column strips).
// The fill stages both through shared memory and writes each period as fully coalesced
// 16-byte stores.
__host__ __device__ inline int64_t ba16_nper(int64_t n) { return (n + 31 + BA16_PERIOD - 1) / BA16_PERIOD; }
__host__ __device__ inline int64_t ba16_colck_bytes(int64_t n) { return ba16_nper(n) * 32 * 32; }
__host__ __device__ inline int64_t ba16_rowck_bytes(int64_t n, int cols)
{
    (void)cols;
    return ba16_nper(n) * 32 * 32;
}
__host__ __device__ inline int64_t ba16_pass_bytes(int64_t n, int cols)
{
    return ba16_colck_bytes(n) + ba16_rowck_bytes(n, cols);
}
// After the passes: one candidate record per (pass, lane): {packed best score of the two
// strips, packed last period attaining it} -- the start-cell candidates of ba_trace16.
__host__ __device__ inline int64_t ba16_npass(int cols, int64_t m) { return (m + 32LL * cols - 1) / (32LL * cols); }
__host__ __device__ inline int64_t ba16_ck_bytes(int cols, int64_t n, int64_t m)
{
    if (n <= 0 || m <= 0) return 0;
    const int64_t np = ba16_npass(cols, m);
    return np * ba16_pass_bytes(n, cols) + np * 16 * 8;
}
// Columns are RIGHT-aligned: strip (pass, v) covers table columns
//   jshift + pass*W + v*COLS + 1 .. + COLS   with jshift = m - npass*W  (<= 0),
// so the last strip ends exactly at column m and the padding falls on columns <= 0, which
// behave as the all-zero SW border (substitution 0 or -1 keeps M at 0 there).

#define BA16_STAGE_ROW 17  // padded row length (entries) of the staging buffers

// Write one period of staged checkpoints: buf[j][lane] = {packed A, packed B} for j = 0..7 (step
// or column), lanes 0..15; the low halves belong to virtual thread `lane`, the high halves to
// `lane+16`.  Output: 32 sectors of 32 bytes, sector v = 8 x {A.half, B.half}; lane t writes the
// 16-byte chunks t, 16+t, 32+t, 48+t, i.e. every store instruction covers 256 contiguous bytes.
__device__ __forceinline__ void ba16_flush_period(const uint2 *buf, uint4 *gdst, int t)
{
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int c = q * 16 + t;
        const int v = c >> 1, hh = c & 1;
        const unsigned selw = (v & 16) ? 0x7632u : 0x5410u;
        const uint2 *src = buf + (4 * hh) * BA16_STAGE_ROW + (v & 15);
        const uint2 e0 = src[0], e1 = src[BA16_STAGE_ROW], e2 = src[2 * BA16_STAGE_ROW], e3 = src[3 * BA16_STAGE_ROW];
        gdst[c] = make_uint4(ba_prmt(e0.x, e0.y, selw), ba_prmt(e1.x, e1.y, selw), ba_prmt(e2.x, e2.y, selw),
                             ba_prmt(e3.x, e3.y, selw));
    }
}

// MULTI: the launch contains pairs whose query needs more than one 32*COLS-column pass (the
// inter-pass boundary code is compiled out of the common single-pass instance).
template <int COLS, bool MULTI>
__global__ void __launch_bounds__(BA16_WARPS * 32) ba_fill16_sw_kernel(Fill16Params P)
{
    // row profiles: ring[group][i & (RING-1)] = 4 substitution bytes of the letter of row i
    __shared__ unsigned ring[BA16_GROUPS][BA16_PROF_RING + 16];  // +16: the two groups of a warp hit disjoint banks
    __shared__ unsigned proftab[8];  // letter index -> {sub(R,1), sub(R,2), sub(R,3), sub(R,4)} as int8
    __shared__ uint2 colbuf[BA16_GROUPS][8 * BA16_STAGE_ROW];  // staged column checkpoints of the period
    __shared__ uint2 rowbuf[BA16_GROUPS][8 * BA16_STAGE_ROW];  // staged row checkpoint
    const int let = P.sc->let;
    if (threadIdx.x < 8) {
        unsigned w = 0x80808080u;  // virtual row: -128 against everything
        const int R = threadIdx.x;
        if (R >= 1 && R <= 4 && R < let) {
            w = 0;
            for (int q = 1; q <= 4; q++) w |= ((unsigned)(P.sc->la[R * let + q]) & 0xffu) << (8 * (q - 1));
        }
        proftab[threadIdx.x] = w;
    }
    __syncthreads();
    const int er = P.sc->la[1 * let];  // C1: identical for letters 1..4
    const int eq = P.sc->la[1];
    const int go = P.sc->gap_open;
    const unsigned erp = pack16(er, er), eqp = pack16(eq, eq);
    const unsigned goerp = pack16(go + er, go + er), goeqp = pack16(go + eq, go + eq);

    const int lane = threadIdx.x & 31;
    const int t = lane & 15;                       // lane inside the group
    const int grp = (threadIdx.x >> 4);            // group inside the CTA
    const int group_global = blockIdx.x * BA16_GROUPS + grp;
    unsigned *myring = ring[grp];
    uint2 *mycol = colbuf[grp];
    uint2 *myrow = rowbuf[grp];
    int32_t *bnd = P.bnd + (int64_t)group_global * P.bnd_stride;
    constexpr int W = 32 * COLS;
    const unsigned FULL = 0xffffffffu;
    const unsigned fixsel = (t == 0) ? 0x5410u : 0x7654u;  // lane 0: {boundary.lo, neighbour.lo}

    for (;;) {
        // each warp takes two consecutive items (one per group)
        int it0 = 0;
        if (lane == 0) it0 = atomicAdd(P.counter, 2);
        it0 = __shfl_sync(FULL, it0, 0);
        if (it0 >= P.n_items) break;
        const int it = it0 + (lane >> 4);
        const bool have = it < P.n_items;
        WorkItem item;
        item.pair = 0;
        item.cols = COLS;
        item.code_off = 0;
        if (have) item = P.items[it];
        const int p = item.pair;
        int n = have ? P.ref_len[p] : 0, m = have ? P.qry_len[p] : 0;
        if (have && (P.status[p] != 0 || !P.acgt_only[p])) n = 0;
        const bool live = have && n > 0 && m > 0;
        if (!live) {
            n = 0;
            m = 0;
        }
        const uint8_t *rptr = P.enc + (live ? P.ref_off[p] : 0);
        const uint8_t *qptr = P.enc + (live ? P.qry_off[p] : 0);
        const int stride = P.stride;
        const int npass = live ? (m + W - 1) / W : 0;
        const int nsteps = live ? n + 31 : 0;
        // loop bounds must be warp-uniform (shuffles): take the larger group
        const int nsteps_w = max(nsteps, __shfl_xor_sync(FULL, nsteps, 16));
        const int npass_w = MULTI ? max(npass, __shfl_xor_sync(FULL, npass, 16)) : 1;
        const int64_t pass_bytes = ba16_pass_bytes(n, COLS);


        for (int pass = 0; pass < npass_w; pass++) {
            const bool pass_live = live && pass < npass;
            const int jbase = pass * W;
            const int jshift = m - npass * W;
            // ---- per-column constants: PRMT selectors (columns <= 0: sign-select => sub in {0,-1})
            unsigned sel[COLS];
#pragma unroll
            for (int k = 0; k < COLS; k++) {
                const int jl = jshift + jbase + t * COLS + k + 1, jh = jl + 16 * COLS;
                unsigned nl = 0x88u, nh = 0xCC00u;  // virtual column: both bytes = sign of a profile byte
                if (pass_live && jl >= 1) {
                    const unsigned q = (unsigned)qptr[(int64_t)(jl - 1) * stride] - 1u;
                    nl = q | ((q | 8u) << 4);
                }
                if (pass_live && jh >= 1) {
                    const unsigned q = 4u + (unsigned)qptr[(int64_t)(jh - 1) * stride] - 1u;
                    nh = (q << 8) | ((q | 8u) << 12);
                }
                sel[k] = nl | nh;
            }
            // ---- (re)load the profile ring: rows -31..0 virtual, rows 1.. from the reference
            for (int x = t; x < 32; x += 16) myring[(x - 31) & (BA16_PROF_RING - 1)] = 0x80808080u;
            __syncwarp();

            unsigned Mu[COLS], Uu[COLS], Hu[COLS];
#pragma unroll
            for (int k = 0; k < COLS; k++) Mu[k] = Uu[k] = Hu[k] = 0;
            unsigned outM = 0, outL = 0, outH = 0, sdiagH = 0;
            unsigned bmax = 0;  // running max of the current period
            // per virtual thread: best value of this pass and the LAST period attaining it
            unsigned vbest = pack16(1, 1);
            int per_lo = -1, per_hi = -1;
            uint8_t *ck = P.arena + item.code_off + (int64_t)pass * pass_bytes;
            uint4 *colck = (uint4 *)ck;                              // [period][64 chunks of 16 bytes]
            uint4 *rowck = (uint4 *)(ck + ba16_colck_bytes(n));

            // byte offsets of my two rows inside the profile ring, advanced by one entry per step
            unsigned rolo = (unsigned)((1 - t) & (BA16_PROF_RING - 1)) * 4u;
            unsigned rohi = (unsigned)((1 - t - 16) & (BA16_PROF_RING - 1)) * 4u;
#pragma unroll 2
            for (int s = 0; s < nsteps_w; s++) {
                if ((s & (BA16_PROF_CHUNK - 1)) == 0) {
                    // rows s+1 .. s+CHUNK enter the ring (lanes of the group share the load)
                    __syncwarp();
                    for (int x = t; x < BA16_PROF_CHUNK; x += 16) {
                        const int i = s + 1 + x;
                        unsigned w = 0x80808080u;
                        if (pass_live && i <= n) w = proftab[rptr[(int64_t)(i - 1) * stride] & 7];
                        myring[i & (BA16_PROF_RING - 1)] = w;
                    }
                    __syncwarp();
                }
                const int ilo = s + 1 - t;
                const unsigned profA = *(const unsigned *)((const char *)myring + rolo);
                const unsigned profB = *(const unsigned *)((const char *)myring + rohi);
                rolo = (rolo + 4u) & (BA16_PROF_RING * 4u - 1u);
                rohi = (rohi + 4u) & (BA16_PROF_RING * 4u - 1u);

                // boundary column of the left neighbour strip (rotate inside the 16-lane group)
                unsigned rM = __shfl_sync(FULL, outM, (t + 15) & 15, 16);
                unsigned rL = __shfl_sync(FULL, outL, (t + 15) & 15, 16);
                unsigned rH = __shfl_sync(FULL, outH, (t + 15) & 15, 16);
                unsigned bM = 0, bL = 0, bH = 0;  // pass 0: DP column 0 is all zero for SW
                if (MULTI && pass > 0) {  // warp-uniform
                    if (t == 0 && ilo >= 1 && ilo <= n) {
                        bM = (unsigned)bnd[3 * ilo + 0];
                        bL = (unsigned)bnd[3 * ilo + 1];
                        bH = (unsigned)bnd[3 * ilo + 2];
                    }
                }
                rM = ba_prmt(bM, rM, fixsel);
                rL = ba_prmt(bL, rL, fixsel);
                rH = ba_prmt(bH, rH, fixsel);

                unsigned Ml = rM, Ll = rL, Hd = sdiagH;
#pragma unroll
                for (int k = 0; k < COLS; k++) {
                    const unsigned sub = ba_prmt(profA, profB, sel[k]);
                    const unsigned Mn = __viaddmax_s16x2_relu(Hd, sub, erp);  // erp <= 0: same as max(.,0)
                    const unsigned c1 = __vadd2(Uu[k], erp);
                    const unsigned Un = __viaddmax_s16x2(Mu[k], goerp, c1);
                    const unsigned c2 = __vadd2(Ll, eqp);
                    const unsigned Ln = __viaddmax_s16x2(Ml, goeqp, c2);
                    Hd = Hu[k];
                    Hu[k] = __vimax3_s16x2(Mn, Un, Ln);
                    Mu[k] = Mn;
                    Uu[k] = Un;
                    Ml = Mn;
                    Ll = Ln;
                }
#pragma unroll
                for (int k = 0; k + 1 < COLS; k += 2) bmax = __vimax3_s16x2(bmax, Mu[k], Mu[k + 1]);
                if (COLS & 1) bmax = __vmaxs2(bmax, Mu[COLS - 1]);
                sdiagH = rH;
                outM = Mu[COLS - 1];
                outL = Ll;
                outH = Hu[COLS - 1];

                // column checkpoint: (M, clamped L) of my last column, staged for this period
                mycol[(s & 7) * BA16_STAGE_ROW + t] = make_uint2(outM, __vimax_s16x2_relu(outL, erp));
                if (MULTI && pass_live && s < nsteps && t == 15 && pass + 1 < npass) {
                    const int ihi = ilo - 16;
                    if (ihi >= 1 && ihi <= n) {
                        bnd[3 * ihi + 0] = hi16(outM) & 0xffff;
                        bnd[3 * ihi + 1] = hi16(outL) & 0xffff;
                        bnd[3 * ihi + 2] = hi16(outH) & 0xffff;
                    }
                }
                if ((s & (BA16_PERIOD - 1)) == BA16_PERIOD - 1) {
                    const int e = s >> BA16_PSHIFT;
#pragma unroll
                    for (int k = 0; k < COLS; k++)
                        myrow[k * BA16_STAGE_ROW + t] = make_uint2(Mu[k], __vimax_s16x2_relu(Uu[k], erp));
                    __syncwarp();
                    if (pass_live && (e << BA16_PSHIFT) < nsteps) ba16_flush_period(mycol, colck + (int64_t)e * 64, t);
                    if (pass_live && s < nsteps) ba16_flush_period(myrow, rowck + (int64_t)e * 64, t);
                    __syncwarp();
                    // best tracking at period granularity ('>=' keeps the LAST period)
                    bool ph, pl;
                    vbest = __vibmax_s16x2(bmax, vbest, &ph, &pl);
                    if (pl) per_lo = e;
                    if (ph) per_hi = e;
                    bmax = 0;
                }
            }
            // the tail period (a partial one): its column checkpoints, and the best cell
            __syncwarp();
            if ((nsteps_w & (BA16_PERIOD - 1)) != 0 && pass_live &&
                ((nsteps_w >> BA16_PSHIFT) << BA16_PSHIFT) < nsteps)
                ba16_flush_period(mycol, colck + (int64_t)(nsteps_w >> BA16_PSHIFT) * 64, t);
            if (nsteps_w & (BA16_PERIOD - 1)) {
                const int e = nsteps_w >> BA16_PSHIFT;
                bool ph, pl;
                vbest = __vibmax_s16x2(bmax, vbest, &ph, &pl);
                if (pl) per_lo = e;
                if (ph) per_hi = e;
            }
            if (pass_live) {
                uint2 *cand = (uint2 *)(P.arena + item.code_off + (int64_t)npass * pass_bytes) + pass * 16 + t;
                *cand = make_uint2(vbest, pack16(per_lo, per_hi));
            }
            __syncwarp();
        }

    }
}
