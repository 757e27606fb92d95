// ba_fill16.cuh -- packed 16-bit DPX fill for the affine aligners with checkpointed traceback.
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
// Layout: a warp holds FOUR pairs, two per 16-lane group: the low 16-bit halves of every
// register belong to pair A, the high halves to pair B (two pairs of the same size class,
// neighbours in the sorted work list).  Lane t owns 2*COLS consecutive columns of a 32*COLS wide
// pass (two "strips" of COLS columns: 2t and 2t+1) and handles row i at step s = i-1+t; one
// SHFL.UP per boundary register hands the right-most column to lane t+1 for both pairs at once.
//
// Exactness conditions checked by the host before this path is chosen (else ba_fill32):
//   C1  every gap entry la[R][0] (resp. la[0][Q]) is one value er <= 0 (eq <= 0) for letters 1..4
//   C2  go + er + max(sub) < 0 and go + eq + max(sub) < 0   => every cell that attains the table
//       maximum has max3(diag) == diag M, i.e. qualifies for bíogo's best-cell rule
//       (align/sw_affine_letters.go:126-134), so "last cell in row-major order with the maximum
//       score" is the reference's start cell; U and L may skip the clamp at 0 inside the
//       recurrence (it is re-applied when ba_trace16 reads the checkpoints)
//   C3  letters restricted to alphabet indices 1..4 and |la| < 128: the substitution score of
//       a packed cell pair is ONE PRMT (byte permute with sign replication) of two 4-byte row
//       profiles
//   C4  min(n, m) * max(sub) < 2^15 - 256
#pragma once

#include "ba_common.cuh"

#ifndef BA16_PROF_RING
#define BA16_PROF_RING 512           // ring of row profiles per group (entries)
#endif
#define BA16_PROF_CHUNK (BA16_PROF_RING / 2)   // rows loaded per refill
#define BA16_WARPS 2                 // warps per CTA
#define BA16_PERIOD 8                // steps between two row checkpoints (= tile rows of ba_trace16)
#define BA16_PSHIFT 3
#define BA16_GROUPS (2 * BA16_WARPS) // 16-lane groups per CTA (two pairs each)
#define BA16_SKEW 15                 // steps of systolic skew: nsteps = n + BA16_SKEW
#define BA16_NEG (-20000)            // "minus infinity" of the NW/Fitted kernels (finite values stay above -14000)

struct Fill16Params {
    const uint8_t *enc;
    int stride;
    const int64_t *ref_off;
    const int32_t *ref_len;
    const int64_t *qry_off;
    const int32_t *qry_len;
    const int32_t *status;
    const uint8_t *acgt_only;  // per pair letter flags: bit0 = all indices in 1..4, bit1 = no index 0
    int flag_mask;             // which of those bits this launch requires (else: exact path)
    const WorkItem *items;
    int n_items;
    int *counter;
    uint8_t *arena;          // checkpoints, WorkItem.code_off bytes into it
    int32_t *bnd;            // inter-pass boundary rows (16 bytes each), bnd_stride ints per group (LONG: per couple)
    int64_t bnd_stride;
    // LONG (wavefront over the passes of kb-scale pairs): items are (pair, pass) tasks in pass-major
    // order, 2*long_couples per pass; the boundary rows carry the writer's pass as a tag, so the
    // reader needs neither flags nor fences (bnd is zeroed before the launch)
    int long_couples;
    int *long_err;           // set when a wait timed out (never expected)
    // 0: persistent CTAs (loop until the work counter runs out).  k > 0: a warp takes at most k
    // claims and exits, so that the grid is larger than what is resident and CTAs of other,
    // higher-priority kernels (the traceback of the previous wave) get SM slots as these retire.
    int max_claims;
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
// A COUPLE (pairs A and B of a 16-lane group: A in the low, B in the high 16-bit halves of every register)
// shares one checkpoint region, and the fill stores its packed registers AS THEY ARE: no instruction is
// spent on separating the two pairs (round 1 stored per pair and paid 64 PRMT + 16 VSUB per 8-step period
// and lane for it -- 6.5 % of the kernel's ALU work).  Per pass and period e (BA16_PERIOD = 8 steps; strip
// g = 2*lane + s belongs to lane g>>1, which handles row i at step i-1+lane):
//   column block  [lane 0..15][strip s][step j] x element   last column of the strip at step 8e+j
//                 affine element = {M word, L word} (8 bytes), linear element = {T word}; linear interleaves
//                 the strips ([lane][step][strip]) so that a step is one 8-byte store
//   row block     [lane 0..15][column kk = s*COLS + k] x element   the lane's row at step 8e+7
//                 affine element = {M (or G = M + go + e for the SYM kernels) word, U word}, linear {T word};
//                 a lane's row is RS = ba16_row_stride(COLS) double-elements long
// so the traceback thread of one pair finds a tile's left column in one 64-byte piece (two sectors, which it
// shares with its couple partner -- the neighbouring thread) and its top row in another.  Raw values: SW's
// deferred clamp of U and L is re-applied by the reader, as is the SYM bias.
// The fill stages a period in shared memory lane-major with an odd row stride (conflict-free stores) and
// flushes it with fully coalesced 8/16-byte copies whose addresses are per-lane constants plus immediates.
__host__ __device__ inline int64_t ba16_nper(int64_t n) { return (n + BA16_SKEW + BA16_PERIOD - 1) / BA16_PERIOD; }
// double-elements (two columns) per lane row in GLOBAL memory: COLS, except that an even COLS that is not a
// power of two gets one pad (then the staged rows can be copied linearly)
__host__ __device__ constexpr int ba16_row_stride(int cols)
{
    return ((cols & 1) || (cols & (cols - 1)) == 0) ? cols : cols + 1;
}
__host__ __device__ inline int64_t ba16_colblk(bool lin) { return lin ? 1024 : 2048; }
__host__ __device__ inline int64_t ba16_rowblk(int cols, bool lin) { return 16LL * ba16_row_stride(cols) * (lin ? 8 : 16); }
__host__ __device__ inline int64_t ba16_colck_bytes(int64_t n, bool lin) { return ba16_nper(n) * ba16_colblk(lin); }
__host__ __device__ inline int64_t ba16_rowck_bytes(int cols, int64_t n, bool lin) { return ba16_nper(n) * ba16_rowblk(cols, lin); }
__host__ __device__ inline int64_t ba16_pass_bytes(int cols, int64_t n, bool lin)
{
    return ba16_colck_bytes(n, lin) + ba16_rowck_bytes(cols, n, lin);
}
// After the passes: per (pass, lane) one candidate record per pair {best score of the lane's two strips, last
// period attaining it} -- the start-cell candidates of ba_trace16 (A at [2x], B at [2x+1]).
__host__ __device__ inline int64_t ba16_npass(int cols, int64_t m) { return (m + 32LL * cols - 1) / (32LL * cols); }
// bytes of a couple's region: n = rows (max over the couple), np = passes (max over the couple)
__host__ __device__ inline int64_t ba16_ck_bytes(int cols, int64_t n, int64_t np, bool lin)
{
    if (n <= 0 || np <= 0) return 0;
    return np * ba16_pass_bytes(cols, n, lin) + np * 16 * 16;
}
// Columns are RIGHT-aligned: strip g of pass p covers table columns
//   jshift + p*W + g*COLS + 1 .. + COLS   with W = 32*COLS, jshift = m - npass*W  (<= 0),
// so the last strip ends exactly at column m and the padding falls on columns <= 0, which
// behave as the all-zero SW border (substitution 0 or -1 keeps M at 0 there).  (Per PAIR: each pair of a
// couple has its own m, npass and jshift; only the region's size is common.)

// Staging (shared memory, per group): lane-major rows with an odd stride.
//   column stage: 16 (affine) / 8 (linear) elements of 8 bytes per lane, row stride 17 / 9 elements
//   row stage:    COLS double-elements of 16 (affine) / 8 (linear) bytes per lane, row stride COLS|1
#define BA16_COLST_WORDS (16 * 17 * 2)
__host__ __device__ constexpr int ba16_rowst_stride(int cols) { return cols | 1; }

// MULTI: the launch contains pairs whose query needs more than one 32*COLS-column pass (the
// inter-pass boundary code is compiled out of the common single-pass instance).
// Dynamic shared memory (GSUB score tables); the emulator backs it with a static buffer.
// A replica serves BA16_GSUB_SHARE lanes and is confined to as many banks: byte (R, q) of replica r
// lives at ((R*32 + q) / (4*SHARE)) * 128 + 4*SHARE*r + (R*32 + q) % (4*SHARE) of the warp's slab.
#ifndef BA16_GSUB_SHARE
#define BA16_GSUB_SHARE 2
#endif
#define BA16_GSUB_SPAN (4 * BA16_GSUB_SHARE)            // bytes of a replica per 128-byte bank row
__host__ __device__ inline size_t ba16_gsub_bytes_per_warp(int let)
{
    return (size_t)(let + 1) * 32 / BA16_GSUB_SPAN * 128;
}
#ifdef BA_EMULATE
static inline uint8_t *ba_dyn_smem()
{
    static uint8_t buf[80 * 1024];   // GSUB score tables of the fill, value tiles of the affine traceback
    return buf;
}
#else
__device__ __forceinline__ uint8_t *ba_dyn_smem()
{
    extern __shared__ __align__(16) uint8_t ba_dyn_smem_[];
    return ba_dyn_smem_;
}
#endif

#ifdef BA_EMULATE
#define BA_SPIN_PAUSE() emu::yield()
#define BA_LDCG(p) (*(p))
#else
#define BA_SPIN_PAUSE() __nanosleep(64)
#define BA_LDCG(p) __ldcg(p)
#endif

// Inter-pass boundary row: the last column's (M, L, H) of a pass for both pairs of a couple,
// 16 bytes = two 64-bit words that each carry the writer's tag in their top 16 bits:
//   w0 = M (A|B packed) | L_A << 32 | tag << 48      w1 = H (A|B packed) | L_B << 32 | tag << 48
// A 64-bit access is single-copy atomic, so a reader that sees both tags sees the whole row.
__device__ __forceinline__ ulonglong2 ba16_bnd_pack(unsigned M, unsigned L, unsigned H, unsigned tag)
{
    ulonglong2 w;
    w.x = (unsigned long long)M | ((unsigned long long)((L & 0xffffu) | (tag << 16)) << 32);
    w.y = (unsigned long long)H | ((unsigned long long)((L >> 16) | (tag << 16)) << 32);
    return w;
}
__device__ __forceinline__ bool ba16_bnd_ready(const ulonglong2 &w, unsigned tag)
{
    return (unsigned)(w.x >> 48) == tag && (unsigned)(w.y >> 48) == tag;
}
__device__ __forceinline__ uint4 ba16_bnd_unpack(const ulonglong2 &w)
{
    const unsigned hx = (unsigned)(w.x >> 32), hy = (unsigned)(w.y >> 32);
    return make_uint4((unsigned)w.x, (hx & 0xffffu) | (hy << 16), (unsigned)w.y, 0u);
}

// KIND: KIND_SW (right-aligned columns, zero borders, clamps, best-cell candidates) or KIND_NW /
// KIND_FIT (left-aligned, -inf/gap borders, no clamps).  GSUB: substitution scores from two
// shared-memory tables (any alphabet, e.g. Protein) instead of the 4-letter PRMT profiles.
// LONG: intra-pair wavefront for kb-scale pairs -- every work item is ONE pass of a pair couple;
// the passes of a couple run concurrently on different warps (anywhere on the GPU), each trailing
// its left neighbour by about three periods; the boundary column travels through a per-couple
// global row buffer whose rows are tagged with the writer's pass.  Items are claimed in pass-major order by a persistent
// grid, so the pass a warp waits for is always held by a warp that is already running.
// LIN: the linear-gap aligners (align/sw_letters.go, nw_letters.go, fitted_letters.go): one table T,
// T = max3(diag + sub, up + er, left + eq) [clamped at 0 for SW]; it lives in the H registers, the
// checkpoints carry it in their M field (the second field is zero).  SW needs er < 0 and eq < 0
// so that every cell holding the table maximum came through the diagonal (best-cell rule,
// align/sw_letters.go:110-114).
// SYM (affine, er == eq, go + er <= 0 -- every BASELINE scoring): the M array holds G = M + go + e
// instead of M, because both gap recurrences want exactly that sum:
//   U' = max(U + e, G_prev_row)     L = max(L_left + e, G_left)     one add-max each,
// so a packed cell costs PRMT, add-max (M), add (G), add-max (U), add-max (L), max3 (H) = 6
// instructions (+ half a max3 of best tracking) instead of 7.5.  M itself is transient inside
// the step; checkpoints recover it as G - (go + e).
template <int KIND, int COLS, bool MULTI, bool GSUB, bool LONG, bool LIN, bool SYM = false>
__global__ void __launch_bounds__(BA16_WARPS * 32)
    ba_fill16_kernel(Fill16Params P)
{
    constexpr int NC = 2 * COLS;  // columns per lane
    // row profiles: ring[group][pair][i & (RING-1)] = 4 substitution bytes of the letter of row i
    // GSUB kernels trade ring capacity for their score tables (shared memory bounds their occupancy)
    // (so do the linear kernels, whose small register footprint leaves shared memory as the limit)
    constexpr int RING = GSUB ? 128 : LIN ? 256 : BA16_PROF_RING, CHUNK = GSUB ? 64 : LIN ? 128 : BA16_PROF_CHUNK;
    __shared__ unsigned ring[BA16_GROUPS][2][RING + 8];
    __shared__ unsigned proftab[BA_MAX_LET + 1];  // PRMT: letter -> 4 sub bytes; GSUB: letter -> row byte offset
    // GSUB: signed-byte score tables in dynamic shared memory, replicated per BA16_GSUB_SHARE lanes and
    // confined to as many banks, so a lookup meets at most SHARE - 1 other lanes instead of 31;
    // row `let` is the virtual letter (-128 against everything).
    uint8_t *gtab = GSUB ? ba_dyn_smem() + (size_t)(threadIdx.x >> 5) * ba16_gsub_bytes_per_warp(P.sc->let) : nullptr;
    // staged checkpoints of the current period, per group (couple): lane-major rows, odd row stride
    constexpr int CST = LIN ? 9 : 17;                         // column stage: 8-byte elements per lane row
    constexpr int RST = ba16_rowst_stride(COLS);              // row stage: double-elements per lane row
    constexpr int RS = ba16_row_stride(COLS);                 // ... and in global memory
    constexpr int REW = LIN ? 2 : 4;                          // words per row double-element
    __shared__ __align__(16) unsigned colbuf[BA16_GROUPS][16 * CST * 2];
    __shared__ __align__(16) unsigned rowbuf[BA16_GROUPS][16 * RST * REW];
    __shared__ uint4 bbuf[MULTI ? BA16_GROUPS : 1][BA16_PERIOD];  // boundary rows (M, L, H) of the current period
    const int let = P.sc->let;
    if (!GSUB) {
        if (threadIdx.x < 8) {
            unsigned w = 0x80808080u;  // virtual row: -128 against everything
            const int R = threadIdx.x;
            if (R >= 1 && R <= 4 && R < let) {
                w = 0;
                for (int q = 1; q <= 4; q++) w |= ((unsigned)(P.sc->la[R * let + q]) & 0xffu) << (8 * (q - 1));
            }
            proftab[threadIdx.x] = w;
        }
    } else {
        constexpr int NREP = 32 / BA16_GSUB_SHARE;
        for (int x = threadIdx.x & 31; x < (let + 1) * 32 * NREP; x += 32) {
            const int r = x % NREP, bo = x / NREP, R = bo >> 5, q = bo & 31;
            const int v = (R == let) ? -128 : (q < let ? P.sc->la[R * let + q] : 0);
            gtab[(bo / BA16_GSUB_SPAN) * 128 + BA16_GSUB_SPAN * r + (bo % BA16_GSUB_SPAN)] = (uint8_t)v;
        }
        for (int x = threadIdx.x; x <= BA_MAX_LET; x += blockDim.x)
            proftab[x] = (unsigned)(min(x, let) * (32 / BA16_GSUB_SPAN * 128));
    }
    const unsigned VROW = GSUB ? (unsigned)(let * (32 / BA16_GSUB_SPAN * 128)) : 0x80808080u;  // ring entry of a virtual row
    __syncthreads();
    const int er = P.sc->la[1 * let];  // C1: identical for letters 1..4
    const int eq = P.sc->la[1];
    const int go = LIN ? 0 : P.sc->gap_open;  // linear borders: j*eq and i*er
    const unsigned erp = pack16(er, er), eqp = pack16(eq, eq);
    const unsigned goerp = pack16(go + er, go + er), goeqp = pack16(go + eq, go + eq);
    const unsigned NEGP = pack16(BA16_NEG, BA16_NEG);
    // BIAS (the SYM SW kernels): every value of the tables carries +vb, vb = -(go + e) >= 0.  Then M >= vb in both
    // halves of a packed word, G = M + go + e = M - vb >= 0, and the packed add is ONE 32-bit subtraction that never
    // borrows between the halves -- an integer add ptxas may issue on the FMA pipe (IMAD.IADD) instead of the
    // saturated ALU pipe that every DPX instruction needs.  Zero borders become vb, the clamp of M is the third
    // operand of the add-max, checkpoints hold biased values (ba_trace16 CkView subtracts vb), the start-cell
    // candidates are stored unbiased.  The host bounds vb (g_sym_scoring) so that C4's head room covers it.
    constexpr bool BIAS = SYM && KIND == KIND_SW && !LIN;
    const int vb = BIAS ? -(go + er) : 0;
    const unsigned vbp = pack16(vb, vb);

    const int lane = threadIdx.x & 31;
    const int t = lane & 15;                       // lane inside the group
    const int grp = (threadIdx.x >> 4);            // group inside the CTA
    const int group_global = blockIdx.x * BA16_GROUPS + grp;
    unsigned *ringA = ring[grp][0], *ringB = ring[grp][1];
    unsigned *colst = colbuf[grp] + t * CST * 2;   // my row of the column stage
    unsigned *rowst = rowbuf[grp] + t * RST * REW;   // my row of the row stage
    ulonglong2 *bnd = (ulonglong2 *)(P.bnd + (int64_t)group_global * P.bnd_stride);  // LONG: re-pointed per task
    constexpr int W = 32 * COLS;
    const unsigned FULL = 0xffffffffu;

    for (int claim = 0; P.max_claims == 0 || claim < P.max_claims; claim++) {
        // each warp takes four consecutive items: two (pairs A and B) per 16-lane group
        int it0 = 0;
        if (lane == 0) it0 = atomicAdd(P.counter, 4);
        it0 = __shfl_sync(FULL, it0, 0);
        if (it0 >= P.n_items) break;
        const int itA = it0 + 2 * (lane >> 4), itB = itA + 1;
        int long_pass = 0, couple = 0;
        if (LONG) {
            long_pass = itA / (2 * P.long_couples);
            couple = (itA % (2 * P.long_couples)) >> 1;
            bnd = (ulonglong2 *)(P.bnd + (int64_t)couple * P.bnd_stride);
        }
        int nA = 0, mA = 0, nB = 0, mB = 0, pA = 0, pB = 0;
        int64_t offC = 0;          // the couple's checkpoint region (both items carry the same offset and geometry)
        int n_ck = 0, np_ck = 0;
        bool liveA = false, liveB = false;
        if (itA < P.n_items) {
            const WorkItem item = P.items[itA];
            pA = max(item.pair, 0);
            offC = item.code_off;
            n_ck = item.n_ck;
            np_ck = item.ck & 0xffff;
            nA = P.ref_len[pA];
            mA = P.qry_len[pA];
            liveA = item.pair >= 0 && P.status[pA] == 0 && (P.acgt_only[pA] & P.flag_mask) == P.flag_mask && nA > 0 && mA > 0;
        }
        if (itB < P.n_items) {
            const WorkItem item = P.items[itB];
            pB = max(item.pair, 0);
            if (itA >= P.n_items || P.items[itA].pair < 0) {   // (LONG task lists may hold B without A)
                offC = item.code_off;
                n_ck = item.n_ck;
                np_ck = item.ck & 0xffff;
            }
            nB = P.ref_len[pB];
            mB = P.qry_len[pB];
            liveB = item.pair >= 0 && P.status[pB] == 0 && (P.acgt_only[pB] & P.flag_mask) == P.flag_mask && nB > 0 && mB > 0;
        }
        if (!liveA) nA = mA = 0;
        if (!liveB) nB = mB = 0;
        const uint8_t *rA = P.enc + (liveA ? P.ref_off[pA] : 0), *qA = P.enc + (liveA ? P.qry_off[pA] : 0);
        const uint8_t *rB = P.enc + (liveB ? P.ref_off[pB] : 0), *qB = P.enc + (liveB ? P.qry_off[pB] : 0);
        const int stride = P.stride;
        const int npassA = liveA ? (mA + W - 1) / W : 0, npassB = liveB ? (mB + W - 1) / W : 0;
        const int nstepsA = liveA ? nA + BA16_SKEW : 0, nstepsB = liveB ? nB + BA16_SKEW : 0;
        // loop bounds must be warp-uniform (shuffles): take the largest of the four pairs
        int nsteps_w = max(nstepsA, nstepsB), npass_w = max(npassA, npassB);
        nsteps_w = max(nsteps_w, __shfl_xor_sync(FULL, nsteps_w, 16));
        npass_w = MULTI ? max(npass_w, __shfl_xor_sync(FULL, npass_w, 16)) : 1;
        const int64_t passb = ba16_pass_bytes(COLS, n_ck, LIN);
        const int nstepsC = max(nstepsA, nstepsB);   // the couple's region is written while either pair is live
        const int jshiftA = (KIND == KIND_SW) ? mA - npassA * W : 0, jshiftB = (KIND == KIND_SW) ? mB - npassB * W : 0;
        // rows staged between passes: only pairs that really have a next pass (bounds the scratch)
        const int nbnd = max(npassA > 1 ? nA : 0, npassB > 1 ? nB : 0);

        for (int pass = (LONG ? long_pass : 0); pass < (LONG ? long_pass + 1 : npass_w); pass++) {
            const bool plA = liveA && pass < npassA, plB = liveB && pass < npassB;
            const bool rd_bnd = MULTI && pass > 0 && (plA || plB);               // my column 0 comes from bnd
            const bool wr_bnd = MULTI && pass + 1 < max(npassA, npassB);          // the next pass reads my last column
            // boundary rows are fetched one period ahead by lanes 0..7 of the group (row 8e+1+t)
            ulonglong2 pre = make_ulonglong2(0ull, 0ull);
#define BA16_BND_LAND(row_)                                                       \
    do {                                                                          \
        if (LONG) {                                                               \
            int spins_ = 0;                                                       \
            while (!ba16_bnd_ready(pre, (unsigned)pass)) {                        \
                BA_SPIN_PAUSE();                                                  \
                if (++spins_ > (1 << 22) || *(volatile int *)P.long_err != 0) {   \
                    *P.long_err = 1;                                              \
                    break;                                                        \
                }                                                                 \
                pre = BA_LDCG(bnd + (row_));                                      \
            }                                                                     \
        }                                                                         \
        bbuf[grp][t] = ba16_bnd_unpack(pre);                                      \
    } while (0)
            if (rd_bnd && t < BA16_PERIOD && 1 + t <= nbnd) {
                pre = BA_LDCG(bnd + 1 + t);
                BA16_BND_LAND(1 + t);
            }
            // ---- per-column constants.  PRMT path: byte selectors (virtual columns <= 0 of SW:
            // sign-select => sub in {0,-1}).  GSUB path: byte offsets of the query letters in a
            // table row (SW virtual columns use letter 0 = gap: any finite entry keeps M <= 0 only
            // approximately, so SW+GSUB clamps them through the profile row instead -- see below).
            unsigned sel[NC];
            unsigned qb4[GSUB ? NC : 1];
#pragma unroll
            for (int k = 0; k < NC; k++) {
                const int ja = jshiftA + pass * W + t * NC + k + 1, jb = jshiftB + pass * W + t * NC + k + 1;
                const bool oka = plA && ja >= 1 && ja <= mA, okb = plB && jb >= 1 && jb <= mB;
                if (!GSUB) {
                    unsigned nl = 0x88u, nh = 0xCC00u;  // virtual column: both bytes = sign of a profile byte
                    if (KIND != KIND_SW) {
                        nl = 0x80u;                     // padding column of NW/Fitted: any letter will do
                        nh = 0xC400u;
                    }
                    if (oka) {
                        const unsigned q = (unsigned)qA[(int64_t)(ja - 1) * stride] - 1u;
                        nl = q | ((q | 8u) << 4);
                    }
                    if (okb) {
                        const unsigned q = 4u + (unsigned)qB[(int64_t)(jb - 1) * stride] - 1u;
                        nh = (q << 8) | ((q | 8u) << 12);
                    }
                    sel[k] = nl | nh;
                } else {
                    // byte offset of the query letter inside a table row of my replica (lane couple)
                    const unsigned la_ = oka ? (unsigned)qA[(int64_t)(ja - 1) * stride] : 0u;
                    const unsigned lb_ = okb ? (unsigned)qB[(int64_t)(jb - 1) * stride] : 0u;
                    const unsigned rep_ = (unsigned)BA16_GSUB_SPAN * (unsigned)(lane / BA16_GSUB_SHARE);
                    sel[k] = (la_ / BA16_GSUB_SPAN) * 128u + (la_ % BA16_GSUB_SPAN) + rep_;
                    qb4[k] = (lb_ / BA16_GSUB_SPAN) * 128u + (lb_ % BA16_GSUB_SPAN) + rep_;
                }
            }
            // ---- profile rings: rows -15..0 virtual, rows 1.. loaded in chunks inside the step loop
            ringA[(t - 15) & (RING - 1)] = VROW;
            ringB[(t - 15) & (RING - 1)] = VROW;
            __syncwarp();

            unsigned Mu[NC], Uu[NC], Hu[NC];
            // the right-most column handed to the next lane is (Mu[NC-1], outL, Hu[NC-1]): M and H stay in
            // their arrays until the next step's shuffles have read them
            unsigned outL = 0, sdiagH = 0, mlast = 0;  // mlast (SYM): M of my last column at the current row
#pragma unroll
            for (int k = 0; k < NC; k++) {
                if (KIND == KIND_SW) {
                    Mu[k] = BIAS ? 0u : SYM ? goerp : 0u;   // SYM: Mu holds M + go + e (BIAS: + vb, i.e. zero)
                    Uu[k] = Hu[k] = vbp;
                } else {
                    // row 0: M = U = -inf, L = go + j*eq (nw_affine_letters.go:119-131); H = L
                    const int j = pass * W + t * NC + k + 1;
                    Mu[k] = SYM ? __vadd2(NEGP, goerp) : NEGP;
                    Uu[k] = NEGP;
                    Hu[k] = pack16(go + j * eq, go + j * eq);
                }
            }
            if (KIND != KIND_SW) {
                const int j0 = pass * W + t * NC;  // the column left of my first one: H[0][j0]
                sdiagH = (j0 == 0) ? 0u : pack16(go + j0 * eq, go + j0 * eq);
                outL = NEGP;
            } else if (BIAS) {
                outL = sdiagH = mlast = vbp;
            }
            unsigned bmax = 0;  // running max of the current period
            // per lane (both pairs packed): best value of this pass and the LAST period attaining it
            unsigned vbest = pack16(vb + 1, vb + 1);
            int per_lo = -1, per_hi = -1;
            const bool plC = plA || plB;
            uint8_t *colck = P.arena + offC + (int64_t)pass * passb;
            uint8_t *rowck = colck + ba16_colck_bytes(n_ck, LIN);
            unsigned ro = (unsigned)((1 - t) & (RING - 1)) * 4u;  // byte offset of my row in the rings
            unsigned ckM0 = 0, ckL0 = 0; // affine: (M, L) of my first strip's last column, as they are (A | B packed)
            const int nper_w = (nsteps_w + BA16_PERIOD - 1) >> BA16_PSHIFT;
            // letters of the first ring chunk (prefetching variant, see the refill below)
            constexpr int PFN = (CHUNK <= 64) ? CHUNK / 16 : 0;
            unsigned pfA[PFN > 0 ? PFN : 1], pfB[PFN > 0 ? PFN : 1];
            if (PFN > 0) {
#pragma unroll
                for (int y = 0; y < (PFN > 0 ? PFN : 1); y++) {
                    const int i = 1 + t + 16 * y;
                    pfA[y] = (plA && i <= nA) ? rA[(int64_t)(i - 1) * stride] : 0;
                    pfB[y] = (plB && i <= nB) ? rB[(int64_t)(i - 1) * stride] : 0;
                }
            }
            for (int e = 0; e < nper_w; e++) {
                const int pre_row = 8 * (e + 1) + 1 + t;  // my boundary row of the NEXT period
                const bool pre_on = rd_bnd && t < BA16_PERIOD && pre_row <= nbnd;
                if (pre_on) pre = BA_LDCG(bnd + pre_row);  // lands at the end of this period
                if ((e & (CHUNK / BA16_PERIOD - 1)) == 0) {
                    // rows 8e+1 .. 8e+CHUNK enter the rings (lanes of the group share the load)
                    __syncwarp();
                    if (PFN > 0) {
                        // small chunks (GSUB): the letters were fetched at the previous refill, the
                        // next chunk's are requested now and land while this one is consumed
#pragma unroll
                        for (int y = 0; y < (PFN > 0 ? PFN : 1); y++) {
                            const int i = 8 * e + 1 + t + 16 * y;
                            unsigned wa = VROW, wb = VROW;
                            if (plA && i <= nA) wa = proftab[pfA[y] & (GSUB ? 31 : 7)];
                            if (plB && i <= nB) wb = proftab[pfB[y] & (GSUB ? 31 : 7)];
                            ringA[i & (RING - 1)] = wa;
                            ringB[i & (RING - 1)] = wb;
                            const int i2 = i + CHUNK;
                            pfA[y] = (plA && i2 <= nA) ? rA[(int64_t)(i2 - 1) * stride] : 0;
                            pfB[y] = (plB && i2 <= nB) ? rB[(int64_t)(i2 - 1) * stride] : 0;
                        }
                    } else {
                        // four rows per lane and pair at a time: the eight letter loads of a batch are in flight
                        // together (one load, one lookup, one store at a time cost 16 exposed latencies per
                        // refill -- 11 % of the stall samples of config 5's wavefront kernel, ncu r02)
                        static_assert(CHUNK % 64 == 0, "a refill is a whole number of 4-row batches");
                        for (int x0 = t; x0 < CHUNK; x0 += 64) {
                            unsigned la[4], lb[4];
#pragma unroll
                            for (int y = 0; y < 4; y++) {
                                const int i = 8 * e + 1 + x0 + 16 * y;
                                la[y] = (plA && i <= nA) ? (unsigned)rA[(int64_t)(i - 1) * stride] : 0u;
                                lb[y] = (plB && i <= nB) ? (unsigned)rB[(int64_t)(i - 1) * stride] : 0u;
                            }
#pragma unroll
                            for (int y = 0; y < 4; y++) {
                                const int i = 8 * e + 1 + x0 + 16 * y;
                                unsigned wa = VROW, wb = VROW;
                                if (plA && i <= nA) wa = proftab[la[y] & (GSUB ? 31 : 7)];
                                if (plB && i <= nB) wb = proftab[lb[y] & (GSUB ? 31 : 7)];
                                ringA[i & (RING - 1)] = wa;
                                ringB[i & (RING - 1)] = wb;
                            }
                        }
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < BA16_PERIOD; j++) {
                    const int s = 8 * e + j;
                const int i = s + 1 - t;
                const unsigned profA = *(const unsigned *)((const char *)ringA + ro);
                const unsigned profB = *(const unsigned *)((const char *)ringB + ro);
                ro = (ro + 4u) & (RING * 4u - 1u);

                // boundary column of the left neighbour lane; lane 0: DP column 0 (all zero for SW)
                // or the staged last column of the previous pass
                unsigned rM = __shfl_up_sync(FULL, Mu[NC - 1], 1, 16);
                unsigned rL = __shfl_up_sync(FULL, outL, 1, 16);
                unsigned rH = __shfl_up_sync(FULL, Hu[NC - 1], 1, 16);
                if (t == 0) {
                    if (KIND == KIND_SW) {
                        rM = BIAS ? 0u : SYM ? goerp : 0u;
                        rL = rH = vbp;
                    } else {
                        // DP column 0: M = L = -inf; U = go + i*er (NW, nw_affine_letters.go:132-142)
                        // or 0 (Fitted, fitted_affine_letters.go:123-132); H = U
                        rM = SYM ? __vadd2(NEGP, goerp) : NEGP;
                        rL = NEGP;
                        rH = (KIND == KIND_NW) ? pack16(go + i * er, go + i * er) : 0u;
                    }
                    if (MULTI && pass > 0 && i >= 1 && i <= nbnd) {
                        const uint4 b = bbuf[grp][j];
                        rM = b.x;
                        rL = b.y;
                        rH = b.z;
                    }
                }

                // NW/Fitted: a lane whose first row has not come yet keeps its row-0 border state
                if (KIND == KIND_SW || i >= 1) {
                if (LIN) {
                    // one sweep along the row: X = max(diag + sub, up + er), T = max(left + eq, X [, 0])
                    unsigned Hd = sdiagH, Tl = rH;
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        unsigned sub;
                        if (!GSUB)
                            sub = ba_prmt(profA, profB, sel[k]);
                        else
                            sub = ba_prmt((unsigned)(int)*(const signed char *)(gtab + profA + sel[k]),
                                          (unsigned)(int)*(const signed char *)(gtab + profB + qb4[k]), 0x5410u);
                        const unsigned old = Hu[k];
                        const unsigned X = __viaddmax_s16x2(Hd, sub, __vadd2(old, erp));
                        Tl = (KIND == KIND_SW) ? __viaddmax_s16x2_relu(Tl, eqp, X) : __viaddmax_s16x2(Tl, eqp, X);
                        Hu[k] = Tl;
                        Hd = old;
                    }
                    sdiagH = rH;
                } else if (SYM) {
                    // U of this row from the previous row's (G, U)
#pragma unroll
                    for (int k = 0; k < NC; k++) Uu[k] = __viaddmax_s16x2(Uu[k], erp, Mu[k]);
                    unsigned Hd = sdiagH, Gl = rM, Ll = rL, Mprev = 0;
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        unsigned sub;
                        if (!GSUB)
                            sub = ba_prmt(profA, profB, sel[k]);
                        else
                            sub = ba_prmt((unsigned)(int)*(const signed char *)(gtab + profA + sel[k]),
                                          (unsigned)(int)*(const signed char *)(gtab + profB + qb4[k]), 0x5410u);
                        const unsigned M = BIAS                ? __viaddmax_s16x2(Hd, sub, vbp)   // clamp at the biased zero
                                           : (KIND == KIND_SW) ? __viaddmax_s16x2_relu(Hd, sub, erp)
                                                               : __vadd2(Hd, sub);
                        Hd = Hu[k];
                        Ll = __viaddmax_s16x2(Ll, eqp, Gl);
                        Hu[k] = __vimax3_s16x2(M, Uu[k], Ll);
                        Gl = BIAS ? M - vbp : __vadd2(M, goerp);   // BIAS: both halves >= vb, no borrow
                        Mu[k] = Gl;
                        if (KIND == KIND_SW) {
                            if (k & 1)
                                bmax = __vimax3_s16x2(bmax, Mprev, M);
                            else
                                Mprev = M;
                        }
                        if (k == COLS - 1) {
                            ckM0 = M;
                            ckL0 = Ll;
                        }
                        if (k == NC - 1) mlast = M;
                    }
                    sdiagH = rH;
                    outL = Ll;
                } else {
                // Three sweeps, each updating its array in place (no register rotation):
                //  U from the previous row's (M, U); M from the previous row's H; then L along the
                //  row and H = max3 for the next row's diagonal.
#pragma unroll
                for (int k = 0; k < NC; k++) Uu[k] = __viaddmax_s16x2(Mu[k], goerp, __vadd2(Uu[k], erp));
                {
                    unsigned Hd = sdiagH;
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        unsigned sub;
                        if (!GSUB)
                            sub = ba_prmt(profA, profB, sel[k]);
                        else
                            sub = ba_prmt((unsigned)(int)*(const signed char *)(gtab + profA + sel[k]),
                                          (unsigned)(int)*(const signed char *)(gtab + profB + qb4[k]), 0x5410u);
                        if (KIND == KIND_SW)
                            Mu[k] = __viaddmax_s16x2_relu(Hd, sub, erp);  // erp <= 0: same as max(.,0)
                        else
                            Mu[k] = __vadd2(Hd, sub);
                        Hd = Hu[k];
                    }
                }
                unsigned Ml = rM, Ll = rL;
#pragma unroll
                for (int k = 0; k < NC; k++) {
                    Ll = __viaddmax_s16x2(Ml, goeqp, __vadd2(Ll, eqp));
                    Ml = Mu[k];
                    Hu[k] = __vimax3_s16x2(Ml, Uu[k], Ll);
                    if (k == COLS - 1) {
                        // checkpoint words of my first strip, taken where (M, L) are live
                        ckM0 = Ml;
                        ckL0 = Ll;
                    }
                }
                sdiagH = rH;
                outL = Ll;
                }
                }
                if (KIND == KIND_SW && !SYM) {
#pragma unroll
                for (int k = 0; k + 1 < NC; k += 2)
                    bmax = LIN ? __vimax3_s16x2(bmax, Hu[k], Hu[k + 1]) : __vimax3_s16x2(bmax, Mu[k], Mu[k + 1]);
                }

                // column checkpoints of my two strips for this step, raw and packed as they are (SW's
                // deferred clamp of U and L is re-applied by the reader, ba_trace16 CkView)
                if (LIN) {
                    // element j = {T of strip 2t, T of strip 2t+1}: one 8-byte store per step
                    *(uint2 *)(colst + 2 * j) = make_uint2(Hu[COLS - 1], Hu[NC - 1]);
                } else {
                    // elements [strip][step] = {M, L}
                    *(uint2 *)(colst + 2 * j) = make_uint2(ckM0, ckL0);
                    *(uint2 *)(colst + 2 * (8 + j)) = make_uint2(SYM ? mlast : Mu[NC - 1], outL);
                }
                if (wr_bnd && t == 15 && i >= 1 && i <= nbnd) bnd[i] = ba16_bnd_pack(LIN ? 0u : Mu[NC - 1], outL, Hu[NC - 1], (unsigned)pass + 1u);
                }
                // ---- end of the period: row checkpoint, flush both arrays, best-cell bookkeeping
                // my row as it is: double-element c = columns 2c, 2c+1 -> {M, U, M, U} (affine; M is G for SYM) / {T, T}
#pragma unroll
                for (int c = 0; c < COLS; c++) {
                    if (LIN)
                        *(uint2 *)(rowst + 2 * c) = make_uint2(Hu[2 * c], Hu[2 * c + 1]);
                    else
                        *(uint4 *)(rowst + 4 * c) = make_uint4(Mu[2 * c], Uu[2 * c], Mu[2 * c + 1], Uu[2 * c + 1]);
                }
                __syncwarp();
                if (plC && (e << BA16_PSHIFT) < nstepsC) {
                    // column block: global [lane][element] (16 / 8 elements of 8 bytes per lane); lane t copies
                    // element t of lane row x (+ 8 lanes further: row x + 1 for the 8-element linear rows)
                    uint2 *g = (uint2 *)(colck + (int64_t)e * ba16_colblk(LIN));
                    const unsigned *sb = colbuf[grp];
                    if (LIN) {
#pragma unroll
                        for (int x = 0; x < 8; x++)
                            g[x * 16 + t] = *(const uint2 *)(sb + ((2 * x + (t >> 3)) * CST + (t & 7)) * 2);
                    } else {
#pragma unroll
                        for (int x = 0; x < 16; x++) g[x * 16 + t] = *(const uint2 *)(sb + (x * CST + t) * 2);
                    }
                }
                if (plC && 8 * e + 7 < nstepsC) {
                    // row block: global [lane][RS double-elements]
                    uint8_t *gb = rowck + (int64_t)e * ba16_rowblk(COLS, LIN);
                    const unsigned *sb = rowbuf[grp];
                    constexpr bool POW2 = (COLS & (COLS - 1)) == 0;
                    if (POW2 && COLS < 16) {
                        // 16 lanes cover 16 / COLS lane rows per copy: stage element (lane row * RST + c) with
                        // lane row = x * (16 / COLS) + t / COLS, c = t % COLS -- a per-lane constant plus an immediate
                        constexpr int LPR = 16 / COLS;
                        const int mine = ((t / COLS) * RST + (t % COLS)) * REW;
#pragma unroll
                        for (int x = 0; x < COLS; x++) {
                            if (LIN)
                                ((uint2 *)gb)[x * 16 + t] = *(const uint2 *)(sb + mine + x * LPR * RST * REW);
                            else
                                ((uint4 *)gb)[x * 16 + t] = *(const uint4 *)(sb + mine + x * LPR * RST * REW);
                        }
                    } else {
                        // RS == RST: the staged rows are the global rows, copy them linearly
                        static_assert(POW2 || RS == RST, "row stage and global row strides must agree");
#pragma unroll
                        for (int x = 0; x < RST; x++) {
                            if (LIN)
                                ((uint2 *)gb)[x * 16 + t] = *(const uint2 *)(sb + (x * 16 + t) * 2);
                            else
                                ((uint4 *)gb)[x * 16 + t] = *(const uint4 *)(sb + (x * 16 + t) * 4);
                        }
                    }
                }
                __syncwarp();
                if (MULTI) {
                    // lane 0 is done with this period's boundary rows: land the next ones
                    if (pre_on) BA16_BND_LAND(pre_row);
                    __syncwarp();
                }
                if (KIND == KIND_SW) {
                    // best tracking at period granularity ('>=' keeps the LAST period)
                    bool ph, pl;
                    vbest = __vibmax_s16x2(bmax, vbest, &ph, &pl);
                    if (pl) per_lo = e;
                    if (ph) per_hi = e;
                    bmax = 0;
                }
            }
#undef BA16_BND_LAND
            if (KIND == KIND_SW && plC) {
                // start-cell candidates of this (pass, lane): pair A's record, then pair B's
                int4 *cand = (int4 *)(P.arena + offC + (int64_t)np_ck * passb) + pass * 16 + t;
                *cand = make_int4(lo16(vbest) - vb, plA ? per_lo : -1, hi16(vbest) - vb, plB ? per_hi : -1);
            }
            __syncwarp();
        }
    }
}
