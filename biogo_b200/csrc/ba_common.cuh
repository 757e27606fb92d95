// ba_common.cuh -- shared definitions for the biogo align B200 kernels.
//
// Score domain.  Go ints are 64-bit; on the device every DP value is held in a 32-bit
// lane scaled by BA_SCALE (= 8) so that the three low bits are free for the tie-break
// "tags" described in DESIGN.md: a candidate c_k competing for a layer value X is turned
// into the unsigned key (c_k + k) - X, which equals k when c_k == X, is >= 8 when
// c_k > X and wraps to >= 2^31 when c_k < X.  An unsigned 3-way minimum over the keys
// therefore yields the FIRST matching candidate in the reference's `switch` order
// (align/sw_affine_letters.go:171-275, align/nw_affine_letters.go:198-299).
#pragma once

#ifdef BA_EMULATE
#include "cuda_emu.h"  // tests/emu: CPU SIMT emulator used by the no-GPU test tier only
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#define BA_SCALE 8
#define BA_SCALE_SHIFT 3
// "minus infinity" (align/align.go:54 minInt) in the scaled 32-bit domain.  The host
// guarantees every finite scaled value lies in (-2^26, 2^26), so NEG plus a handful of
// penalties never reaches a finite value and never wraps.
#define BA_NEG (-(1 << 29))
#define BA_MAX_LET 32          // largest supported matrix dimension (Protein is 26)
#define BA_WARP 32
#define BA_MAX_COLS 8          // columns per lane in the 32-bit systolic kernels
#define BA16_MAX_COLS 10       // columns per strip in the packed 16-bit kernels (10: the 257..320-column single-pass class)
#define BA_FILL_WARPS 4        // warps per CTA in the fill kernels

enum { KIND_SW = 0, KIND_NW = 1, KIND_FIT = 2 };

// Per-pair record produced by the validation kernel and consumed by fill/walk.
// status uses the BA_PAIR_* values of include/biogo_align.h.
struct PairStart {
    int32_t score;   // SW: best score (unscaled); others unused
    int32_t i, j;    // traceback start cell
    int32_t layer;   // NWAffine: start layer (0 M, 1 U, 2 L)
};

// One unit of fill work: a pair and where its traceback codes live.
struct WorkItem {
    int32_t pair;        // index into the batch
    int32_t cols;        // columns per lane used for this pair's code layout
    int64_t code_off;    // byte offset of this pair's codes in the arena
    // packed 16-bit path (ba_fill16): the two pairs of a COUPLE -- items 2k and 2k+1 of a launch -- share one
    // checkpoint region (their values sit in the low / high halves of the same words); the region's
    // geometry comes from the larger of the two
    int32_t n_ck;        // rows the region is laid out for: max(ref_len) over the couple
    int32_t ck;          // bits 0..15: passes of the region (max over the couple); bit 16: this pair is the high half
};

struct DeviceScoring {
    int32_t la[BA_MAX_LET * BA_MAX_LET];  // unscaled matrix, row-major, stride `let`
    int32_t let;
    int32_t gap_open;                     // unscaled
    int32_t alpha_len;
    int32_t kind, affine;
    uint8_t index[256];                   // alphabet.Index, 0xFF = illegal
};

// Code layout helpers (shared by fill and walk).  A pair with query length m and
// reference length n is processed in passes of W = 32*cols columns; in every pass lane t
// owns columns [t*cols+1, (t+1)*cols] of the pass and handles row i at step s = i-1+t.
// Codes are stored "skewed": for (pass, step, lane) one group of `cb` bytes, groups of a
// step contiguous over the lanes, so every warp-wide store is one coalesced segment.
__host__ __device__ inline int ba_code_group_bytes(int affine, int cols)
{
    if (affine) return cols <= 4 ? 4 : 8;      // 1 byte per cell
    return cols <= 4 ? 1 : 2;                  // 2 bits per cell
}
__host__ __device__ inline int64_t ba_code_bytes(int affine, int cols, int64_t n, int64_t m)
{
    if (n <= 0 || m <= 0) return 0;
    int64_t W = 32LL * cols;
    int64_t npass = (m + W - 1) / W;
    int64_t nsteps = n + 31;
    return npass * nsteps * 32 * ba_code_group_bytes(affine, cols);
}
