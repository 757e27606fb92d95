// ba_format.cuh -- align.Format for a whole batch (align/align.go:146-170), SURVEY section 8(f) row 1.
//
// For every feature pair of an alignment, each of the two rows receives either its own letters
// [start, end) or, when its feature is empty, the gap letter repeated len(other feature) times.
// Kernel 1 sizes the two rows of every pair, the host scheduler scans the sizes, kernel 2 writes
// both rows with one warp per pair (coalesced byte copies).  HBM-bound: 2 bytes read + 2 bytes
// written per alignment column at most.
#pragma once

#include "ba_common.cuh"

struct FormatParams {
    const uint8_t *letters;     // the packed batch (raw letters, not indices)
    int stride;                 // 1 = Letters, 2 = QLetters (the L byte of every {L, Q})
    const int64_t *ref_off, *qry_off;
    const int64_t *seg_start;   // per pair: first segment
    const int32_t *seg_count;   // per pair
    const int32_t *segs;        // 5 ints per segment, final order
    int64_t n;
    int32_t *len_a, *len_b;     // per pair row lengths (kernel 1 out)
    const int64_t *off_a, *off_b;  // exclusive scans of the lengths (kernel 2 in)
    uint8_t *row_a, *row_b;
    uint8_t gap;
};

__global__ void ba_format_size_kernel(FormatParams P)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const int32_t *s = P.segs + 5 * P.seg_start[p];
    int la = 0, lb = 0;
    for (int k = 0; k < P.seg_count[p]; k++, s += 5) {
        const int alen = s[1] - s[0], blen = s[3] - s[2];
        la += alen ? alen : blen;   // fc[i].Len() == 0 -> gap.Repeat(fc[1-i].Len())
        lb += blen ? blen : alen;
    }
    P.len_a[p] = la;
    P.len_b[p] = lb;
}

__global__ void __launch_bounds__(256) ba_format_write_kernel(FormatParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t p = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); p < P.n; p += warps) {
        const int32_t *s = P.segs + 5 * P.seg_start[p];
        const uint8_t *ra = P.letters + P.ref_off[p], *rb = P.letters + P.qry_off[p];
        uint8_t *da = P.row_a + P.off_a[p], *db = P.row_b + P.off_b[p];
        const int cnt = P.seg_count[p];
        for (int k = 0; k < cnt; k++, s += 5) {
            const int a0 = s[0], alen = s[1] - s[0], b0 = s[2], blen = s[3] - s[2];
            const int wa = alen ? alen : blen, wb = blen ? blen : alen;
            for (int x = lane; x < wa; x += 32) da[x] = alen ? ra[(int64_t)(a0 + x) * P.stride] : P.gap;
            for (int x = lane; x < wb; x += 32) db[x] = blen ? rb[(int64_t)(b0 + x) * P.stride] : P.gap;
            da += wa;
            db += wb;
        }
    }
}
