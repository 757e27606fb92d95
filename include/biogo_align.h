/*
 * biogo_align.h -- C ABI of libbiogoalign.so, the B200 (sm_100a) replacement for the
 * dynamic-programming hot path of biogo's `align` package.
 *
 * This is the drop-in boundary: plain pointers and sizes, no CUDA or torch types.
 * The twelve unexported Go methods
 *     (a SW|NW|Fitted|SWAffine|NWAffine|FittedAffine) alignLetters / alignQLetters
 *     (align/sw_letters.go:68, sw_qletters.go:68, nw_letters.go:75, nw_qletters.go:75,
 *      fitted_letters.go:70, fitted_qletters.go:70, sw_affine_letters.go:85,
 *      sw_affine_qletters.go:85, nw_affine_letters.go:98, nw_affine_qletters.go:98,
 *      fitted_affine_letters.go:93, fitted_affine_qletters.go:93)
 * each become one ba_align_batch() call with n == 1; the new AlignBatch entry point is
 * the same call with n > 1.  The cgo stub a maintainer would add is in INTEGRATION.md.
 *
 * There is no CPU fallback: every entry point that needs the GPU returns BA_ERR_CUDA
 * (with ba_last_error() text) when no sm_100 device/driver is usable.
 */
#ifndef BIOGO_ALIGN_H
#define BIOGO_ALIGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BA_ABI_VERSION 3 /* 3: ba_result.waves, overlap options, ba_create_multi; 2: wavefront/warp-trace counters, format, seqio, gather */

/* Aligner kind: align/sw.go:18, align/nw.go:16, align/fitted.go:16 (and *_affine.go:16). */
enum { BA_SW = 0, BA_NW = 1, BA_FITTED = 2 };

/* Call status (return value of the ba_* functions). */
enum {
    BA_OK = 0,
    BA_ERR_ARG = -1,          /* bad argument (null pointer, negative size, unknown kind) */
    BA_ERR_CUDA = -2,         /* CUDA runtime/driver failure; no CPU fallback exists */
    BA_ERR_MATRIX_SIZE = -3,  /* align.ErrMatrixWrongSize{Size,Len} (align/align.go:66-73); SW/NW kinds only */
    BA_ERR_RANGE = -4,        /* score bound does not fit the 32-bit device lanes */
    BA_ERR_NOMEM = -5,        /* host or device allocation failed */
    BA_ERR_NO_SCORING = -6,   /* ba_align_batch before ba_set_scoring */
    BA_ERR_FORMAT = -7        /* ba_fasta_parse / ba_fastq_parse: malformed text (ba_seqfile.err) */
};

/* Per-pair status (ba_result.status[k]). */
enum {
    BA_PAIR_OK = 0,
    BA_PAIR_ILLEGAL_REF = 1,  /* "align: illegal letter %q at position %d in rSeq"; err_pos[k] = position */
    BA_PAIR_ILLEGAL_QRY = 2,  /* "... in qSeq" */
    BA_PAIR_PANIC = 3,        /* the Go code panics on this input (index out of range: empty or
                                 illegal-letter input in the NW/Fitted border set-up) */
    BA_PAIR_NO_PATH = 4       /* "align: ... internal error: no path" (never expected) */
};

typedef struct ba_ctx ba_ctx;

/*
 * Result of one batch.  All arrays are owned by the library (pinned host memory) and
 * stay valid until ba_free_result().  Segment k (0 <= k < seg_count[p]) of pair p is
 *     segs[5*(seg_start[p]+k) .. +5] = { a.start, a.end, b.start, b.end, score }
 * in final order (already reversed like align/sw_letters.go:188-190), 0-based half-open
 * slice coordinates; a = reference, b = query.  This is featPair (align/align.go:121-124).
 * Pairs are laid out in the scheduler's work order, so seg_start is NOT monotone in p.
 */
typedef struct {
    int64_t n;           /* number of pairs */
    int32_t *status;     /* [n] BA_PAIR_* */
    int64_t *err_pos;    /* [n] letter position for BA_PAIR_ILLEGAL_*, else -1 */
    int64_t *seg_start;  /* [n] first segment of pair p inside segs (in segments) */
    int32_t *seg_count;  /* [n] number of segments of pair p (0 unless BA_PAIR_OK) */
    int32_t *segs;       /* [5*n_segs] */
    int64_t n_segs;
    /* device-side timing of this call, CUDA events on the launching stream (ms) */
    float ms_h2d, ms_kernels, ms_d2h;
    float ms_fill;        /* DP fill kernels only (the dominant kernel family) */
    float ms_trace;       /* traceback kernels (walk / tile trace + compaction) */
    float ms_host;        /* host wall time inside the call spent planning/scattering (GPU idle) */
    int64_t cells;        /* sum of len(ref)*len(query) over the batch */
    int64_t gpu_launches; /* kernels launched for this batch */
    int64_t fill_launches;/* of which DP fill kernels */
    int64_t code_bytes;   /* traceback bytes the fill kernels wrote to HBM */
    int64_t pairs_fast16; /* pairs served by the packed 16-bit DPX path (the rest: 32-bit exact-code path) */
    int64_t wavefront_launches; /* fill launches that spread the passes of long pairs over warps */
    int64_t warp_trace_launches; /* traceback launches that used one warp per pair */
    int64_t waves;        /* waves the batch was cut into (the traceback of one overlaps the fill of the next) */
} ba_result;

/* Context: one per (host thread, device).  Thread-compatible: calls on ONE ctx must be
 * serialised by the caller; different ctxs may be used concurrently. */
int ba_create(int device, ba_ctx **out);
void ba_destroy(ba_ctx *ctx);
const char *ba_last_error(const ba_ctx *ctx);

/*
 * Module (1): scoring upload.  Replaces the matrix check/flatten at the top of every
 * alignLetters (e.g. align/sw_letters.go:69-79) and alpha.LetterIndex()
 * (alphabet/alphabet.go:255).
 *   kind      BA_SW / BA_NW / BA_FITTED
 *   affine    0 = Linear [][]int, 1 = Affine{Matrix, GapOpen} (align/align.go:34-40)
 *   la        the square matrix flattened row-major, let*let Go ints
 *   alpha_len alphabet.Len(); let < alpha_len is BA_ERR_MATRIX_SIZE for SW and NW kinds
 *             (absent in Fitted/FittedAffine: align/fitted_letters.go:71-78)
 *   index     alphabet.Index, 256 entries, negative = letter not in the alphabet
 * The matrix is copied; callers may mutate theirs afterwards.
 *
 * LIMITS that the reference does not have (each is reported, never silently wrong):
 *   - let > 32                                  -> BA_ERR_RANGE ("scoring matrix larger than 32x32")
 *   - |entry| > 2^20 or |gap_open| > 2^20        -> BA_ERR_RANGE
 *   - ba_align_batch: |gap_open| + (len_r + len_q + 2) * max|entry| >= 2^23 for some pair
 *     (about 1.6 Mbp combined with NUC_4)       -> BA_ERR_RANGE; there is no 64-bit path
 *   - Fitted / FittedAffine with a matrix SMALLER than the alphabet (legal in the reference, which
 *     has no size check there) and letters beyond the matrix -> BA_PAIR_PANIC for that pair, where Go
 *     may silently read a neighbouring matrix row.
 * The host mirrors surface these as errors (Go: ErrDevice wrapping the text; Python: AlignError).
 */
int ba_set_scoring(ba_ctx *ctx, int kind, int affine, const int64_t *la, int let,
                   int alpha_len, int64_t gap_open, const int32_t *index);

/*
 * align/matrix as a library: the 78 scoring tables of align/matrix/matrices.go:14-2400
 * (NUC_4, NUC_4_4, DAYHOFF, GONNET, IDENTITY, MATCH, BLOSUM30..100, BLOSUMN, PAM10..500 and the *_cdi
 * variants) are compiled into the library, laid out for biogo's alphabets (gap letter at index 0, gap
 * row/column zero).  ba_matrix_get returns the dimension, the letters in index order and the row-major
 * table.  ba_set_scoring_named = ba_set_scoring on a named table with `gap` written into the gap letter's
 * row and column ([0][0] keeps the table's value), the usual way these tables are used with bíogo's
 * aligners (align/sw_example_test.go builds its matrix the same way by hand).
 * Every context keeps the device copies of its last eight scorings: switching back to one of them
 * (named or not) is a pointer swap, not an upload.
 */
int ba_matrix_count(void);
const char *ba_matrix_name(int k);
int ba_matrix_get(const char *name, int *let, const char **letters, const int16_t **table);
int ba_set_scoring_named(ba_ctx *ctx, int kind, int affine, const char *name, int64_t gap,
                         int alpha_len, int64_t gap_open, const int32_t *index);

/*
 * Modules (2)+(3): DP fill + traceback for n independent pairs.
 *   letters   one contiguous HOST buffer holding every sequence
 *   stride    1 = alphabet.Letters, 2 = alphabet.QLetters ({L,Q} pairs; Q is ignored,
 *             align/sw_qletters.go is sw_letters.go with x -> x.L)
 *   *_off     BYTE offsets of element 0 of each sequence inside `letters`
 *   *_len     lengths in letters
 * Inputs are read-only and not retained after return.
 */
int ba_align_batch(ba_ctx *ctx, const uint8_t *letters, int stride,
                   const int64_t *ref_off, const int32_t *ref_len,
                   const int64_t *qry_off, const int32_t *qry_len,
                   int64_t n, ba_result *out);

/*
 * Same, with the packed inputs already resident in DEVICE memory (all five pointers are
 * device pointers; letters_bytes = size of the letters buffer).  `cuda_stream` is a
 * cudaStream_t cast to void* (NULL = the library's own stream).  Used by bench.py to
 * time the kernels alone; results are still gathered to the host in `out`.
 */
int ba_align_batch_device(ba_ctx *ctx, const uint8_t *d_letters, int64_t letters_bytes, int stride,
                          const int64_t *d_ref_off, const int32_t *d_ref_len,
                          const int64_t *d_qry_off, const int32_t *d_qry_len,
                          const int32_t *h_ref_len, const int32_t *h_qry_len,
                          int64_t n, void *cuda_stream, ba_result *out);

void ba_free_result(ba_ctx *ctx, ba_result *res);

/* align.Format for a batch (reference: align/align.go:146-170): the two gapped rows of every
 * alignment.  `letters` and the offset/length arrays describe the packed batch exactly as for ba_align_batch;
 * pass letters == NULL to format the batch of this context's last ba_align_batch call, whose letters
 * are still on the device.  seg_start/seg_count/segs are a ba_result's arrays.  Rows hold letters
 * only (for QLetters the L byte; quality bytes are not carried).  Row p of side x is
 * row_x[off_x[p] .. off_x[p+1]). */
typedef struct ba_formatted {
    int64_t n;
    int64_t *off_a, *off_b; /* n + 1 entries each */
    uint8_t *row_a, *row_b;
    int64_t bytes_a, bytes_b;
    float ms_kernels;
} ba_formatted;
int ba_format_batch(ba_ctx *ctx, const uint8_t *letters, int stride, const int64_t *ref_off,
                    const int32_t *ref_len, const int64_t *qry_off, const int32_t *qry_len, int64_t n,
                    const int64_t *seg_start, const int32_t *seg_count, const int32_t *segs, int64_t n_segs,
                    uint8_t gap, ba_formatted *out);
void ba_free_formatted(ba_ctx *ctx, ba_formatted *f);

/* FASTA / FASTQ text -> packed batch (reference readers: io/seqio/fasta/fasta.go:60-113,
 * io/seqio/fastq/fastq.go:62-141).  Host-side: no context, no device.  `letters` is pinned memory
 * laid out like ba_align_batch's input (stride 1: Letters; stride 2: QLetters {L, Q}), so a pair
 * (record i, record j) is aligned by passing seq_off[i], seq_len[i], seq_off[j], seq_len[j].
 * Names and descriptions live in `text` (not NUL-terminated).  On BA_ERR_FORMAT `err` and
 * `err_line` describe the first bad line, with the reference readers' messages. */
typedef struct ba_seqfile {
    int64_t n;
    int stride;
    uint8_t *letters;
    int64_t letters_bytes;
    int64_t *seq_off;
    int32_t *seq_len;
    char *text;
    int64_t text_bytes;
    int64_t *name_off, *desc_off;
    int32_t *name_len, *desc_len;
    int64_t err_line;
    char err[160];
} ba_seqfile;
int ba_fasta_parse(const uint8_t *data, int64_t len, int stride, ba_seqfile *out);
/* quality_offset: 33 (Sanger, Illumina 1.8+) or 64 (Illumina 1.3/1.5); Q = byte - offset as in
 * alphabet/letters.go:66-72 */
int ba_fastq_parse(const uint8_t *data, int64_t len, int quality_offset, ba_seqfile *out);
void ba_seqfile_free(ba_seqfile *f);

/* One batch over several devices (pairs never communicate: no collective).  The host deals pair
 * ids to devices (largest cells first, so every device gets the same work within ~1 %), calls
 * this to pack a device's share contiguously -- dst needs stride * sum(ref_len + qry_len) bytes over
 * the share -- runs ba_align_batch per device from its own host thread and scatters the results
 * back by pair id (biogo_b200/shard.py is that host logic; go/align would do the same with one
 * goroutine per device). */
int ba_gather_pairs(const uint8_t *letters, int stride, const int64_t *ref_off, const int32_t *ref_len,
                    const int64_t *qry_off, const int32_t *qry_len, const int64_t *ids, int64_t k,
                    uint8_t *dst, int64_t dst_bytes, int64_t *new_ref_off, int64_t *new_qry_off);

/*
 * ONE batch over several GPUs of one box (north_star: "batches of independent pairs shard across the 8 GPUs
 * of one box by length-bucketed work lists, with a host-side gather and no NCCL").  A ba_multi owns one
 * ba_ctx and one host thread per device.  ba_align_batch_multi deals the pairs to the devices (cells bucketed
 * in quarter-octave classes, largest first, snake order: equal cells and equal length mix per device; batches
 * of equal-shaped pairs are cut into contiguous ranges), gathers each share's letters into that device's
 * pinned staging buffer piece by piece under the upload, runs the shares concurrently and gathers the results
 * by pair id into ONE ba_result with exactly the layout ba_align_batch produces.  The Go side's
 * AlignBatch calls this for large batches (go/align/cuda_bridge.go); there is no reference counterpart.
 *   devices / n_devices   CUDA device ordinals; n_devices <= 0 (or devices == NULL) = every visible device
 * Thread-compatible like ba_ctx: calls on one ba_multi must be serialised by the caller.
 * ba_multi_set_option forwards to every device context; its own key: "contiguous_shares" (default 1).
 */
typedef struct ba_multi ba_multi;
int ba_create_multi(const int *devices, int n_devices, ba_multi **out);
void ba_destroy_multi(ba_multi *m);
int ba_multi_devices(const ba_multi *m);
const char *ba_multi_last_error(const ba_multi *m);
int ba_multi_set_scoring(ba_multi *m, int kind, int affine, const int64_t *la, int let,
                         int alpha_len, int64_t gap_open, const int32_t *index);
int ba_multi_set_option(ba_multi *m, const char *key, int64_t value);
int ba_align_batch_multi(ba_multi *m, const uint8_t *letters, int stride,
                         const int64_t *ref_off, const int32_t *ref_len,
                         const int64_t *qry_off, const int32_t *qry_len,
                         int64_t n, ba_result *out);
void ba_multi_free_result(ba_multi *m, ba_result *res);

/*
 * align/pals/dp: PALS' banded x-drop dynamic programming over filter trapezoids,
 * batched.  One problem = dp.NewAligner(target, query, k, minLength, minId) + Aligner.Costs +
 * AlignTraps(trapezoids) (align/pals/dp/align.go:43-141; the DP itself: kernel.go:52-454).  A problem is
 * processed by one warp -- AlignTraps is sequential between the trapezoids of a problem (covered marks,
 * hit order) -- with the rows of every trace computed 32 columns at a time; many problems fill the GPU.
 *   letters            one host buffer holding every target and query (raw letters, compared byte for byte)
 *   target_off/len,    per problem
 *   query_off/len
 *   index              the TARGET alphabet's LetterIndex (256 entries, negative = not in the alphabet):
 *                      a query letter outside it never matches (kernel.go:227)
 *   trap_off           n + 1 offsets into traps (in trapezoids); traps: {Top, Bottom, Left, Right} each
 *                      (align/pals/filter/trapezoid.go:8-11)
 *   kmer               Aligner.k (trapezoids lower than k are skipped, align.go:79)
 * Hits of problem p are hits[hit_off[p] .. hit_off[p+1]), de-duplicated like align.go:88-138 (the order of
 * hits with EQUAL Abpos resp. Aepos inside that step follows a stable sort; Go's sort.Sort is not stable
 * and the reference pins no such case).  Integer arithmetic is 32-bit on the device:
 * (len(target) + len(query) + 8) * max(|MatchCost|, DiffCost) must stay below 2^29 (else BA_ERR_RANGE);
 * DiffCost >= 1.
 */
typedef struct {
    int32_t max_igap, diff_cost, same_cost, match_cost, block_cost;   /* dp.Costs, align/pals/dp/align.go:24-31 */
    double rmatch_cost;
} ba_pals_costs;
typedef struct {
    int64_t abpos, bbpos, aepos, bepos, low_diagonal, high_diagonal, score;   /* dp.Hit, align.go:143-150 */
    double error;
} ba_pals_hit;
typedef struct {
    int64_t n;
    int64_t *hit_off;     /* n + 1 */
    ba_pals_hit *hits;
    int64_t n_hits;
    int32_t *status;      /* per problem, 0 = ok */
    float ms_kernels;
} ba_pals_result;
int ba_pals_dp_batch(ba_ctx *ctx, const uint8_t *letters, const int64_t *target_off, const int32_t *target_len,
                     const int64_t *query_off, const int32_t *query_len, int64_t n, const int32_t *index,
                     const int64_t *trap_off, const int32_t *traps, int kmer, int min_hit_length, double min_id,
                     const ba_pals_costs *costs, ba_pals_result *out);
void ba_pals_free_result(ba_ctx *ctx, ba_pals_result *res);

/* Pinned host allocation so a caller (cgo, ctypes) can pack straight into DMA-able memory. */
void *ba_alloc_host(size_t bytes);
void ba_free_host(void *p);

/* Tunables (optional). key: "arena_bytes" (traceback-code arena per wave, default 48 GiB
 * or 60 % of free HBM), "force_path" (0 auto, 1 = 32-bit exact-code kernels only), "slot_cap"
 * (test hook: segment slots per pair on the 16-bit path, 0 = automatic), "long_mode" (wavefront
 * over the passes of kb-scale pairs: 0 never, 1 automatic, 2 whenever a launch has >= 4 passes),
 * "trace_mode" (16-bit path traceback: 0 thread per pair, 1 automatic, 2 warp per pair, 3 CTA per
 * pair),
 * "exclusive_compute" (default 1: ba_align_batch calls on the same device run their kernel phases
 * one at a time, so that with several contexts on several threads the upload of one batch overlaps
 * the kernels of another), "plan_cache" (default 1: a batch whose length arrays equal the previous
 * batch's reuses its work lists). */
int ba_set_option(ba_ctx *ctx, const char *key, int64_t value);

/* Build/ABI information. */
int ba_abi_version(void);
const char *ba_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* BIOGO_ALIGN_H */
