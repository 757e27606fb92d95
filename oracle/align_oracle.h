/*
 * align_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, int64 like Go's int on amd64) of the six pairwise
 * aligners of biogo's `align` package.  It is the parity checker for the CUDA
 * path and the "port" CPU baseline of bench.py; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use it.
 * The product library (libbiogoalign.so) never links or calls this file.
 *
 * Parity pin: checked against every golden vector the reference holds for the
 * path (G1..G8 of SURVEY.md section 4: the seven Example* outputs in
 * align/sw_example_test.go, align/nw_example_test.go,
 * align/fitted_example_test.go and TestIssue58 in align/align_test.go:28-51)
 * by tests/test_oracle_golden.py.  The Go reference itself cannot run in this
 * image (no Go toolchain), so oracle/_ref does not exist.
 */
#ifndef BIOGO_ALIGN_ORACLE_H
#define BIOGO_ALIGN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* aligner kinds: align/sw.go:18, align/nw.go:16, align/fitted.go:16 (+ *_affine.go) */
enum { ORC_SW = 0, ORC_NW = 1, ORC_FITTED = 2 };

/* per-pair outcome */
enum {
    ORC_OK = 0,
    ORC_ILLEGAL_REF = 1,   /* fmt.Errorf("align: illegal letter %q at position %d in rSeq") */
    ORC_ILLEGAL_QRY = 2,   /* ... in qSeq */
    ORC_PANIC = 3,         /* the Go code would panic (index out of range on empty/illegal input) */
    ORC_NO_PATH = 4        /* panic("align: ... internal error: no path ...") */
};

typedef struct {
    int kind;            /* ORC_SW / ORC_NW / ORC_FITTED */
    int affine;          /* 0: Linear [][]int, 1: Affine{Matrix, GapOpen} (align/align.go:34-40) */
    int let;             /* len(matrix) */
    const int64_t *la;   /* matrix flattened row-major, let*let entries */
    int64_t gap_open;    /* Affine.GapOpen */
    const int32_t *index;/* alphabet.Index: 256 entries, -1 = not in alphabet (alphabet/alphabet.go:113,183-219) */
} orc_scoring;

typedef struct {
    int64_t a_start, a_end, b_start, b_end, score; /* featPair, align/align.go:121-124 */
} orc_seg;

/*
 * Align one pair.  `stride` is 1 for alphabet.Letters and 2 for
 * alphabet.QLetters ({L,Q} byte pairs, alphabet/letters.go:182-185; Q is never
 * read, as in the generated *_qletters.go files).
 * On ORC_OK *segs (malloc'd, caller frees) holds *nsegs segments in final
 * (reversed) order.  err_pos receives the letter position for ORC_ILLEGAL_*.
 */
int orc_align(const orc_scoring *sc,
              const uint8_t *ref, int64_t nref,
              const uint8_t *qry, int64_t nqry, int stride,
              orc_seg **segs, int64_t *nsegs, int64_t *err_pos);

/*
 * Batch driver used for the CPU baseline: pairs are pulled from a shared
 * counter by `nthreads` pthreads (the shape a Go user gets from
 * concurrent.Processor, concurrent/processor.go:31-77).  Outputs:
 *   status[n], err_pos[n], seg_off[n+1] and a malloc'd segment array
 *   (int64 x5 per segment) returned through *segs_out.
 */
int orc_align_batch(const orc_scoring *sc, const uint8_t *letters, int stride,
                    const int64_t *ref_off, const int32_t *ref_len,
                    const int64_t *qry_off, const int32_t *qry_len,
                    int64_t n, int nthreads,
                    int32_t *status, int64_t *err_pos, int64_t *seg_off,
                    int64_t **segs_out);

void orc_free(void *p);

/* alphabet.newAlphabet index construction (alphabet/alphabet.go:183-219) */
void orc_build_index(const char *letters, int case_sensitive, int32_t index[256]);

/* featPair.String (align/align.go:129-144) for a whole alignment, fmt.Sprint style: "[a b c]" */
int orc_format_pairs(const orc_seg *segs, int64_t n, char *buf, int64_t buflen);

/* align.Format (align/align.go:148-170) for Letters; returns lengths through out_len */
int orc_format_alignment(const uint8_t *ref, const uint8_t *qry, int stride,
                         const orc_seg *segs, int64_t n, uint8_t gap,
                         uint8_t *out_a, uint8_t *out_b, int64_t cap, int64_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
