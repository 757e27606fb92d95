// cuda_bridge.go -- the cgo layer a maintainer adds to package align to make the B200 library a
// drop-in for the DP hot path.  NOT compiled in this repository (the image has no Go toolchain);
// it is kept mechanical and small on purpose.  See INTEGRATION.md.
//
// It replaces the bodies of the twelve generated methods
//   (a SW|NW|Fitted|SWAffine|NWAffine|FittedAffine) alignLetters / alignQLetters
// (align/sw_letters.go:68 ... align/fitted_affine_qletters.go:93) with one call into
// libbiogoalign.so, and adds AlignBatch.  align.go, the six wrapper files (sw.go ...) and
// featPair stay as they are, so every exported name, error identity and String() is unchanged.

package align

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../biogo_b200/lib -lbiogoalign -Wl,-rpath,${SRCDIR}/../../biogo_b200/lib
#include <stdlib.h>
#include "biogo_align.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	"github.com/biogo/biogo/alphabet"
	"github.com/biogo/biogo/feat"
	"github.com/biogo/biogo/seq"
)

const (
	kindSW = iota
	kindNW
	kindFitted
)

// One ba_ctx per OS thread that ever calls Align: Align stays reentrant like the reference
// (concurrent/processor.go:48-74 calls it from many goroutines).
var ctxPool = sync.Pool{New: func() interface{} {
	var c *C.ba_ctx
	if rc := C.ba_create(0, &c); rc != C.BA_OK {
		panic(fmt.Sprintf("align: CUDA context: %s", C.GoString(C.ba_last_error(c))))
	}
	return c
}}

// alignPacked runs one batch.  letters is a single contiguous buffer (no Go pointers to Go
// pointers cross the boundary); offsets are byte offsets, stride is 1 (Letters) or 2 (QLetters).
func alignPacked(kind int, affine bool, matrix Linear, gapOpen int, alpha alphabet.Alphabet,
	letters []byte, stride int, refOff []int64, refLen []int32, qryOff []int64, qryLen []int32,
	checkSize bool) ([][]feat.Pair, []error) {

	n := len(refOff)
	alns, errs := make([][]feat.Pair, n), make([]error, n)
	fail := func(err error) ([][]feat.Pair, []error) {
		for i := range errs {
			errs[i] = err
		}
		return alns, errs
	}

	// the checks at the top of every alignLetters (align/sw_letters.go:69-79)
	let := len(matrix)
	if checkSize && let < alpha.Len() {
		return fail(ErrMatrixWrongSize{Size: let, Len: alpha.Len()})
	}
	la := make([]C.int64_t, 0, let*let)
	for _, row := range matrix {
		if len(row) != let {
			return fail(ErrMatrixNotSquare)
		}
		for _, v := range row {
			la = append(la, C.int64_t(v))
		}
	}
	var index [256]C.int32_t
	for i, v := range alpha.LetterIndex() {
		index[i] = C.int32_t(v)
	}

	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	ctx := ctxPool.Get().(*C.ba_ctx)
	defer ctxPool.Put(ctx)

	aff := C.int(0)
	if affine {
		aff = 1
	}
	if rc := C.ba_set_scoring(ctx, C.int(kind), aff, &la[0], C.int(let), C.int(alpha.Len()),
		C.int64_t(gapOpen), &index[0]); rc != C.BA_OK {
		panic(fmt.Sprintf("align: %s", C.GoString(C.ba_last_error(ctx))))
	}
	var res C.ba_result
	var lp *C.uint8_t
	if len(letters) > 0 {
		lp = (*C.uint8_t)(unsafe.Pointer(&letters[0]))
	}
	rc := C.ba_align_batch(ctx, lp, C.int(stride),
		(*C.int64_t)(unsafe.Pointer(&refOff[0])), (*C.int32_t)(unsafe.Pointer(&refLen[0])),
		(*C.int64_t)(unsafe.Pointer(&qryOff[0])), (*C.int32_t)(unsafe.Pointer(&qryLen[0])),
		C.int64_t(n), &res)
	if rc != C.BA_OK {
		panic(fmt.Sprintf("align: %s", C.GoString(C.ba_last_error(ctx)))) // no CPU fallback
	}
	defer C.ba_free_result(ctx, &res)

	status := unsafe.Slice((*int32)(unsafe.Pointer(res.status)), n)
	errPos := unsafe.Slice((*int64)(unsafe.Pointer(res.err_pos)), n)
	segStart := unsafe.Slice((*int64)(unsafe.Pointer(res.seg_start)), n)
	segCount := unsafe.Slice((*int32)(unsafe.Pointer(res.seg_count)), n)
	segs := unsafe.Slice((*int32)(unsafe.Pointer(res.segs)), 5*int(res.n_segs))
	for p := 0; p < n; p++ {
		switch status[p] {
		case C.BA_PAIR_OK:
			aln := make([]feat.Pair, segCount[p])
			for k := range aln {
				s := segs[5*(int(segStart[p])+k):]
				aln[k] = &featPair{
					a:     feature{start: int(s[0]), end: int(s[1])},
					b:     feature{start: int(s[2]), end: int(s[3])},
					score: int(s[4]),
				}
			}
			alns[p] = aln
		case C.BA_PAIR_ILLEGAL_REF:
			pos := int(errPos[p])
			errs[p] = fmt.Errorf("align: illegal letter %q at position %d in rSeq",
				letters[int(refOff[p])+pos*stride], pos)
		case C.BA_PAIR_ILLEGAL_QRY:
			pos := int(errPos[p])
			errs[p] = fmt.Errorf("align: illegal letter %q at position %d in qSeq",
				letters[int(qryOff[p])+pos*stride], pos)
		case C.BA_PAIR_PANIC:
			panic("runtime error: index out of range") // the reference panics on this input
		default:
			panic("align: internal error: no path")
		}
	}
	return alns, errs
}

// one pair: Letters
func alignOneLetters(kind int, affine bool, m Linear, gapOpen int, checkSize bool,
	rSeq, qSeq alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	buf := make([]byte, 0, len(rSeq)+len(qSeq))
	buf = append(buf, alphabet.LettersToBytes(rSeq)...)
	buf = append(buf, alphabet.LettersToBytes(qSeq)...)
	a, e := alignPacked(kind, affine, m, gapOpen, alpha, buf, 1,
		[]int64{0}, []int32{int32(len(rSeq))}, []int64{int64(len(rSeq))}, []int32{int32(len(qSeq))}, checkSize)
	return a[0], e[0]
}

// one pair: QLetters ({L,Q} byte pairs; Q is ignored exactly like align/sw_qletters.go)
func alignOneQLetters(kind int, affine bool, m Linear, gapOpen int, checkSize bool,
	rSeq, qSeq alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	buf := make([]byte, 0, 2*(len(rSeq)+len(qSeq)))
	for _, l := range rSeq {
		buf = append(buf, byte(l.L), byte(l.Q))
	}
	for _, l := range qSeq {
		buf = append(buf, byte(l.L), byte(l.Q))
	}
	a, e := alignPacked(kind, affine, m, gapOpen, alpha, buf, 2,
		[]int64{0}, []int32{int32(len(rSeq))}, []int64{int64(2 * len(rSeq))}, []int32{int32(len(qSeq))}, checkSize)
	return a[0], e[0]
}

// The twelve replaced methods (only the SW family is spelled out; NW / Fitted differ in the
// kind constant, Fitted/FittedAffine pass checkSize=false: align/fitted_letters.go:71-78).
func (a SW) alignLetters(r, q alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return alignOneLetters(kindSW, false, Linear(a), 0, true, r, q, alpha)
}
func (a SW) alignQLetters(r, q alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return alignOneQLetters(kindSW, false, Linear(a), 0, true, r, q, alpha)
}
func (a SWAffine) alignLetters(r, q alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return alignOneLetters(kindSW, true, a.Matrix, a.GapOpen, true, r, q, alpha)
}
func (a SWAffine) alignQLetters(r, q alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return alignOneQLetters(kindSW, true, a.Matrix, a.GapOpen, true, r, q, alpha)
}

// AlignBatch is the new entry point: many independent pairs, one device call.  Pairs that fail
// the wrapper validation (align/sw.go:23-48) get their error without touching the GPU.
func (a SWAffine) AlignBatch(refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	return alignBatch(kindSW, true, a.Matrix, a.GapOpen, true, refs, queries)
}

func alignBatch(kind int, affine bool, m Linear, gapOpen int, checkSize bool,
	refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	n := len(refs)
	alns, errs := make([][]feat.Pair, n), make([]error, n)
	var (
		buf            []byte
		refOff, qryOff []int64
		refLen, qryLen []int32
		live           []int
		alpha0         alphabet.Alphabet
		stride         = 1
	)
	for k := 0; k < n; k++ {
		alpha := refs[k].Alphabet()
		switch {
		case alpha == nil:
			errs[k] = ErrNoAlphabet
			continue
		case alpha != queries[k].Alphabet():
			errs[k] = ErrMismatchedAlphabets
			continue
		case alpha.IndexOf(alpha.Gap()) != 0:
			errs[k] = ErrNotGappedAlphabet
			continue
		}
		if alpha0 == nil {
			alpha0 = alpha
		}
		switch r := refs[k].Slice().(type) {
		case alphabet.Letters:
			q, ok := queries[k].Slice().(alphabet.Letters)
			if !ok {
				errs[k] = ErrMismatchedTypes
				continue
			}
			refOff, refLen = append(refOff, int64(len(buf))), append(refLen, int32(len(r)))
			buf = append(buf, alphabet.LettersToBytes(r)...)
			qryOff, qryLen = append(qryOff, int64(len(buf))), append(qryLen, int32(len(q)))
			buf = append(buf, alphabet.LettersToBytes(q)...)
		case alphabet.QLetters:
			q, ok := queries[k].Slice().(alphabet.QLetters)
			if !ok {
				errs[k] = ErrMismatchedTypes
				continue
			}
			stride = 2
			refOff, refLen = append(refOff, int64(len(buf))), append(refLen, int32(len(r)))
			for _, l := range r {
				buf = append(buf, byte(l.L), byte(l.Q))
			}
			qryOff, qryLen = append(qryOff, int64(len(buf))), append(qryLen, int32(len(q)))
			for _, l := range q {
				buf = append(buf, byte(l.L), byte(l.Q))
			}
		default:
			errs[k] = ErrTypeNotHandled
			continue
		}
		live = append(live, k)
	}
	if len(live) == 0 {
		return alns, errs
	}
	a, e := alignPacked(kind, affine, m, gapOpen, alpha0, buf, stride, refOff, refLen, qryOff, qryLen, checkSize)
	for p, k := range live {
		alns[k], errs[k] = a[p], e[p]
	}
	return alns, errs
}

// FormatBatch is align.Format (align/align.go:146-170) for many alignments in one device call
// (ba_format_batch).  The rows come back as letters; for QLetters the qualities are re-attached
// here along the same segments (the gap QLetter has Q = 0, align/align.go:161).
func FormatBatch(as, bs []seq.Slicer, alns [][]feat.Pair, gap alphabet.Letter) [][2]alphabet.Slice {
	n := len(as)
	out := make([][2]alphabet.Slice, n)
	var (
		buf              []byte
		refOff, qryOff   []int64
		refLen, qryLen   []int32
		segStart         = make([]int64, n)
		segCount         = make([]int32, n)
		segs             []int32
		stride           = 1
	)
	pack := func(s alphabet.Slice) (off int64, ln int32) {
		off = int64(len(buf))
		switch v := s.(type) {
		case alphabet.Letters:
			buf = append(buf, alphabet.LettersToBytes(v)...)
			ln = int32(len(v))
		case alphabet.QLetters:
			stride = 2
			for _, l := range v {
				buf = append(buf, byte(l.L), byte(l.Q))
			}
			ln = int32(len(v))
		}
		return
	}
	for k := 0; k < n; k++ {
		o, l := pack(as[k].Slice())
		refOff, refLen = append(refOff, o), append(refLen, l)
		o, l = pack(bs[k].Slice())
		qryOff, qryLen = append(qryOff, o), append(qryLen, l)
		segStart[k], segCount[k] = int64(len(segs)/5), int32(len(alns[k]))
		for _, fp := range alns[k] {
			f := fp.Features()
			segs = append(segs, int32(f[0].Start()), int32(f[0].End()), int32(f[1].Start()), int32(f[1].End()),
				int32(fp.Score()))
		}
	}
	if n == 0 {
		return out
	}
	ctx := ctxPool.Get().(*C.ba_ctx)
	defer ctxPool.Put(ctx)
	var res C.ba_formatted
	var segp *C.int32_t
	if len(segs) > 0 {
		segp = (*C.int32_t)(unsafe.Pointer(&segs[0]))
	}
	rc := C.ba_format_batch(ctx, (*C.uint8_t)(unsafe.Pointer(&buf[0])), C.int(stride),
		(*C.int64_t)(unsafe.Pointer(&refOff[0])), (*C.int32_t)(unsafe.Pointer(&refLen[0])),
		(*C.int64_t)(unsafe.Pointer(&qryOff[0])), (*C.int32_t)(unsafe.Pointer(&qryLen[0])), C.int64_t(n),
		(*C.int64_t)(unsafe.Pointer(&segStart[0])), (*C.int32_t)(unsafe.Pointer(&segCount[0])), segp,
		C.int64_t(len(segs)/5), C.uint8_t(gap), &res)
	if rc != C.BA_OK {
		panic("align: " + C.GoString(C.ba_last_error(ctx))) // no CPU fallback
	}
	defer C.ba_free_formatted(ctx, &res)
	offA := unsafe.Slice((*int64)(unsafe.Pointer(res.off_a)), n+1)
	offB := unsafe.Slice((*int64)(unsafe.Pointer(res.off_b)), n+1)
	rowA := unsafe.Slice((*byte)(unsafe.Pointer(res.row_a)), int(res.bytes_a))
	rowB := unsafe.Slice((*byte)(unsafe.Pointer(res.row_b)), int(res.bytes_b))
	for k := 0; k < n; k++ {
		rows := [2][]byte{rowA[offA[k]:offA[k+1]], rowB[offB[k]:offB[k+1]]}
		for side, src := range [2]alphabet.Slice{as[k].Slice(), bs[k].Slice()} {
			switch v := src.(type) {
			case alphabet.Letters:
				out[k][side] = alphabet.BytesToLetters(append([]byte(nil), rows[side]...))
			case alphabet.QLetters:
				ql := make(alphabet.QLetters, 0, len(rows[side]))
				x := 0
				for _, fp := range alns[k] {
					f := fp.Features()
					if f[side].Len() == 0 {
						for i := 0; i < f[1-side].Len(); i, x = i+1, x+1 {
							ql = append(ql, alphabet.QLetter{L: alphabet.Letter(rows[side][x])})
						}
					} else {
						for i := f[side].Start(); i < f[side].End(); i, x = i+1, x+1 {
							ql = append(ql, alphabet.QLetter{L: alphabet.Letter(rows[side][x]), Q: v[i].Q})
						}
					}
				}
				out[k][side] = ql
			}
		}
	}
	return out
}
