// cuda_bridge.go -- the cgo layer a maintainer adds to package align to make the B200 library a
// drop-in for the DP hot path.  See INTEGRATION.md.
//
// STATUS: written against include/biogo_align.h but NEVER COMPILED OR RUN -- this repository's
// image has no Go toolchain.  tests/test_go_overlay_static.py checks, by parsing this file, that
// every method the reference declares exists with the reference's signature and that every C
// symbol used here is declared in the header; that is all the verification it has had.
//
// It replaces the bodies of the twelve generated methods
//   (a SW|NW|Fitted|SWAffine|NWAffine|FittedAffine) alignLetters / alignQLetters
// (align/sw_letters.go:68, sw_qletters.go:68, nw_letters.go:75, nw_qletters.go:75,
// fitted_letters.go:70, fitted_qletters.go:70, sw_affine_letters.go:85, sw_affine_qletters.go:85,
// nw_affine_letters.go:98, nw_affine_qletters.go:98, fitted_affine_letters.go:93,
// fitted_affine_qletters.go:93) with one call into libbiogoalign.so, and adds AlignBatch on all six
// types.  align.go, the six wrapper files (sw.go ...) and featPair stay as they are, so every
// exported name, error identity and String() is unchanged.

package align

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../biogo_b200/lib -lbiogoalign -Wl,-rpath,${SRCDIR}/../../biogo_b200/lib
#include <stdlib.h>
#include "biogo_align.h"
*/
import "C"

import (
	"errors"
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

// ErrDevice reports a failure of the CUDA library (no usable B200, out of device memory, a score
// range the device lanes cannot hold).  There is no CPU fallback: Align returns it, it never
// silently computes elsewhere.
var ErrDevice = errors.New("align: CUDA device error")

// MultiDeviceMinPairs is the batch size from which AlignBatch spreads one batch over every GPU of
// the box (ba_align_batch_multi); smaller batches stay on one device.
var MultiDeviceMinPairs = 20000

// ---- contexts -----------------------------------------------------------------------------
//
// A devCtx is one ba_ctx plus a pinned packing buffer that grows as needed.  Contexts live in an
// explicit bounded free list -- not a sync.Pool, whose entries the GC may drop without running
// ba_destroy (leaking streams, device arenas and page-locked memory).  A context that does not
// fit back into the list is destroyed on the spot.  Shutdown destroys the idle ones.

type devCtx struct {
	c      *C.ba_ctx
	pinned unsafe.Pointer // ba_alloc_host block the batch is packed into
	cap    int
}

var (
	ctxFree  = make(chan *devCtx, 16)
	multiMu  sync.Mutex  // one multi-device batch at a time: it uses every GPU anyway
	multiCtx *C.ba_multi // created on first use
	multiBuf devCtx      // pinned packing buffer of the multi-device path (c unused)
)

func getCtx() (*devCtx, error) {
	select {
	case d := <-ctxFree:
		return d, nil
	default:
	}
	var c *C.ba_ctx
	if rc := C.ba_create(0, &c); rc != C.BA_OK {
		msg := "ba_create failed"
		if c != nil {
			msg = C.GoString(C.ba_last_error(c))
			C.ba_destroy(c)
		}
		return nil, fmt.Errorf("%w: %s", ErrDevice, msg)
	}
	return &devCtx{c: c}, nil
}

func putCtx(d *devCtx) {
	select {
	case ctxFree <- d:
	default:
		d.destroy()
	}
}

func (d *devCtx) destroy() {
	if d.pinned != nil {
		C.ba_free_host(d.pinned)
		d.pinned = nil
	}
	if d.c != nil {
		C.ba_destroy(d.c)
		d.c = nil
	}
}

// buffer returns the context's pinned packing buffer with room for n bytes.
func (d *devCtx) buffer(n int) ([]byte, error) {
	if n == 0 {
		return nil, nil
	}
	if n > d.cap {
		if d.pinned != nil {
			C.ba_free_host(d.pinned)
			d.pinned, d.cap = nil, 0
		}
		want := n + n/4
		p := C.ba_alloc_host(C.size_t(want))
		if p == nil {
			return nil, fmt.Errorf("%w: pinned host allocation of %d bytes failed", ErrDevice, want)
		}
		d.pinned, d.cap = p, want
	}
	return unsafe.Slice((*byte)(d.pinned), n), nil
}

// Shutdown releases every idle device context (streams, device arenas, pinned memory).  Contexts
// in use by concurrent Align calls are released when those calls return them to a full list, or by
// a later Shutdown.
func Shutdown() {
	for {
		select {
		case d := <-ctxFree:
			d.destroy()
			continue
		default:
		}
		break
	}
	multiMu.Lock()
	if multiCtx != nil {
		C.ba_destroy_multi(multiCtx)
		multiCtx = nil
	}
	multiBuf.destroy()
	multiMu.Unlock()
}

// ---- one packed batch -----------------------------------------------------------------------

type scoring struct {
	kind      int
	affine    bool
	matrix    Linear
	gapOpen   int
	checkSize bool // false for Fitted / FittedAffine: align/fitted_letters.go:71-78 has no size check
}

// flatten performs the checks at the top of every alignLetters (align/sw_letters.go:69-79) in Go,
// so ErrMatrixWrongSize and ErrMatrixNotSquare keep their identity.
func (s scoring) flatten(alpha alphabet.Alphabet) ([]C.int64_t, error) {
	let := len(s.matrix)
	if s.checkSize && let < alpha.Len() {
		return nil, ErrMatrixWrongSize{Size: let, Len: alpha.Len()}
	}
	la := make([]C.int64_t, 0, let*let)
	for _, row := range s.matrix {
		if len(row) != let {
			return nil, ErrMatrixNotSquare
		}
		for _, v := range row {
			la = append(la, C.int64_t(v))
		}
	}
	return la, nil
}

// packed is one sub-batch: pairs that share an alphabet and a slice type, packed into one
// contiguous pinned buffer (no Go pointer to a Go pointer crosses the boundary).
type packed struct {
	alpha          alphabet.Alphabet
	stride         int // 1 alphabet.Letters, 2 alphabet.QLetters ({L, Q}; Q is never read)
	letters        []byte
	refOff, qryOff []int64
	refLen, qryLen []int32
}

// run aligns one sub-batch.  A device failure is returned for every pair; reference panics stay
// panics (SURVEY 8b).
func (s scoring) run(pk *packed, d *devCtx, multi bool) ([][]feat.Pair, []error) {
	n := len(pk.refOff)
	alns, errs := make([][]feat.Pair, n), make([]error, n)
	fail := func(err error) ([][]feat.Pair, []error) {
		for i := range errs {
			errs[i] = err
		}
		return alns, errs
	}
	if n == 0 {
		return alns, errs
	}
	la, err := s.flatten(pk.alpha)
	if err != nil {
		return fail(err)
	}
	if len(la) == 0 {
		return fail(ErrMatrixWrongSize{Size: 0, Len: pk.alpha.Len()})
	}
	var index [256]C.int32_t
	for i, v := range pk.alpha.LetterIndex() {
		index[i] = C.int32_t(v)
	}
	aff := C.int(0)
	if s.affine {
		aff = 1
	}
	var lp *C.uint8_t
	if len(pk.letters) > 0 {
		lp = (*C.uint8_t)(unsafe.Pointer(&pk.letters[0]))
	}
	ro, rl := (*C.int64_t)(unsafe.Pointer(&pk.refOff[0])), (*C.int32_t)(unsafe.Pointer(&pk.refLen[0]))
	qo, ql := (*C.int64_t)(unsafe.Pointer(&pk.qryOff[0])), (*C.int32_t)(unsafe.Pointer(&pk.qryLen[0]))

	var res C.ba_result
	if multi {
		if rc := C.ba_multi_set_scoring(multiCtx, C.int(s.kind), aff, &la[0], C.int(len(s.matrix)),
			C.int(pk.alpha.Len()), C.int64_t(s.gapOpen), &index[0]); rc != C.BA_OK {
			return fail(fmt.Errorf("%w: %s", ErrDevice, C.GoString(C.ba_multi_last_error(multiCtx))))
		}
		if rc := C.ba_align_batch_multi(multiCtx, lp, C.int(pk.stride), ro, rl, qo, ql, C.int64_t(n), &res); rc != C.BA_OK {
			return fail(fmt.Errorf("%w: %s", ErrDevice, C.GoString(C.ba_multi_last_error(multiCtx))))
		}
		defer C.ba_multi_free_result(multiCtx, &res)
	} else {
		if rc := C.ba_set_scoring(d.c, C.int(s.kind), aff, &la[0], C.int(len(s.matrix)),
			C.int(pk.alpha.Len()), C.int64_t(s.gapOpen), &index[0]); rc != C.BA_OK {
			return fail(fmt.Errorf("%w: %s", ErrDevice, C.GoString(C.ba_last_error(d.c))))
		}
		if rc := C.ba_align_batch(d.c, lp, C.int(pk.stride), ro, rl, qo, ql, C.int64_t(n), &res); rc != C.BA_OK {
			return fail(fmt.Errorf("%w: %s", ErrDevice, C.GoString(C.ba_last_error(d.c))))
		}
		defer C.ba_free_result(d.c, &res)
	}

	status := unsafe.Slice((*int32)(unsafe.Pointer(res.status)), n)
	errPos := unsafe.Slice((*int64)(unsafe.Pointer(res.err_pos)), n)
	segStart := unsafe.Slice((*int64)(unsafe.Pointer(res.seg_start)), n)
	segCount := unsafe.Slice((*int32)(unsafe.Pointer(res.seg_count)), n)
	var segs []int32
	if res.n_segs > 0 {
		segs = unsafe.Slice((*int32)(unsafe.Pointer(res.segs)), 5*int(res.n_segs))
	}
	for p := 0; p < n; p++ {
		switch status[p] {
		case C.BA_PAIR_OK:
			// one backing array per pair: featPair values, handed out as *featPair (align/align.go:121-144)
			fps := make([]featPair, segCount[p])
			aln := make([]feat.Pair, segCount[p])
			for k := range fps {
				s := segs[5*(int(segStart[p])+k):]
				fps[k] = featPair{
					a:     feature{start: int(s[0]), end: int(s[1])},
					b:     feature{start: int(s[2]), end: int(s[3])},
					score: int(s[4]),
				}
				aln[k] = &fps[k]
			}
			alns[p] = aln
		case C.BA_PAIR_ILLEGAL_REF:
			pos := int(errPos[p])
			errs[p] = fmt.Errorf("align: illegal letter %q at position %d in rSeq",
				pk.letters[int(pk.refOff[p])+pos*pk.stride], pos)
		case C.BA_PAIR_ILLEGAL_QRY:
			pos := int(errPos[p])
			errs[p] = fmt.Errorf("align: illegal letter %q at position %d in qSeq",
				pk.letters[int(pk.qryOff[p])+pos*pk.stride], pos)
		case C.BA_PAIR_PANIC:
			// the reference panics on this input (empty sequence or illegal letter in the NW/Fitted
			// border set-up, e.g. align/nw_letters.go:91-96)
			panic(runtime.Error(boundsError{}))
		default:
			panic("align: internal error: no path")
		}
	}
	return alns, errs
}

// boundsError stands in for the runtime's index-out-of-range error value.
type boundsError struct{}

func (boundsError) Error() string   { return "runtime error: index out of range" }
func (boundsError) RuntimeError() {}

// ---- single pairs (the twelve replaced methods) ------------------------------------------------

func (s scoring) alignLetters(rSeq, qSeq alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	runtime.LockOSThread() // the library's contexts are bound to the calling thread's current device
	defer runtime.UnlockOSThread()
	d, err := getCtx()
	if err != nil {
		return nil, err
	}
	defer putCtx(d)
	buf, err := d.buffer(len(rSeq) + len(qSeq))
	if err != nil {
		return nil, err
	}
	copy(buf, alphabet.LettersToBytes(rSeq))
	copy(buf[len(rSeq):], alphabet.LettersToBytes(qSeq))
	pk := &packed{alpha: alpha, stride: 1, letters: buf,
		refOff: []int64{0}, refLen: []int32{int32(len(rSeq))},
		qryOff: []int64{int64(len(rSeq))}, qryLen: []int32{int32(len(qSeq))}}
	a, e := s.run(pk, d, false)
	return a[0], e[0]
}

func (s scoring) alignQLetters(rSeq, qSeq alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	d, err := getCtx()
	if err != nil {
		return nil, err
	}
	defer putCtx(d)
	buf, err := d.buffer(2 * (len(rSeq) + len(qSeq)))
	if err != nil {
		return nil, err
	}
	packQ(buf, rSeq)
	packQ(buf[2*len(rSeq):], qSeq)
	pk := &packed{alpha: alpha, stride: 2, letters: buf,
		refOff: []int64{0}, refLen: []int32{int32(len(rSeq))},
		qryOff: []int64{int64(2 * len(rSeq))}, qryLen: []int32{int32(len(qSeq))}}
	a, e := s.run(pk, d, false)
	return a[0], e[0]
}

// packQ writes QLetters as {L, Q} byte pairs (alphabet/letters.go:182-185); the kernels read L only,
// exactly like align/sw_qletters.go reads x.L.
func packQ(dst []byte, s alphabet.QLetters) {
	for i, l := range s {
		dst[2*i], dst[2*i+1] = byte(l.L), byte(l.Q)
	}
}

func (a SW) alignLetters(rSeq, qSeq alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindSW, false, Linear(a), 0, true}.alignLetters(rSeq, qSeq, alpha)
}
func (a SW) alignQLetters(rSeq, qSeq alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindSW, false, Linear(a), 0, true}.alignQLetters(rSeq, qSeq, alpha)
}
func (a NW) alignLetters(rSeq, qSeq alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindNW, false, Linear(a), 0, true}.alignLetters(rSeq, qSeq, alpha)
}
func (a NW) alignQLetters(rSeq, qSeq alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindNW, false, Linear(a), 0, true}.alignQLetters(rSeq, qSeq, alpha)
}
func (a Fitted) alignLetters(rSeq, qSeq alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindFitted, false, Linear(a), 0, false}.alignLetters(rSeq, qSeq, alpha)
}
func (a Fitted) alignQLetters(rSeq, qSeq alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindFitted, false, Linear(a), 0, false}.alignQLetters(rSeq, qSeq, alpha)
}
func (a SWAffine) alignLetters(rSeq, qSeq alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindSW, true, a.Matrix, a.GapOpen, true}.alignLetters(rSeq, qSeq, alpha)
}
func (a SWAffine) alignQLetters(rSeq, qSeq alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindSW, true, a.Matrix, a.GapOpen, true}.alignQLetters(rSeq, qSeq, alpha)
}
func (a NWAffine) alignLetters(rSeq, qSeq alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindNW, true, a.Matrix, a.GapOpen, true}.alignLetters(rSeq, qSeq, alpha)
}
func (a NWAffine) alignQLetters(rSeq, qSeq alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindNW, true, a.Matrix, a.GapOpen, true}.alignQLetters(rSeq, qSeq, alpha)
}
func (a FittedAffine) alignLetters(rSeq, qSeq alphabet.Letters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindFitted, true, a.Matrix, a.GapOpen, false}.alignLetters(rSeq, qSeq, alpha)
}
func (a FittedAffine) alignQLetters(rSeq, qSeq alphabet.QLetters, alpha alphabet.Alphabet) ([]feat.Pair, error) {
	return scoring{kindFitted, true, a.Matrix, a.GapOpen, false}.alignQLetters(rSeq, qSeq, alpha)
}

// ---- AlignBatch (new) ----------------------------------------------------------------------------

// AlignBatch aligns many independent pairs in one device call per (alphabet, slice type) sub-batch.
// Pairs that fail the wrapper validation (align/sw.go:23-48) get their error without touching the
// GPU.  Batches of MultiDeviceMinPairs pairs or more are spread over every GPU of the box.
func (a SW) AlignBatch(refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	return scoring{kindSW, false, Linear(a), 0, true}.alignBatch(refs, queries)
}
func (a NW) AlignBatch(refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	return scoring{kindNW, false, Linear(a), 0, true}.alignBatch(refs, queries)
}
func (a Fitted) AlignBatch(refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	return scoring{kindFitted, false, Linear(a), 0, false}.alignBatch(refs, queries)
}
func (a SWAffine) AlignBatch(refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	return scoring{kindSW, true, a.Matrix, a.GapOpen, true}.alignBatch(refs, queries)
}
func (a NWAffine) AlignBatch(refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	return scoring{kindNW, true, a.Matrix, a.GapOpen, true}.alignBatch(refs, queries)
}
func (a FittedAffine) AlignBatch(refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	return scoring{kindFitted, true, a.Matrix, a.GapOpen, false}.alignBatch(refs, queries)
}

// subBatch collects the pairs of one (alphabet, slice type) class before they are packed.
type subBatch struct {
	alpha  alphabet.Alphabet
	stride int
	ids    []int // positions in the caller's batch
	bytes  int
}

func (s scoring) alignBatch(refs, queries []AlphabetSlicer) ([][]feat.Pair, []error) {
	n := len(refs)
	alns, errs := make([][]feat.Pair, n), make([]error, n)
	if len(queries) != n {
		for i := range errs {
			errs[i] = errors.New("align: AlignBatch needs as many queries as references")
		}
		return alns, errs
	}
	// ---- wrapper validation per pair, in the reference's order (align/sw.go:23-48), and classes
	var subs []*subBatch
	find := func(alpha alphabet.Alphabet, stride int) *subBatch {
		for _, sb := range subs {
			if sb.alpha == alpha && sb.stride == stride {
				return sb
			}
		}
		sb := &subBatch{alpha: alpha, stride: stride}
		subs = append(subs, sb)
		return sb
	}
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
		switch r := refs[k].Slice().(type) {
		case alphabet.Letters:
			q, ok := queries[k].Slice().(alphabet.Letters)
			if !ok {
				errs[k] = ErrMismatchedTypes
				continue
			}
			sb := find(alpha, 1)
			sb.ids = append(sb.ids, k)
			sb.bytes += len(r) + len(q)
		case alphabet.QLetters:
			q, ok := queries[k].Slice().(alphabet.QLetters)
			if !ok {
				errs[k] = ErrMismatchedTypes
				continue
			}
			sb := find(alpha, 2)
			sb.ids = append(sb.ids, k)
			sb.bytes += 2 * (len(r) + len(q))
		default:
			errs[k] = ErrTypeNotHandled
		}
	}

	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	for _, sb := range subs {
		multi := len(sb.ids) >= MultiDeviceMinPairs
		var d *devCtx
		if multi {
			multiMu.Lock()
			if multiCtx == nil {
				if rc := C.ba_create_multi(nil, 0, &multiCtx); rc != C.BA_OK { // every visible device
					msg := "ba_create_multi failed"
					if multiCtx != nil {
						msg = C.GoString(C.ba_multi_last_error(multiCtx))
						C.ba_destroy_multi(multiCtx)
						multiCtx = nil
					}
					multiMu.Unlock()
					for _, k := range sb.ids {
						errs[k] = fmt.Errorf("%w: %s", ErrDevice, msg)
					}
					continue
				}
			}
			d = &multiBuf
		} else {
			var err error
			if d, err = getCtx(); err != nil {
				for _, k := range sb.ids {
					errs[k] = err
				}
				continue
			}
		}
		a, e := s.packAndRun(sb, refs, queries, d, multi)
		if multi {
			multiMu.Unlock()
		} else {
			putCtx(d)
		}
		for p, k := range sb.ids {
			alns[k], errs[k] = a[p], e[p]
		}
	}
	return alns, errs
}

// packAndRun packs a sub-batch straight into the context's pinned buffer (the layout the library
// uploads as is: reference then query of every pair, in pair order) and runs it.
func (s scoring) packAndRun(sb *subBatch, refs, queries []AlphabetSlicer, d *devCtx, multi bool) ([][]feat.Pair, []error) {
	m := len(sb.ids)
	pk := &packed{alpha: sb.alpha, stride: sb.stride,
		refOff: make([]int64, m), qryOff: make([]int64, m), refLen: make([]int32, m), qryLen: make([]int32, m)}
	buf, err := d.buffer(sb.bytes)
	if err != nil {
		alns, errs := make([][]feat.Pair, m), make([]error, m)
		for i := range errs {
			errs[i] = err
		}
		return alns, errs
	}
	pk.letters = buf
	off := 0
	for p, k := range sb.ids {
		if sb.stride == 1 {
			r, q := refs[k].Slice().(alphabet.Letters), queries[k].Slice().(alphabet.Letters)
			pk.refOff[p], pk.refLen[p] = int64(off), int32(len(r))
			off += copy(buf[off:], alphabet.LettersToBytes(r))
			pk.qryOff[p], pk.qryLen[p] = int64(off), int32(len(q))
			off += copy(buf[off:], alphabet.LettersToBytes(q))
		} else {
			r, q := refs[k].Slice().(alphabet.QLetters), queries[k].Slice().(alphabet.QLetters)
			pk.refOff[p], pk.refLen[p] = int64(off), int32(len(r))
			packQ(buf[off:], r)
			off += 2 * len(r)
			pk.qryOff[p], pk.qryLen[p] = int64(off), int32(len(q))
			packQ(buf[off:], q)
			off += 2 * len(q)
		}
	}
	return s.run(pk, d, multi)
}

// ---- FormatBatch -----------------------------------------------------------------------------------

// FormatBatch is align.Format (align/align.go:146-170) for many alignments in one device call
// (ba_format_batch).  The rows come back as letters; for QLetters the qualities are re-attached
// here along the same segments (the gap QLetter has Q = 0, align/align.go:161).  All sequences of
// one call must be of one slice type.
func FormatBatch(as, bs []seq.Slicer, alns [][]feat.Pair, gap alphabet.Letter) ([][2]alphabet.Slice, error) {
	n := len(as)
	out := make([][2]alphabet.Slice, n)
	if n == 0 {
		return out, nil
	}
	if len(bs) != n || len(alns) != n {
		return nil, errors.New("align: FormatBatch needs as, bs and alns of one length")
	}
	stride, total := 0, 0
	size := func(s alphabet.Slice) (int, error) {
		st := 0
		switch v := s.(type) {
		case alphabet.Letters:
			st, total = 1, total+len(v)
		case alphabet.QLetters:
			st, total = 2, total+2*len(v)
		default:
			return 0, ErrTypeNotHandled
		}
		if stride != 0 && st != stride {
			return 0, ErrMismatchedTypes
		}
		stride = st
		return st, nil
	}
	for k := 0; k < n; k++ {
		if _, err := size(as[k].Slice()); err != nil {
			return nil, err
		}
		if _, err := size(bs[k].Slice()); err != nil {
			return nil, err
		}
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	d, err := getCtx()
	if err != nil {
		return nil, err
	}
	defer putCtx(d)
	buf, err := d.buffer(total)
	if err != nil {
		return nil, err
	}
	refOff, qryOff := make([]int64, n), make([]int64, n)
	refLen, qryLen := make([]int32, n), make([]int32, n)
	segStart, segCount := make([]int64, n), make([]int32, n)
	var segs []int32
	off := 0
	pack := func(s alphabet.Slice) (int64, int32) {
		o := off
		switch v := s.(type) {
		case alphabet.Letters:
			off += copy(buf[off:], alphabet.LettersToBytes(v))
			return int64(o), int32(len(v))
		case alphabet.QLetters:
			packQ(buf[off:], v)
			off += 2 * len(v)
			return int64(o), int32(len(v))
		}
		return int64(o), 0
	}
	for k := 0; k < n; k++ {
		refOff[k], refLen[k] = pack(as[k].Slice())
		qryOff[k], qryLen[k] = pack(bs[k].Slice())
		segStart[k], segCount[k] = int64(len(segs)/5), int32(len(alns[k]))
		for _, fp := range alns[k] {
			f := fp.Features()
			segs = append(segs, int32(f[0].Start()), int32(f[0].End()), int32(f[1].Start()), int32(f[1].End()),
				int32(fp.Score()))
		}
	}
	var res C.ba_formatted
	var lp *C.uint8_t
	if total > 0 {
		lp = (*C.uint8_t)(unsafe.Pointer(&buf[0]))
	} else {
		// every sequence is empty: the library still needs a non-NULL letters pointer to tell this
		// call from "format the previous ba_align_batch"
		var dummy [16]byte
		lp = (*C.uint8_t)(unsafe.Pointer(&dummy[0]))
	}
	var segp *C.int32_t
	if len(segs) > 0 {
		segp = (*C.int32_t)(unsafe.Pointer(&segs[0]))
	}
	rc := C.ba_format_batch(d.c, lp, C.int(stride),
		(*C.int64_t)(unsafe.Pointer(&refOff[0])), (*C.int32_t)(unsafe.Pointer(&refLen[0])),
		(*C.int64_t)(unsafe.Pointer(&qryOff[0])), (*C.int32_t)(unsafe.Pointer(&qryLen[0])), C.int64_t(n),
		(*C.int64_t)(unsafe.Pointer(&segStart[0])), (*C.int32_t)(unsafe.Pointer(&segCount[0])), segp,
		C.int64_t(len(segs)/5), C.uint8_t(gap), &res)
	if rc != C.BA_OK {
		return nil, fmt.Errorf("%w: %s", ErrDevice, C.GoString(C.ba_last_error(d.c)))
	}
	defer C.ba_free_formatted(d.c, &res)
	offA := unsafe.Slice((*int64)(unsafe.Pointer(res.off_a)), n+1)
	offB := unsafe.Slice((*int64)(unsafe.Pointer(res.off_b)), n+1)
	var rowA, rowB []byte
	if res.bytes_a > 0 {
		rowA = unsafe.Slice((*byte)(unsafe.Pointer(res.row_a)), int(res.bytes_a))
	}
	if res.bytes_b > 0 {
		rowB = unsafe.Slice((*byte)(unsafe.Pointer(res.row_b)), int(res.bytes_b))
	}
	for k := 0; k < n; k++ {
		rows := [2][]byte{rowA[offA[k]:offA[k+1]], rowB[offB[k]:offB[k+1]]}
		for side, src := range [2]alphabet.Slice{as[k].Slice(), bs[k].Slice()} {
			switch v := src.(type) {
			case alphabet.Letters:
				out[k][side] = alphabet.Letters(alphabet.BytesToLetters(append([]byte(nil), rows[side]...)))
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
	return out, nil
}
