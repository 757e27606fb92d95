"""Host-side mirror of biogo's ``align`` package surface for the DP hot path.

Same names, argument meaning and error behaviour as the Go package so that tests read
like the reference's own (align/sw_example_test.go etc.):

    smith = align.SWAffine(Matrix=[[...]], GapOpen=-5)
    aln, err = smith.Align(swsa, swsb)          # ([]feat.Pair, error)
    fa = align.Format(swsa, swsb, aln, '-')

* ``Aligner.Align(reference, query)``   align/align.go:28-30 and the six wrappers
  align/sw.go:22-49, nw.go:20-47, fitted.go:21-48, sw_affine.go:20-47, nw_affine.go:20-47,
  fitted_affine.go:21-48 (validation order reproduced below).
* ``AlignBatch(references, queries)``   new: one C-ABI call for many independent pairs.
* sentinel errors / ErrMatrixWrongSize  align/align.go:57-73 (compared by identity / value).
* ``Format``                            align/align.go:148-170.

Go panics (index out of range on empty or illegal input in the NW/Fitted border set-up,
"internal error: no path") surface as the Python exception ``GoPanic``.
The DP itself always runs in the CUDA library; nothing here computes alignments.
"""
from __future__ import annotations

import threading

import numpy as np

from . import _lib
from . import alphabet as _alphabet
from .feat import Pair


class AlignErr(Exception):
    """An error *value* returned (not raised) by Align, like a Go error."""

    def Error(self) -> str:
        return str(self)


# align/align.go:57-64 -- compared by identity
ErrMismatchedTypes = AlignErr("align: mismatched sequence types")
ErrMismatchedAlphabets = AlignErr("align: mismatched alphabets")
ErrNoAlphabet = AlignErr("align: no alphabet")
ErrNotGappedAlphabet = AlignErr("align: alphabet does not have gap at position 0")
ErrTypeNotHandled = AlignErr("align: sequence type not handled")
ErrMatrixNotSquare = AlignErr("align: scoring matrix is not square")


class ErrMatrixWrongSize(AlignErr):
    """align/align.go:66-73 -- compared by value."""

    def __init__(self, Size: int, Len: int):
        super().__init__(f"align: scoring matrix size {Size} does not match alphabet length {Len}")
        self.Size, self.Len = Size, Len

    def __eq__(self, other):
        return isinstance(other, ErrMatrixWrongSize) and (other.Size, other.Len) == (self.Size, self.Len)

    def __hash__(self):
        return hash((self.Size, self.Len))


class IllegalLetterError(AlignErr):
    """fmt.Errorf("align: illegal letter %q at position %d in rSeq|qSeq") (align/sw_letters.go:97-102)."""

    def __init__(self, letter: int, position: int, which: str):
        super().__init__(f"align: illegal letter {_go_quote(letter)} at position {position} in {which}")
        self.letter, self.position, self.which = letter, position, which


class GoPanic(RuntimeError):
    """The reference panics on this input; the drop-in does the same."""


def _go_quote(b: int) -> str:
    # %q of an alphabet.Letter (a byte): Go-syntax single-quoted rune literal
    ch = chr(b)
    if ch == "'":
        return "'\\''"
    if ch == "\\":
        return "'\\\\'"
    if 0x20 <= b < 0x7F:
        return f"'{ch}'"
    esc = {7: "\\a", 8: "\\b", 12: "\\f", 10: "\\n", 13: "\\r", 9: "\\t", 11: "\\v"}
    if b in esc:
        return f"'{esc[b]}'"
    if b < 0x80:
        return f"'\\x{b:02x}'"
    return f"'{ch}'" if ch.isprintable() else f"'\\u{b:04x}'"


# ---------------------------------------------------------------------------
# context handling: one ba_ctx per (thread, device); Align is reentrant like the reference
_tls = threading.local()
_default_device = 0
_default_lib_path = None


def set_default_device(device: int) -> None:
    global _default_device
    _default_device = device


def set_library_path(path) -> None:
    """Tests only: point the mirror at another build of the same C ABI."""
    global _default_lib_path
    _default_lib_path = path
    _tls.__dict__.clear()


def context(device=None) -> _lib.Context:
    device = _default_device if device is None else device
    key = (device, _default_lib_path)
    ctxs = _tls.__dict__.setdefault("ctxs", {})
    if key not in ctxs:
        ctxs[key] = _lib.Context(device, _default_lib_path)
    return ctxs[key]


# ---------------------------------------------------------------------------
def _flatten(matrix, alpha, check_size: bool):
    """The matrix check at the top of every alignLetters (align/sw_letters.go:69-79)."""
    let = len(matrix)
    if check_size and let < alpha.Len():
        return None, let, ErrMatrixWrongSize(Size=let, Len=alpha.Len())
    la = []
    for row in matrix:
        if len(row) != let:
            return None, let, ErrMatrixNotSquare
        la.extend(int(v) for v in row)
    return la, let, None


def _slice_bytes(sl):
    if isinstance(sl, _alphabet.QLetters):
        return sl.raw, 2
    return bytes(sl), 1


class _Aligner:
    _kind = _lib.BA_SW
    _affine = False
    _check_size = True
    _name = "sw"

    # overridden by Linear / Affine
    def _matrix(self):
        raise NotImplementedError

    def _gap_open(self) -> int:
        return 0

    def _validate(self, reference, query):
        """The wrapper's checks, in the reference's order (align/sw.go:23-48)."""
        alpha = reference.Alphabet()
        if alpha is None:
            return None, None, None, ErrNoAlphabet
        if alpha is not query.Alphabet():
            return None, None, None, ErrMismatchedAlphabets
        if alpha.IndexOf(alpha.Gap()) != 0:
            return None, None, None, ErrNotGappedAlphabet
        rs = reference.Slice()
        qs = query.Slice()
        if isinstance(rs, _alphabet.Letters):
            if not isinstance(qs, _alphabet.Letters):
                return None, None, None, ErrMismatchedTypes
        elif isinstance(rs, _alphabet.QLetters):
            if not isinstance(qs, _alphabet.QLetters):
                return None, None, None, ErrMismatchedTypes
        else:
            return None, None, None, ErrTypeNotHandled
        return alpha, rs, qs, None

    def _upload(self, ctx, alpha):
        la, let, err = _flatten(self._matrix(), alpha, self._check_size)
        if err is not None:
            return err
        ctx.set_scoring(self._kind, self._affine, la, let, alpha.Len(), self._gap_open(), alpha.LetterIndex())
        return None

    def _pair_outcome(self, res, p, rs_raw, qs_raw, stride):
        st = int(res.status[p])
        if st == _lib.BA_PAIR_OK:
            return [Pair(*map(int, row)) for row in res.pairs(p)], None
        if st == _lib.BA_PAIR_ILLEGAL_REF:
            pos = int(res.err_pos[p])
            return None, IllegalLetterError(rs_raw[pos * stride], pos, "rSeq")
        if st == _lib.BA_PAIR_ILLEGAL_QRY:
            pos = int(res.err_pos[p])
            return None, IllegalLetterError(qs_raw[pos * stride], pos, "qSeq")
        if st == _lib.BA_PAIR_PANIC:
            raise GoPanic("runtime error: index out of range")
        raise GoPanic(f"align: {self._name} internal error: no path")

    def Align(self, reference, query):
        """([]feat.Pair, error) -- align.Aligner (align/align.go:28-30)."""
        alpha, rs, qs, err = self._validate(reference, query)
        if err is not None:
            return None, err
        ctx = context()
        err = self._upload(ctx, alpha)
        if err is not None:
            return None, err
        r_raw, stride = _slice_bytes(rs)
        q_raw, _ = _slice_bytes(qs)
        letters = np.frombuffer(r_raw + q_raw, dtype=np.uint8)
        res = ctx.align_packed(letters, stride, [0], [len(r_raw) // stride], [len(r_raw)], [len(q_raw) // stride])
        return self._pair_outcome(res, 0, r_raw, q_raw, stride)

    def AlignBatch(self, references, queries):
        """([][]feat.Pair, []error): one batched device call for many independent pairs.

        Pairs that fail the wrapper's validation get their error without touching the GPU;
        all others must share one alphabet and one slice type (one scoring upload per call).
        """
        if len(references) != len(queries):
            raise ValueError("align: AlignBatch needs as many queries as references")
        n = len(references)
        alns, errs = [None] * n, [None] * n
        alpha0, stride0 = None, None
        chunks, ref_off, ref_len, qry_off, qry_len, live, raws = [], [], [], [], [], [], []
        off = 0
        for k in range(n):
            alpha, rs, qs, err = self._validate(references[k], queries[k])
            if err is not None:
                errs[k] = err
                continue
            r_raw, stride = _slice_bytes(rs)
            q_raw, _ = _slice_bytes(qs)
            if alpha0 is None:
                alpha0, stride0 = alpha, stride
            elif alpha is not alpha0 or stride != stride0:
                raise ValueError("align: AlignBatch needs one alphabet and one sequence type per call")
            chunks += [r_raw, q_raw]
            ref_off.append(off)
            ref_len.append(len(r_raw) // stride)
            off += len(r_raw)
            qry_off.append(off)
            qry_len.append(len(q_raw) // stride)
            off += len(q_raw)
            live.append(k)
            raws.append((r_raw, q_raw))
        if not live:
            return alns, errs
        ctx = context()
        err = self._upload(ctx, alpha0)
        if err is not None:
            for k in live:
                errs[k] = err
            return alns, errs
        letters = np.frombuffer(b"".join(chunks), dtype=np.uint8)
        res = ctx.align_packed(letters, stride0, ref_off, ref_len, qry_off, qry_len)
        for p, k in enumerate(live):
            alns[k], errs[k] = self._pair_outcome(res, p, raws[p][0], raws[p][1], stride0)
        return alns, errs


class Linear(list):
    """align.Linear [][]int: square matrix, first row/column = gap penalties (align/align.go:32-34)."""


class Affine:
    """align.Affine{Matrix Linear; GapOpen int} (align/align.go:36-40)."""

    def __init__(self, Matrix=None, GapOpen: int = 0):
        self.Matrix = Linear(Matrix or [])
        self.GapOpen = GapOpen


class _LinearAligner(Linear, _Aligner):
    def _matrix(self):
        return self


class _AffineAligner(Affine, _Aligner):
    _affine = True

    def _matrix(self):
        return self.Matrix

    def _gap_open(self) -> int:
        return int(self.GapOpen)


class SW(_LinearAligner):
    """Smith-Waterman, linear gaps (align/sw.go:15-49)."""
    _kind, _name = _lib.BA_SW, "sw"


class NW(_LinearAligner):
    """Needleman-Wunsch, linear gaps (align/nw.go:13-47)."""
    _kind, _name = _lib.BA_NW, "nw"


class Fitted(_LinearAligner):
    """Fitted (query global, reference local), linear gaps (align/fitted.go:13-48)."""
    _kind, _name, _check_size = _lib.BA_FITTED, "fitted", False


class SWAffine(_AffineAligner):
    """Smith-Waterman, affine gaps (align/sw_affine.go:13-47)."""
    _kind, _name = _lib.BA_SW, "sw affine"


class NWAffine(_AffineAligner):
    """Needleman-Wunsch, affine gaps (align/nw_affine.go:13-47)."""
    _kind, _name = _lib.BA_NW, "nw affine"


class FittedAffine(_AffineAligner):
    """Fitted, affine gaps (align/fitted_affine.go:13-48)."""
    _kind, _name, _check_size = _lib.BA_FITTED, "fitted affine", False


def Format(a, b, f, gap):
    """align.Format (align/align.go:148-170): the two gapped rows of an alignment."""
    if isinstance(gap, str):
        gap = ord(gap)
    src = [a.Slice(), b.Slice()]
    out = [s.Make(0, 0) for s in src]
    for fs in f:
        fc = fs.Features()
        for i in range(2):
            if fc[i].Len() == 0:
                n = fc[1 - i].Len()
                if isinstance(out[i], _alphabet.QLetters):
                    out[i] = out[i].Append(_alphabet.QLetters(bytes([gap]) * n))
                else:
                    out[i] = out[i].Append(_alphabet.Letters(bytes([gap]) * n))
            else:
                out[i] = out[i].Append(src[i].Slice(fc[i].Start(), fc[i].End()))
    return out


def FormatBatch(a_list, b_list, alns, gap):
    """align.Format for many alignments in one device call (ba_format_batch; the reference has only
    the per-alignment Format, align/align.go:146-170).  ``alns[k]`` is the []feat.Pair of pair k
    (None or empty -> two empty rows).  Returns a list of [row_a, row_b] like Format; rows carry
    letters only; the qualities of QLetters rows are re-attached on the host."""
    if isinstance(gap, str):
        gap = ord(gap)
    n = len(a_list)
    if len(b_list) != n or len(alns) != n:
        raise ValueError("align: FormatBatch needs equally long lists")
    if n == 0:
        return []
    chunks, ref_off, ref_len, qry_off, qry_len = [], [], [], [], []
    seg_start, seg_count, segs = [], [], []
    stride0, off = None, 0
    for k in range(n):
        r_raw, stride = _slice_bytes(a_list[k].Slice())
        q_raw, stride_q = _slice_bytes(b_list[k].Slice())
        if stride != stride_q or (stride0 is not None and stride != stride0):
            raise ValueError("align: FormatBatch needs one sequence type per call")
        stride0 = stride
        chunks += [r_raw, q_raw]
        ref_off.append(off)
        ref_len.append(len(r_raw) // stride)
        off += len(r_raw)
        qry_off.append(off)
        qry_len.append(len(q_raw) // stride)
        off += len(q_raw)
        seg_start.append(len(segs))
        f = alns[k] or []
        seg_count.append(len(f))
        for fs in f:
            fc = fs.Features()
            segs.append((fc[0].Start(), fc[0].End(), fc[1].Start(), fc[1].End(), fs.Score()))

    class _Res:
        pass
    res = _Res()
    res.seg_start = np.array(seg_start, np.int64)
    res.seg_count = np.array(seg_count, np.int32)
    res.segs = np.array(segs, np.int32).reshape(-1, 5)
    letters = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    off_a, off_b, row_a, row_b = context().format_packed(letters, stride0, ref_off, ref_len, qry_off, qry_len, res, gap)
    out = []
    for k in range(n):
        rows = []
        for src, off_x, row_x in ((a_list[k], off_a, row_a), (b_list[k], off_b, row_b)):
            raw = row_x[off_x[k]:off_x[k + 1]].tobytes()
            sl = src.Slice()
            if isinstance(sl, _alphabet.QLetters):
                # the device rows carry letters; qualities follow the same segments on the host
                # (gap QLetter has Q = 0, align/align.go:161)
                side = 0 if src is a_list[k] else 1
                q = bytearray()
                for fs in (alns[k] or []):
                    fc = fs.Features()
                    if fc[side].Len() == 0:
                        q += bytes(fc[1 - side].Len())
                    else:
                        q += sl.raw[2 * fc[side].Start() + 1:2 * fc[side].End():2]
                rows.append(_alphabet.QLetters(raw, bytes(q)))
            else:
                rows.append(_alphabet.Letters(raw))
        out.append(rows)
    return out
