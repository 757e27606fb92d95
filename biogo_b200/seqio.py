"""FASTA / FASTQ text -> packed batch: the host-side mirror of bíogo's io/seqio readers
(io/seqio/fasta/fasta.go, io/seqio/fastq/fastq.go) on top of ba_fasta_parse / ba_fastq_parse.

The whole text is parsed in one native call into one letters buffer with offset arrays -- the
layout ba_align_batch takes -- instead of one object per record.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib, alphabet as _alphabet, seq as _seq

SANGER, ILLUMINA_1_3 = 33, 64


class FormatError(ValueError):
    """Malformed FASTA / FASTQ text; ``line`` is the 1-based line number."""

    def __init__(self, msg: str, line: int):
        super().__init__(msg)
        self.line = line


@dataclass
class SeqFile:
    """Parsed records: letters of record k = letters[seq_off[k] : seq_off[k] + stride*seq_len[k]]."""
    stride: int
    letters: np.ndarray
    seq_off: np.ndarray
    seq_len: np.ndarray
    names: list
    descs: list

    def __len__(self):
        return len(self.seq_off)

    def raw(self, k: int) -> bytes:
        o = int(self.seq_off[k])
        return self.letters[o:o + self.stride * int(self.seq_len[k])].tobytes()

    def sequences(self, alpha):
        """seq.Seq objects like linear.NewSeq / linear.NewQSeq filled by the reference readers."""
        out = []
        for k in range(len(self)):
            raw = self.raw(k)
            if self.stride == 2:
                s = _seq.QSeq(Seq=_alphabet.QLetters(raw=raw), Alpha=alpha, ID=self.names[k])
            else:
                s = _seq.Seq(Seq=_alphabet.Letters(raw), Alpha=alpha, ID=self.names[k])
            s.Desc = self.descs[k]
            out.append(s)
        return out

    def pairs(self, ref_idx, qry_idx):
        """(letters, ref_off, ref_len, qry_off, qry_len) for Context.align_packed."""
        r, q = np.asarray(ref_idx, np.int64), np.asarray(qry_idx, np.int64)
        return self.letters, self.seq_off[r], self.seq_len[r], self.seq_off[q], self.seq_len[q]


def _parse(fn, data: bytes, arg: int, lib_path=None) -> SeqFile:
    lib = _lib.load(lib_path)
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    f = _lib.BaSeqFile()
    rc = getattr(lib, fn)(buf.ctypes.data, len(data), arg, C.byref(f))
    try:
        if rc == _lib.BA_ERR_FORMAT:
            raise FormatError(f.err.decode(), int(f.err_line))
        if rc != _lib.BA_OK:
            raise _lib.AlignError(rc, fn + " failed")
        n = int(f.n)
        text = C.string_at(f.text, int(f.text_bytes)) if f.text_bytes else b""
        no, nl = _lib._as_array(f.name_off, n, np.int64), _lib._as_array(f.name_len, n, np.int32)
        do, dl = _lib._as_array(f.desc_off, n, np.int64), _lib._as_array(f.desc_len, n, np.int32)
        return SeqFile(int(f.stride), _lib._as_array(f.letters, int(f.letters_bytes), np.uint8),
                       _lib._as_array(f.seq_off, n, np.int64), _lib._as_array(f.seq_len, n, np.int32),
                       [text[no[k]:no[k] + nl[k]].decode("latin-1") for k in range(n)],
                       [text[do[k]:do[k] + dl[k]].decode("latin-1") for k in range(n)])
    finally:
        lib.ba_seqfile_free(C.byref(f))


def parse_fasta(data: bytes, qletters: bool = False, lib_path=None) -> SeqFile:
    return _parse("ba_fasta_parse", data, 2 if qletters else 1, lib_path)


def parse_fastq(data: bytes, quality_offset: int = SANGER, lib_path=None) -> SeqFile:
    return _parse("ba_fastq_parse", data, quality_offset, lib_path)
