"""seq/linear containers that implement align.AlphabetSlicer.

Mirrors linear.Seq (seq/linear/seq.go:17-57) and linear.QSeq (seq/linear/qseq.go:17-64)
down to what Align needs: an ``Alpha`` field, ``Alphabet()`` (seq/annotation.go:75) and
``Slice()``.  Everything else in biogo/seq is out of scope.
"""
from __future__ import annotations

from . import alphabet as _alphabet


class Seq:
    """linear.Seq{Annotation; Seq alphabet.Letters}"""

    def __init__(self, Seq=b"", Alpha=None, ID: str = ""):
        self.ID = ID
        self.Alpha = Alpha
        self.Seq = Seq if isinstance(Seq, _alphabet.Letters) else _alphabet.Letters(Seq)

    def Alphabet(self):
        return self.Alpha

    def Slice(self):
        return self.Seq

    def Len(self) -> int:
        return len(self.Seq)


class QSeq:
    """linear.QSeq{Annotation; Seq alphabet.QLetters; ...}"""

    def __init__(self, Seq=None, Alpha=None, ID: str = ""):
        self.ID = ID
        self.Alpha = Alpha
        self.Seq = Seq if isinstance(Seq, _alphabet.QLetters) else _alphabet.QLetters(Seq or b"")

    def Alphabet(self):
        return self.Alpha

    def Slice(self):
        return self.Seq

    def Len(self) -> int:
        return self.Seq.Len()


def NewSeq(id: str, b: bytes, alpha) -> Seq:
    return Seq(Seq=_alphabet.Letters(b), Alpha=alpha, ID=id)


def NewQSeq(id: str, ql: "_alphabet.QLetters", alpha) -> QSeq:
    return QSeq(Seq=ql, Alpha=alpha, ID=id)
