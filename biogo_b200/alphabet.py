"""Boundary types of biogo/alphabet that the align hot path touches.

Mirrors: Letter/Letters (alphabet/letters.go:98,128), QLetter/QLetters
(alphabet/letters.go:182,211), the Alphabet interface subset used by align
(alphabet/alphabet.go:116-160: Len, IndexOf, Gap, LetterIndex) and the predefined
alphabets (alphabet/alphabet.go:26-79).  Quality maths, pairing and complementing are
out of scope (never called by align).
"""
from __future__ import annotations

import numpy as np


class Letters(bytes):
    """alphabet.Letters: a slice of single-byte letters (alphabet/letters.go:128-136)."""

    def Len(self) -> int:
        return len(self)

    def Slice(self, start: int, end: int) -> "Letters":
        return Letters(self[start:end])

    def Append(self, src: "Letters") -> "Letters":
        return Letters(bytes(self) + bytes(src))

    def Make(self, length: int, cap: int = 0) -> "Letters":
        return Letters(bytes(length))

    def String(self) -> str:
        return self.decode("latin-1")

    def __str__(self) -> str:  # fmt.Printf("%s", letters)
        return self.decode("latin-1")


class QLetters:
    """alphabet.QLetters: a slice of {L Letter; Q Qphred} pairs (alphabet/letters.go:182-219).

    Stored exactly like Go's []QLetter: interleaved (L, Q) bytes, 2 bytes per element.
    """

    __slots__ = ("raw",)

    def __init__(self, letters: bytes = b"", quals=None, raw: bytes | None = None):
        if raw is not None:
            self.raw = bytes(raw)
            return
        letters = bytes(letters)
        if quals is None:
            quals = bytes(len(letters))
        elif isinstance(quals, int):
            quals = bytes([quals]) * len(letters)
        quals = bytes(quals)
        if len(quals) != len(letters):
            raise ValueError("alphabet: letters and qualities differ in length")
        a = np.empty((len(letters), 2), dtype=np.uint8)
        a[:, 0] = np.frombuffer(letters, dtype=np.uint8)
        a[:, 1] = np.frombuffer(quals, dtype=np.uint8)
        self.raw = a.tobytes()

    def Len(self) -> int:
        return len(self.raw) // 2

    __len__ = Len

    def L(self) -> bytes:
        return self.raw[0::2]

    def Q(self) -> bytes:
        return self.raw[1::2]

    def Slice(self, start: int, end: int) -> "QLetters":
        return QLetters(raw=self.raw[2 * start : 2 * end])

    def Append(self, src: "QLetters") -> "QLetters":
        return QLetters(raw=self.raw + src.raw)

    def Make(self, length: int, cap: int = 0) -> "QLetters":
        return QLetters(raw=bytes(2 * length))

    def String(self) -> str:
        return self.L().decode("latin-1")

    __str__ = String

    def __eq__(self, other) -> bool:
        return isinstance(other, QLetters) and other.raw == self.raw


class Alphabet:
    """The subset of alphabet.Alphabet align uses (alphabet/alphabet.go:183-255)."""

    def __init__(self, letters: str, gap: str, ambiguous: str, case_sensitive: bool = False, name: str = ""):
        if any(ord(ch) > 127 for ch in letters):
            raise ValueError("alphabet: letters contains non-ASCII rune")
        self._length = len(letters)
        self._gap = ord(gap)
        self._ambiguous = ord(ambiguous)
        self._case_sensitive = case_sensitive
        self._name = name
        index = np.full(256, -1, dtype=np.int32)
        if case_sensitive:
            self._letters = letters
            for i, ch in enumerate(letters):
                index[ord(ch)] = i
        else:
            # alphabet.go:206-216: lower-case letters get their position, the
            # upper-case twins copy the index of the lower-case form
            lower, upper = letters.lower(), letters.upper()
            self._letters = lower + upper
            for i, ch in enumerate(lower):
                index[ord(ch)] = i
            for i, ch in enumerate(upper):
                index[ord(ch)] = index[ord(lower[i])]
        self._index = index

    def Len(self) -> int:
        return self._length

    def IsCased(self) -> bool:
        return self._case_sensitive

    def Gap(self) -> int:
        return self._gap

    def Ambiguous(self) -> int:
        return self._ambiguous

    def IndexOf(self, letter: int) -> int:
        return int(self._index[letter])

    def IsValid(self, letter: int) -> bool:
        return self._index[letter] >= 0

    def LetterIndex(self) -> np.ndarray:
        """alphabet.Index: 256 ints, -1 for letters outside the alphabet."""
        return self._index

    def Letters(self) -> str:
        return self._letters

    def __repr__(self) -> str:
        return f"alphabet.{self._name or 'Alphabet'}"


# alphabet/alphabet.go:26-79
DNA = Alphabet("acgt", "-", "n", name="DNA")
DNAgapped = Alphabet("-acgt", "-", "n", name="DNAgapped")
DNAredundant = Alphabet("-acmgrsvtwyhkdbn", "-", "n", name="DNAredundant")
RNA = Alphabet("acgu", "-", "n", name="RNA")
RNAgapped = Alphabet("-acgu", "-", "n", name="RNAgapped")
RNAredundant = Alphabet("-acmgrsvuwyhkdbn", "-", "n", name="RNAredundant")
Protein = Alphabet("-abcdefghijklmnpqrstvwxyz*", "-", "x", name="Protein")


def BytesToLetters(b: bytes) -> Letters:
    return Letters(b)
