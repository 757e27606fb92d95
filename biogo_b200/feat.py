"""The feat.Pair result type of Align.

Mirrors ``feature`` / ``featPair`` of align/align.go:99-144 (which implement feat.Feature
and feat.Pair of feat/feature.go:74,97): unnamed half-open intervals on the reference (a)
and the query (b) plus the segment score.
"""
from __future__ import annotations


class Feature:
    __slots__ = ("start", "end", "loc")

    def __init__(self, start: int, end: int, loc=None):
        self.start, self.end, self.loc = start, end, loc

    def Name(self) -> str:
        return self.loc.Name() if self.loc is not None else ""

    def Description(self) -> str:
        return self.loc.Description() if self.loc is not None else ""

    def Location(self):
        return self.loc

    def Start(self) -> int:
        return self.start

    def End(self) -> int:
        return self.end

    def Len(self) -> int:
        return self.end - self.start


class Pair:
    """featPair (align/align.go:121-144)."""

    __slots__ = ("a", "b", "score")

    def __init__(self, a_start: int, a_end: int, b_start: int, b_end: int, score: int):
        self.a = Feature(a_start, a_end)
        self.b = Feature(b_start, b_end)
        self.score = score

    def Features(self):
        return (self.a, self.b)

    def Score(self) -> int:
        return self.score

    def Invert(self) -> None:
        self.a, self.b = self.b, self.a

    def String(self) -> str:
        if self.a.start == self.a.end:
            return f"-/{self.b.Name()}[{self.b.start},{self.b.end})={self.score}"
        if self.b.start == self.b.end:
            return f"{self.a.Name()}[{self.a.start},{self.a.end})/-={self.score}"
        return (
            f"{self.a.Name()}[{self.a.start},{self.a.end})/"
            f"{self.b.Name()}[{self.b.start},{self.b.end})={self.score}"
        )

    __str__ = String
    __repr__ = String

    def astuple(self):
        return (self.a.start, self.a.end, self.b.start, self.b.end, self.score)

    def __eq__(self, other) -> bool:
        return isinstance(other, Pair) and other.astuple() == self.astuple()


def Sprint(pairs) -> str:
    """fmt.Sprint of a []feat.Pair: ``[p0 p1 ...]``."""
    return "[" + " ".join(str(p) for p in pairs) + "]"
