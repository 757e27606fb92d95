"""Host-side mirror of bíogo's align/pals/dp package on top of ba_pals_dp_batch.

    aligner = pals.NewAligner(target, query, k, minLength, minId)     # align/pals/dp/align.go:43-52
    aligner.Costs = pals.Costs(MaxIGap=5, DiffCost=3, SameCost=1, MatchCost=4, BlockCost=15, RMatchCost=4.0)
    hits = aligner.AlignTraps(trapezoids)                              # align.go:55-141 -> [Hit]

``AlignTrapsBatch`` is the new entry point: many (target, query, trapezoids) problems in one device call.
Sequences are ``seq.Seq``-like objects (``.Seq`` letters, ``.Alpha``) or plain ``bytes`` with an alphabet.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class Trapezoid:
    """filter.Trapezoid (align/pals/filter/trapezoid.go:8-11)"""
    Top: int
    Bottom: int
    Left: int
    Right: int


@dataclass
class Costs:
    """dp.Costs (align/pals/dp/align.go:24-31)"""
    MaxIGap: int
    DiffCost: int
    SameCost: int
    MatchCost: int
    BlockCost: int
    RMatchCost: float


@dataclass
class Hit:
    """dp.Hit (align/pals/dp/align.go:143-150)"""
    Abpos: int
    Bbpos: int
    Aepos: int
    Bepos: int
    LowDiagonal: int
    HighDiagonal: int
    Score: int
    Error: float


def Sum(hits):
    """Hits.Sum (align.go:155-165): summed lengths on both sequences."""
    a = b = 0
    for h in hits:
        la, lb = h.Aepos - h.Abpos, h.Bepos - h.Bbpos
        if la < 0 or lb < 0:
            raise ValueError("dp: negative trapezoid area")
        a, b = a + la, b + lb
    return a, b


def _letters(s) -> bytes:
    if isinstance(s, (bytes, bytearray)):
        return bytes(s)
    v = s.Seq
    return bytes(v) if not hasattr(v, "raw") else bytes(v.raw)


class Aligner:
    def __init__(self, target, query, k: int, minLength: int, minId: float, ctx: "_lib.Context | None" = None,
                 alpha=None):
        self.target, self.query, self.k = target, query, k
        self.minHitLength, self.minId = minLength, minId
        self.Costs: Costs | None = None
        self._ctx = ctx
        self._alpha = alpha if alpha is not None else getattr(target, "Alpha", None)

    def AlignTraps(self, trapezoids):
        return AlignTrapsBatch([self], [trapezoids], self._ctx)[0]


def NewAligner(target, query, k: int, minLength: int, minId: float, ctx=None, alpha=None) -> Aligner:
    return Aligner(target, query, k, minLength, minId, ctx, alpha)


def AlignTrapsBatch(aligners, trapezoid_lists, ctx=None):
    """All problems in one ba_pals_dp_batch call.  The aligners must share k, minLength, minId, Costs and alphabet."""
    if not aligners:
        return []
    a0 = aligners[0]
    if a0.Costs is None:
        raise ValueError("dp: Aligner.Costs is not set")     # the reference dereferences a nil *Costs
    for a in aligners:
        if (a.k, a.minHitLength, a.minId, a.Costs, a._alpha) != (a0.k, a0.minHitLength, a0.minId, a0.Costs, a0._alpha):
            raise ValueError("AlignTrapsBatch: aligners of one batch share k, minLength, minId, Costs and alphabet")
    ctx = ctx or a0._ctx or _lib.Context(0)
    chunks, to, tl, qo, ql, troff, traps = [], [], [], [], [], [0], []
    off = 0
    for a, tz in zip(aligners, trapezoid_lists):
        t, q = _letters(a.target), _letters(a.query)
        to.append(off)
        tl.append(len(t))
        off += len(t)
        qo.append(off)
        ql.append(len(q))
        off += len(q)
        chunks += [t, q]
        for z in tz:
            traps += [z.Top, z.Bottom, z.Left, z.Right] if isinstance(z, Trapezoid) else list(z)
        troff.append(len(traps) // 4)
    letters = np.frombuffer(b"".join(chunks) + b"\0", dtype=np.uint8)[:-1]
    c = a0.Costs
    hits, _ = ctx.pals_dp_batch(letters, to, tl, qo, ql, a0._alpha.LetterIndex(), troff, traps, a0.k, a0.minHitLength,
                                a0.minId, (c.MaxIGap, c.DiffCost, c.SameCost, c.MatchCost, c.BlockCost, c.RMatchCost))
    return [[Hit(*h) for h in hs] for hs in hits]
