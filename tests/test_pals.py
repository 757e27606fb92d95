"""align/pals/dp: banded x-drop DP over filter trapezoids.

Pin: the reference's own golden test align/pals/dp/dp_test.go:22-84 -- target = de Bruijn sequence B(4, 6),
query = B(4, 5) (util.DeBruijn, util/debruijn.go:8-40), six trapezoids, costs (5, 3, 1, 4, 15, 4.0), k = 6,
minLength 50, minId 0.80 -> five hits with their float64 Error values, Sum() = (791, 664).  The oracle
(oracle/pals_oracle.c) must reproduce it, and the device kernel must equal the oracle on it and on random
problems (emulator tier here, GPU tier on the B200)."""
import random

import numpy as np
import pytest

import helpers
from biogo_b200 import _lib, alphabet, pals
from oracle import oracle as O

T = [(452, 0, -128, 3), (433, 237, -1120, -1085), (628, 447, -1984, -1917), (898, 627, -2624, -2557),
     (939, 868, -2880, -2845), (1024, 938, -3072, -3005)]
H = [(1, 0, 290, 242, -3, 54, 101, 0.19421487603305784),
     (365, 286, 435, 345, 74, 96, 26, 0.1864406779661017),
     (437, 341, 507, 400, 91, 113, 26, 0.1864406779661017),
     (3201, 642, 3477, 873, 2553, 2610, 96, 0.19480519480519481),
     (3980, 948, 4066, 1021, 3026, 3054, 30, 0.1917808219178082)]
COSTS = (5, 3, 1, 4, 15, 4.0)
DNA_IDX = np.array(alphabet.DNA.LetterIndex(), np.int32)


def _debruijn_pair():
    a = bytes(b"ACGT"[i] for i in O.debruijn(4, 6))
    b = bytes(b"ACGT"[i] for i in O.debruijn(4, 5))
    assert len(a) == 4096 and len(b) == 1024
    return a, b


def test_oracle_reproduces_the_reference_golden_hits(built):
    a, b = _debruijn_pair()
    hits = O.pals_align_traps(a, b, DNA_IDX, T, 6, 50, 0.80, COSTS)
    assert hits == H                                   # exact, including the float64 Error values
    assert (sum(h[2] - h[0] for h in hits), sum(h[3] - h[1] for h in hits)) == (791, 664)   # Hits.Sum, dp_test.go:80-83


def _random_problem(rng, tlen, qlen, ntraps):
    """a query with planted, mutated copies of target pieces, and trapezoids around the planted diagonals"""
    tgt = bytes(rng.choice(b"ACGT") for _ in range(tlen))
    qry = bytearray(rng.choice(b"ACGT") for _ in range(qlen))
    traps = []
    for _ in range(ntraps):
        ln = rng.randint(30, max(31, min(qlen, tlen) // 2))
        ts, qs = rng.randint(0, tlen - ln), rng.randint(0, qlen - ln)
        piece = bytearray(tgt[ts:ts + ln])
        for x in range(ln):
            if rng.random() < 0.06:
                piece[x] = rng.choice(b"ACGTN")
        qry[qs:qs + ln] = piece
        diag = qs - ts                                # B - A
        w = rng.randint(2, 40)
        traps.append((min(qlen, qs + ln + rng.randint(0, 20)), max(0, qs - rng.randint(0, 20)), diag - w, diag + w))
    traps.sort(key=lambda z: z[1])                     # filter trapezoids arrive sorted by Bottom
    return tgt, bytes(qry), traps


def _check_batch(ctx, problems, kmer, minlen, minid, costs):
    aligners, tzs = [], []
    for tgt, qry, traps in problems:
        a = pals.NewAligner(tgt, qry, kmer, minlen, minid, ctx, alphabet.DNA)
        a.Costs = pals.Costs(*costs)
        aligners.append(a)
        tzs.append(traps)
    got = pals.AlignTrapsBatch(aligners, tzs, ctx)
    for (tgt, qry, traps), g in zip(problems, got):
        want = O.pals_align_traps(tgt, qry, DNA_IDX, traps, kmer, minlen, minid, costs)
        assert [(h.Abpos, h.Bbpos, h.Aepos, h.Bepos, h.LowDiagonal, h.HighDiagonal, h.Score, h.Error) for h in g] == want


def _suite(ctx):
    a, b = _debruijn_pair()
    al = pals.NewAligner(a, b, 6, 50, 0.80, ctx, alphabet.DNA)
    al.Costs = pals.Costs(*COSTS)
    hits = al.AlignTraps([pals.Trapezoid(*t) for t in T])
    assert [(h.Abpos, h.Bbpos, h.Aepos, h.Bepos, h.LowDiagonal, h.HighDiagonal, h.Score, h.Error) for h in hits] == H
    assert pals.Sum(hits) == (791, 664)
    rng = random.Random(17)
    problems = [_random_problem(rng, rng.randint(200, 1500), rng.randint(150, 900), rng.randint(1, 6)) for _ in range(12)]
    problems.append((b"ACGT" * 10, b"ACGT" * 10, []))                       # no trapezoids
    problems.append((b"", b"ACGT", [(3, 0, -2, 2)]))                        # empty target
    _check_batch(ctx, problems, 6, 20, 0.7, COSTS)
    _check_batch(ctx, problems[:6], 4, 10, 0.5, (3, 2, 1, 3, 8, 3.0))        # other costs, wider bands
    _check_batch(ctx, [(a, b, T)] * 3 + problems[:3], 6, 50, 0.80, COSTS)    # the golden problem inside a batch


def test_device_kernel_under_the_emulator(emu_ctx):
    _suite(emu_ctx)


@pytest.mark.gpu
def test_device_kernel_on_the_gpu(gpu_ctx):
    _suite(gpu_ctx)
    rng = random.Random(5)
    big = [_random_problem(rng, rng.randint(3000, 9000), rng.randint(2000, 6000), rng.randint(3, 12)) for _ in range(40)]
    _check_batch(gpu_ctx, big, 6, 50, 0.8, COSTS)
