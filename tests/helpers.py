"""Shared helpers for the parity tests: run one batch through a C-ABI library and through
the oracle and compare status, error positions and segments bit-exactly."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from biogo_b200 import _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libbiogoalign_emu.so")

KINDS = [(k, a) for a in (False, True) for k in (0, 1, 2)]
KIND_NAMES = {(0, False): "SW", (1, False): "NW", (2, False): "Fitted",
              (0, True): "SWAffine", (1, True): "NWAffine", (2, True): "FittedAffine"}


def pack(pairs, stride=1):
    """pairs: list of (ref bytes, qry bytes) -> packed buffer + offset/length arrays."""
    chunks, ro, rl, qo, ql = [], [], [], [], []
    off = 0
    for r, q in pairs:
        if stride == 2:
            r = bytes(b for ch in r for b in (ch, 40))
            q = bytes(b for ch in q for b in (ch, 40))
        chunks += [r, q]
        ro.append(off)
        rl.append(len(r) // stride)
        off += len(r)
        qo.append(off)
        ql.append(len(q) // stride)
        off += len(q)
    letters = np.frombuffer(b"".join(chunks) + b"\0", dtype=np.uint8)[:-1].copy()
    return (letters, np.array(ro, np.int64), np.array(rl, np.int32), np.array(qo, np.int64),
            np.array(ql, np.int32))


def run_lib(ctx, kind, affine, matrix, gap_open, index, alpha_len, packed, stride=1):
    la = np.array(matrix, dtype=np.int64).reshape(-1)
    ctx.set_scoring(kind, affine, la, len(matrix), alpha_len, gap_open, index)
    return ctx.align_packed(packed[0], stride, packed[1], packed[2], packed[3], packed[4])


def run_oracle(kind, affine, matrix, gap_open, index, packed, stride=1, nthreads=4):
    sc = O.Scoring(kind, affine, matrix, gap_open, index)
    return O.align_batch(sc, packed[0] if len(packed[0]) else np.zeros(1, np.uint8), stride, packed[1],
                         packed[2], packed[3], packed[4], nthreads)


def compare(res, orc, label=""):
    """Bit-exact comparison of a BatchResult with the oracle's batch output."""
    status, err_pos, seg_off, segs = orc
    n = len(status)
    assert res.n == n, label
    bad = []
    for p in range(n):
        if int(res.status[p]) != int(status[p]):
            bad.append((p, "status", int(res.status[p]), int(status[p])))
            continue
        if status[p] in (1, 2) and int(res.err_pos[p]) != int(err_pos[p]):
            bad.append((p, "err_pos", int(res.err_pos[p]), int(err_pos[p])))
            continue
        if status[p] == 0:
            want = segs[seg_off[p]:seg_off[p + 1]]
            got = res.pairs(p)
            if got.shape != want.shape or not np.array_equal(got.astype(np.int64), want):
                bad.append((p, "segs", got.tolist(), want.tolist()))
    assert not bad, f"{label}: {len(bad)} of {n} pairs differ, first: {bad[0]}"
