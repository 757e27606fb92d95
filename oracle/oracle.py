"""ctypes wrapper of the CPU oracle (oracle/align_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

ORC_SW, ORC_NW, ORC_FITTED = 0, 1, 2
ORC_OK, ORC_ILLEGAL_REF, ORC_ILLEGAL_QRY, ORC_PANIC, ORC_NO_PATH = 0, 1, 2, 3, 4


class PalsCosts(C.Structure):
    """dp.Costs (align/pals/dp/align.go:24-31)"""
    _fields_ = [("max_igap", C.c_int64), ("diff_cost", C.c_int64), ("same_cost", C.c_int64), ("match_cost", C.c_int64),
                ("block_cost", C.c_int64), ("rmatch_cost", C.c_double)]


class PalsHit(C.Structure):
    """dp.Hit (align/pals/dp/align.go:143-150)"""
    _fields_ = [("abpos", C.c_int64), ("bbpos", C.c_int64), ("aepos", C.c_int64), ("bepos", C.c_int64),
                ("low_diagonal", C.c_int64), ("high_diagonal", C.c_int64), ("score", C.c_int64), ("error", C.c_double)]


class _Scoring(C.Structure):
    _fields_ = [("kind", C.c_int), ("affine", C.c_int), ("let", C.c_int), ("la", C.c_void_p),
                ("gap_open", C.c_int64), ("index", C.c_void_p)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no GPU, no reference sources needed)."""
    srcs = [os.path.join(_HERE, "align_oracle.c"), os.path.join(_HERE, "pals_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        # -ffp-contract=off: the float64 expressions of align/pals/dp are evaluated operation by operation, as Go does on amd64
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-pthread", "-ffp-contract=off", "-shared", "-o", _SO,
                               *srcs])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        l.orc_align.argtypes = [C.POINTER(_Scoring), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        l.orc_align.restype = C.c_int
        l.orc_align_batch.argtypes = [C.POINTER(_Scoring), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_void_p)]
        l.orc_align_batch.restype = C.c_int
        l.orc_free.argtypes = [C.c_void_p]
        l.orc_free.restype = None
        l.orc_build_index.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        l.orc_build_index.restype = None
        l.orc_format_pairs.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]
        l.orc_format_pairs.restype = C.c_int
        l.orc_format_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_uint8,
                                           C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.orc_format_alignment.restype = C.c_int
        l.pals_align_traps.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int64, C.c_int64, C.c_double, C.POINTER(PalsCosts), C.c_int,
                                       C.POINTER(C.c_void_p)]
        l.pals_align_traps.restype = C.c_int64
        l.pals_free.argtypes = [C.c_void_p]
        l.pals_free.restype = None
        l.pals_debruijn.argtypes = [C.c_int, C.c_int, C.c_void_p]
        l.pals_debruijn.restype = C.c_int64
        _lib = l
    return _lib


def build_index(letters: str, case_sensitive: bool = False) -> np.ndarray:
    idx = np.empty(256, dtype=np.int32)
    lib().orc_build_index(letters.encode(), int(case_sensitive), idx.ctypes.data)
    return idx


class Scoring:
    def __init__(self, kind: int, affine: bool, matrix, gap_open: int, index):
        self.la = np.ascontiguousarray(np.array(matrix, dtype=np.int64).reshape(-1))
        self.let = len(matrix)
        self.index = np.ascontiguousarray(index, dtype=np.int32)
        self.c = _Scoring(kind, int(affine), self.let, self.la.ctypes.data, int(gap_open), self.index.ctypes.data)


def align(sc: Scoring, ref: bytes, qry: bytes, stride: int = 1):
    """-> (status, segments as list of 5-tuples, err_pos)"""
    segs = C.c_void_p()
    n = C.c_int64()
    ep = C.c_int64()
    r = np.frombuffer(ref, dtype=np.uint8) if len(ref) else np.zeros(1, np.uint8)
    q = np.frombuffer(qry, dtype=np.uint8) if len(qry) else np.zeros(1, np.uint8)
    st = lib().orc_align(C.byref(sc.c), r.ctypes.data, len(ref) // stride, q.ctypes.data, len(qry) // stride,
                         stride, C.byref(segs), C.byref(n), C.byref(ep))
    out = []
    if st == ORC_OK and n.value:
        a = np.ctypeslib.as_array(C.cast(segs, C.POINTER(C.c_int64)), shape=(n.value, 5)).copy()
        out = [tuple(int(x) for x in row) for row in a]
    if segs:
        lib().orc_free(segs)
    return st, out, int(ep.value)


def align_batch(sc: Scoring, letters, stride, ref_off, ref_len, qry_off, qry_len, nthreads: int = 1):
    """-> (status[n], err_pos[n], seg_off[n+1], segs[total,5]) as numpy arrays"""
    letters = np.ascontiguousarray(letters, dtype=np.uint8)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
    qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
    ref_len = np.ascontiguousarray(ref_len, dtype=np.int32)
    qry_len = np.ascontiguousarray(qry_len, dtype=np.int32)
    n = len(ref_off)
    status = np.zeros(n, dtype=np.int32)
    err_pos = np.zeros(n, dtype=np.int64)
    seg_off = np.zeros(n + 1, dtype=np.int64)
    segs = C.c_void_p()
    rc = lib().orc_align_batch(C.byref(sc.c), letters.ctypes.data, stride, ref_off.ctypes.data,
                               ref_len.ctypes.data, qry_off.ctypes.data, qry_len.ctypes.data, n, nthreads,
                               status.ctypes.data, err_pos.ctypes.data, seg_off.ctypes.data, C.byref(segs))
    if rc != 0:
        raise MemoryError("oracle batch failed")
    total = int(seg_off[n])
    if total:
        a = np.ctypeslib.as_array(C.cast(segs, C.POINTER(C.c_int64)), shape=(total, 5)).copy()
    else:
        a = np.zeros((0, 5), dtype=np.int64)
    lib().orc_free(segs)
    return status, err_pos, seg_off, a


def format_pairs(segs) -> str:
    a = np.ascontiguousarray(np.array(segs, dtype=np.int64).reshape(-1, 5))
    buf = C.create_string_buffer(64 * (len(a) + 2))
    n = lib().orc_format_pairs(a.ctypes.data, len(a), buf, len(buf))
    if n < 0:
        raise ValueError("format buffer too small")
    return buf.value.decode()


def format_alignment(ref: bytes, qry: bytes, segs, gap: str = "-", stride: int = 1):
    a = np.ascontiguousarray(np.array(segs, dtype=np.int64).reshape(-1, 5))
    cap = len(ref) + len(qry) + 8
    oa = np.zeros(cap, np.uint8)
    ob = np.zeros(cap, np.uint8)
    lens = np.zeros(2, np.int64)
    r = np.frombuffer(ref, dtype=np.uint8) if len(ref) else np.zeros(1, np.uint8)
    q = np.frombuffer(qry, dtype=np.uint8) if len(qry) else np.zeros(1, np.uint8)
    rc = lib().orc_format_alignment(r.ctypes.data, q.ctypes.data, stride, a.ctypes.data, len(a), ord(gap),
                                    oa.ctypes.data, ob.ctypes.data, cap, lens.ctypes.data)
    if rc != 0:
        raise ValueError("format_alignment failed")
    return oa[: lens[0]].tobytes().decode(), ob[: lens[1]].tobytes().decode()


# ---- align/pals/dp (oracle/pals_oracle.c) ---------------------------------------------------------

def pals_align_traps(target: bytes, query: bytes, index, traps, kmer: int, min_hit_length: int, min_id: float, costs,
                     dedup: bool = True):
    """dp.Aligner.AlignTraps restated.  traps: [(top, bottom, left, right)]; costs: (MaxIGap, DiffCost, SameCost,
    MatchCost, BlockCost, RMatchCost).  Returns [(Abpos, Bbpos, Aepos, Bepos, LowDiagonal, HighDiagonal, Score, Error)]."""
    l = lib()
    t = np.frombuffer(target, dtype=np.uint8) if len(target) else np.zeros(1, np.uint8)
    q = np.frombuffer(query, dtype=np.uint8) if len(query) else np.zeros(1, np.uint8)
    idx = np.ascontiguousarray(index, dtype=np.int32)
    tr = np.ascontiguousarray(np.array(traps, dtype=np.int64).reshape(-1, 4))
    c = PalsCosts(*[int(x) for x in costs[:5]], float(costs[5]))
    out = C.c_void_p()
    n = l.pals_align_traps(t.ctypes.data, len(target), q.ctypes.data, len(query), idx.ctypes.data, tr.ctypes.data, len(tr),
                           kmer, min_hit_length, float(min_id), C.byref(c), int(dedup), C.byref(out))
    hits = []
    if n > 0:
        arr = C.cast(out, C.POINTER(PalsHit))
        for k in range(n):
            h = arr[k]
            hits.append((h.abpos, h.bbpos, h.aepos, h.bepos, h.low_diagonal, h.high_diagonal, h.score, h.error))
    if out:
        l.pals_free(out)
    return hits


def debruijn(k: int, n: int) -> bytes:
    """util.DeBruijn(k, n) (util/debruijn.go:8-40): letter values 0..k-1."""
    buf = np.zeros(max(k ** n, n, 1), np.uint8)
    m = lib().pals_debruijn(k, n, buf.ctypes.data)
    return buf[:m].tobytes()
