"""ctypes binding of the C ABI declared in include/biogo_align.h.

The product library is ``biogo_b200/lib/libbiogoalign.so`` (built in-tree by
``__graft_entry__.build()``).  Loading fails loudly when it is missing and every call
fails loudly when no B200 is usable -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

BA_SW, BA_NW, BA_FITTED = 0, 1, 2

BA_OK = 0
BA_ERR_ARG, BA_ERR_CUDA, BA_ERR_MATRIX_SIZE, BA_ERR_RANGE, BA_ERR_NOMEM, BA_ERR_NO_SCORING = -1, -2, -3, -4, -5, -6
BA_ERR_FORMAT = -7
BA_PAIR_OK, BA_PAIR_ILLEGAL_REF, BA_PAIR_ILLEGAL_QRY, BA_PAIR_PANIC, BA_PAIR_NO_PATH = 0, 1, 2, 3, 4

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libbiogoalign.so")

# every symbol include/biogo_align.h declares
ABI_SYMBOLS = (
    "ba_create", "ba_destroy", "ba_last_error", "ba_set_scoring", "ba_align_batch",
    "ba_align_batch_device", "ba_free_result", "ba_alloc_host", "ba_free_host",
    "ba_set_option", "ba_abi_version", "ba_build_info", "ba_format_batch", "ba_free_formatted",
    "ba_fasta_parse", "ba_fastq_parse", "ba_seqfile_free", "ba_gather_pairs",
    "ba_create_multi", "ba_destroy_multi", "ba_multi_devices", "ba_multi_last_error", "ba_multi_set_scoring",
    "ba_multi_set_option", "ba_align_batch_multi", "ba_multi_free_result",
    "ba_matrix_count", "ba_matrix_name", "ba_matrix_get", "ba_set_scoring_named",
    "ba_pals_dp_batch", "ba_pals_free_result",
)


class BaResult(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("status", C.POINTER(C.c_int32)),
        ("err_pos", C.POINTER(C.c_int64)),
        ("seg_start", C.POINTER(C.c_int64)),
        ("seg_count", C.POINTER(C.c_int32)),
        ("segs", C.POINTER(C.c_int32)),
        ("n_segs", C.c_int64),
        ("ms_h2d", C.c_float),
        ("ms_kernels", C.c_float),
        ("ms_d2h", C.c_float),
        ("ms_fill", C.c_float),
        ("ms_trace", C.c_float),
        ("ms_host", C.c_float),
        ("cells", C.c_int64),
        ("gpu_launches", C.c_int64),
        ("fill_launches", C.c_int64),
        ("code_bytes", C.c_int64),
        ("pairs_fast16", C.c_int64),
        ("wavefront_launches", C.c_int64),
        ("warp_trace_launches", C.c_int64),
        ("waves", C.c_int64),
    ]


class BaFormatted(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("off_a", C.POINTER(C.c_int64)),
        ("off_b", C.POINTER(C.c_int64)),
        ("row_a", C.POINTER(C.c_uint8)),
        ("row_b", C.POINTER(C.c_uint8)),
        ("bytes_a", C.c_int64),
        ("bytes_b", C.c_int64),
        ("ms_kernels", C.c_float),
    ]


class BaSeqFile(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("stride", C.c_int),
        ("letters", C.POINTER(C.c_uint8)),
        ("letters_bytes", C.c_int64),
        ("seq_off", C.POINTER(C.c_int64)),
        ("seq_len", C.POINTER(C.c_int32)),
        ("text", C.POINTER(C.c_char)),
        ("text_bytes", C.c_int64),
        ("name_off", C.POINTER(C.c_int64)),
        ("desc_off", C.POINTER(C.c_int64)),
        ("name_len", C.POINTER(C.c_int32)),
        ("desc_len", C.POINTER(C.c_int32)),
        ("err_line", C.c_int64),
        ("err", C.c_char * 160),
    ]


class BaPalsCosts(C.Structure):
    _fields_ = [("max_igap", C.c_int32), ("diff_cost", C.c_int32), ("same_cost", C.c_int32), ("match_cost", C.c_int32),
                ("block_cost", C.c_int32), ("rmatch_cost", C.c_double)]


class BaPalsHit(C.Structure):
    _fields_ = [("abpos", C.c_int64), ("bbpos", C.c_int64), ("aepos", C.c_int64), ("bepos", C.c_int64),
                ("low_diagonal", C.c_int64), ("high_diagonal", C.c_int64), ("score", C.c_int64), ("error", C.c_double)]


class BaPalsResult(C.Structure):
    _fields_ = [("n", C.c_int64), ("hit_off", C.POINTER(C.c_int64)), ("hits", C.POINTER(BaPalsHit)),
                ("n_hits", C.c_int64), ("status", C.POINTER(C.c_int32)), ("ms_kernels", C.c_float)]


class AlignError(RuntimeError):
    """A ba_* call failed (CUDA error, range error ...)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libbiogoalign: status {code}: {msg}")
        self.code = code


def _bind(lib):
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.ba_create.argtypes = [i32, C.POINTER(vp)]
    lib.ba_create.restype = i32
    lib.ba_destroy.argtypes = [vp]
    lib.ba_destroy.restype = None
    lib.ba_last_error.argtypes = [vp]
    lib.ba_last_error.restype = C.c_char_p
    lib.ba_set_scoring.argtypes = [vp, i32, i32, vp, i32, i32, i64, vp]
    lib.ba_set_scoring.restype = i32
    lib.ba_align_batch.argtypes = [vp, vp, i32, vp, vp, vp, vp, i64, C.POINTER(BaResult)]
    lib.ba_align_batch.restype = i32
    lib.ba_align_batch_device.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, i64, vp, C.POINTER(BaResult)]
    lib.ba_align_batch_device.restype = i32
    lib.ba_free_result.argtypes = [vp, C.POINTER(BaResult)]
    lib.ba_free_result.restype = None
    lib.ba_alloc_host.argtypes = [C.c_size_t]
    lib.ba_alloc_host.restype = vp
    lib.ba_free_host.argtypes = [vp]
    lib.ba_free_host.restype = None
    lib.ba_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.ba_set_option.restype = i32
    lib.ba_abi_version.argtypes = []
    lib.ba_abi_version.restype = i32
    lib.ba_build_info.argtypes = []
    lib.ba_build_info.restype = C.c_char_p
    lib.ba_format_batch.argtypes = [vp, vp, i32, vp, vp, vp, vp, i64, vp, vp, vp, i64, C.c_uint8,
                                    C.POINTER(BaFormatted)]
    lib.ba_format_batch.restype = i32
    lib.ba_free_formatted.argtypes = [vp, C.POINTER(BaFormatted)]
    lib.ba_free_formatted.restype = None
    lib.ba_fasta_parse.argtypes = [vp, i64, i32, C.POINTER(BaSeqFile)]
    lib.ba_fasta_parse.restype = i32
    lib.ba_fastq_parse.argtypes = [vp, i64, i32, C.POINTER(BaSeqFile)]
    lib.ba_fastq_parse.restype = i32
    lib.ba_seqfile_free.argtypes = [C.POINTER(BaSeqFile)]
    lib.ba_seqfile_free.restype = None
    lib.ba_gather_pairs.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp]
    lib.ba_gather_pairs.restype = i32
    lib.ba_matrix_count.argtypes = []
    lib.ba_matrix_count.restype = i32
    lib.ba_matrix_name.argtypes = [i32]
    lib.ba_matrix_name.restype = C.c_char_p
    lib.ba_matrix_get.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(C.c_char_p), C.POINTER(C.POINTER(C.c_int16))]
    lib.ba_matrix_get.restype = i32
    lib.ba_set_scoring_named.argtypes = [vp, i32, i32, C.c_char_p, i64, i32, i64, vp]
    lib.ba_set_scoring_named.restype = i32
    lib.ba_pals_dp_batch.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, i32, i32, C.c_double,
                                     C.POINTER(BaPalsCosts), C.POINTER(BaPalsResult)]
    lib.ba_pals_dp_batch.restype = i32
    lib.ba_pals_free_result.argtypes = [vp, C.POINTER(BaPalsResult)]
    lib.ba_pals_free_result.restype = None
    lib.ba_create_multi.argtypes = [vp, i32, C.POINTER(vp)]
    lib.ba_create_multi.restype = i32
    lib.ba_destroy_multi.argtypes = [vp]
    lib.ba_destroy_multi.restype = None
    lib.ba_multi_devices.argtypes = [vp]
    lib.ba_multi_devices.restype = i32
    lib.ba_multi_last_error.argtypes = [vp]
    lib.ba_multi_last_error.restype = C.c_char_p
    lib.ba_multi_set_scoring.argtypes = [vp, i32, i32, vp, i32, i32, i64, vp]
    lib.ba_multi_set_scoring.restype = i32
    lib.ba_multi_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.ba_multi_set_option.restype = i32
    lib.ba_align_batch_multi.argtypes = [vp, vp, i32, vp, vp, vp, vp, i64, C.POINTER(BaResult)]
    lib.ba_align_batch_multi.restype = i32
    lib.ba_multi_free_result.argtypes = [vp, C.POINTER(BaResult)]
    lib.ba_multi_free_result.restype = None
    return lib


_libs: dict = {}


def load(path: str | None = None):
    """dlopen the C-ABI library (default: the in-tree CUDA build)."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path not in _libs:
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(biogo_b200 has no CPU fallback)"
            )
        _libs[path] = _bind(C.CDLL(path))
    return _libs[path]


class BatchResult:
    """numpy views of a ba_result, copied out so the pinned buffers can be recycled."""

    __slots__ = ("n", "status", "err_pos", "seg_start", "seg_count", "segs", "stats")

    def pairs(self, p: int):
        s, c = int(self.seg_start[p]), int(self.seg_count[p])
        return self.segs[s : s + c]


def _stats(res: BaResult) -> dict:
    return {k: (float(getattr(res, k)) if k.startswith("ms_") else int(getattr(res, k)))
            for k in ("ms_h2d", "ms_kernels", "ms_d2h", "ms_fill", "ms_trace", "ms_host", "cells", "gpu_launches", "fill_launches",
                      "code_bytes", "pairs_fast16", "wavefront_launches", "warp_trace_launches", "waves", "n_segs")}


def _as_array(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class Context:
    """A ba_ctx: one per (host thread, device)."""

    def __init__(self, device: int = 0, lib_path: str | None = None):
        self.lib = load(lib_path)
        h = C.c_void_p()
        rc = self.lib.ba_create(device, C.byref(h))
        self._h = h
        if rc != BA_OK:
            msg = self.lib.ba_last_error(h).decode() if h else "ba_create failed"
            if h:
                self.lib.ba_destroy(h)
                self._h = None
            raise AlignError(rc, msg)
        self.device = device
        self._scoring_key = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != BA_OK:
            raise AlignError(rc, self.lib.ba_last_error(self._h).decode())

    def set_option(self, key: str, value: int):
        self._check(self.lib.ba_set_option(self._h, key.encode(), int(value)))

    def set_scoring(self, kind: int, affine: bool, la, let: int, alpha_len: int, gap_open: int, index):
        la = np.ascontiguousarray(la, dtype=np.int64).reshape(-1)
        index = np.ascontiguousarray(index, dtype=np.int32)
        key = (kind, bool(affine), la.tobytes(), let, alpha_len, int(gap_open), index.tobytes())
        if key == self._scoring_key:
            return
        rc = self.lib.ba_set_scoring(self._h, kind, int(bool(affine)), la.ctypes.data, let, alpha_len,
                                     int(gap_open), index.ctypes.data)
        self._scoring_key = None
        self._check(rc)
        self._scoring_key = key

    def set_scoring_named(self, kind: int, affine: bool, name: str, gap: int, alpha_len: int, gap_open: int, index):
        """ba_set_scoring_named: one of the 78 align/matrix tables by name, `gap` in the gap row/column."""
        index = np.ascontiguousarray(index, dtype=np.int32)
        self._scoring_key = None
        self._check(self.lib.ba_set_scoring_named(self._h, kind, int(bool(affine)), name.encode(), int(gap), alpha_len,
                                                  int(gap_open), index.ctypes.data))

    def pals_dp_batch(self, letters, target_off, target_len, query_off, query_len, index, trap_off, traps, kmer: int,
                      min_hit_length: int, min_id: float, costs):
        """ba_pals_dp_batch.  costs = (MaxIGap, DiffCost, SameCost, MatchCost, BlockCost, RMatchCost).
        Returns (hits per problem as lists of 8-tuples, kernel ms)."""
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        to, qo = np.ascontiguousarray(target_off, np.int64), np.ascontiguousarray(query_off, np.int64)
        tl, ql = np.ascontiguousarray(target_len, np.int32), np.ascontiguousarray(query_len, np.int32)
        idx = np.ascontiguousarray(index, np.int32)
        troff = np.ascontiguousarray(trap_off, np.int64)
        tr = np.ascontiguousarray(np.asarray(traps, np.int32).reshape(-1))
        if len(tr) == 0:
            tr = np.zeros(4, np.int32)
        n = len(to)
        c = BaPalsCosts(*[int(x) for x in costs[:5]], float(costs[5]))
        res = BaPalsResult()
        rc = self.lib.ba_pals_dp_batch(self._h, letters.ctypes.data if len(letters) else None, to.ctypes.data,
                                       tl.ctypes.data, qo.ctypes.data, ql.ctypes.data, n, idx.ctypes.data,
                                       troff.ctypes.data, tr.ctypes.data, int(kmer), int(min_hit_length), float(min_id),
                                       C.byref(c), C.byref(res))
        self._check(rc)
        out = []
        for p in range(n):
            a, b = int(res.hit_off[p]), int(res.hit_off[p + 1])
            out.append([(h.abpos, h.bbpos, h.aepos, h.bepos, h.low_diagonal, h.high_diagonal, h.score, h.error)
                        for h in (res.hits[k] for k in range(a, b))])
        ms = float(res.ms_kernels)
        self.lib.ba_pals_free_result(self._h, C.byref(res))
        return out, ms

    def _collect(self, res: BaResult) -> BatchResult:
        out = BatchResult()
        n = int(res.n)
        out.n = n
        out.status = _as_array(res.status, n, np.int32)
        out.err_pos = _as_array(res.err_pos, n, np.int64)
        out.seg_start = _as_array(res.seg_start, n, np.int64)
        out.seg_count = _as_array(res.seg_count, n, np.int32)
        ns = int(res.n_segs)
        out.segs = _as_array(res.segs, 5 * ns, np.int32).reshape(ns, 5)
        out.stats = _stats(res)
        self.lib.ba_free_result(self._h, C.byref(res))
        return out

    def align_packed(self, letters, stride: int, ref_off, ref_len, qry_off, qry_len, collect: bool = True):
        """ba_align_batch on packed host buffers (numpy arrays or pinned ctypes memory)."""
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        ref_len = np.ascontiguousarray(ref_len, dtype=np.int32)
        qry_len = np.ascontiguousarray(qry_len, dtype=np.int32)
        n = len(ref_off)
        res = BaResult()
        rc = self.lib.ba_align_batch(self._h, letters.ctypes.data, stride, ref_off.ctypes.data,
                                     ref_len.ctypes.data, qry_off.ctypes.data, qry_len.ctypes.data, n,
                                     C.byref(res))
        self._check(rc)
        if collect:
            return self._collect(res)
        stats = _stats(res)
        self.lib.ba_free_result(self._h, C.byref(res))
        return stats

    def format_packed(self, letters, stride: int, ref_off, ref_len, qry_off, qry_len, result: "BatchResult",
                      gap: int = ord("-")):
        """ba_format_batch: align.Format for every pair of a batch.  ``letters=None`` formats the batch
        of this context's last align_packed call (its letters are still on the device).
        Returns (off_a, off_b, row_a, row_b): row p of side x is row_x[off_x[p]:off_x[p+1]]."""
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        ref_len = np.ascontiguousarray(ref_len, dtype=np.int32)
        qry_len = np.ascontiguousarray(qry_len, dtype=np.int32)
        seg_start = np.ascontiguousarray(result.seg_start, dtype=np.int64)
        seg_count = np.ascontiguousarray(result.seg_count, dtype=np.int32)
        segs = np.ascontiguousarray(result.segs, dtype=np.int32).reshape(-1)
        n = len(ref_off)
        lp = None
        if letters is not None:
            letters = np.ascontiguousarray(letters, dtype=np.uint8)
            lp = letters.ctypes.data
        f = BaFormatted()
        rc = self.lib.ba_format_batch(self._h, lp, stride, ref_off.ctypes.data, ref_len.ctypes.data,
                                      qry_off.ctypes.data, qry_len.ctypes.data, n, seg_start.ctypes.data,
                                      seg_count.ctypes.data, segs.ctypes.data if len(segs) else None, len(segs) // 5,
                                      int(gap), C.byref(f))
        self._check(rc)
        try:
            off_a = _as_array(f.off_a, n + 1, np.int64)
            off_b = _as_array(f.off_b, n + 1, np.int64)
            row_a = _as_array(f.row_a, int(f.bytes_a), np.uint8)
            row_b = _as_array(f.row_b, int(f.bytes_b), np.uint8)
        finally:
            self.lib.ba_free_formatted(self._h, C.byref(f))
        return off_a, off_b, row_a, row_b

    def align_device(self, d_letters: int, letters_bytes: int, stride: int, d_ref_off: int, d_ref_len: int,
                     d_qry_off: int, d_qry_len: int, h_ref_len, h_qry_len, stream: int = 0,
                     collect: bool = True):
        """ba_align_batch_device: inputs already resident in HBM (raw device pointers)."""
        h_ref_len = np.ascontiguousarray(h_ref_len, dtype=np.int32)
        h_qry_len = np.ascontiguousarray(h_qry_len, dtype=np.int32)
        n = len(h_ref_len)
        res = BaResult()
        rc = self.lib.ba_align_batch_device(self._h, d_letters, letters_bytes, stride, d_ref_off, d_ref_len,
                                            d_qry_off, d_qry_len, h_ref_len.ctypes.data,
                                            h_qry_len.ctypes.data, n, stream or None, C.byref(res))
        self._check(rc)
        if collect:
            return self._collect(res)
        stats = _stats(res)
        self.lib.ba_free_result(self._h, C.byref(res))
        return stats


class MultiContext:
    """A ba_multi: one batch over several GPUs of one box (ba_align_batch_multi)."""

    def __init__(self, devices=None, lib_path: str | None = None):
        self.lib = load(lib_path)
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.ba_create_multi(None, 0, C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.ba_create_multi(arr, len(devices), C.byref(h))
        self._h = h
        if rc != BA_OK:
            msg = self.lib.ba_multi_last_error(h).decode() if h else "ba_create_multi failed"
            if h:
                self.lib.ba_destroy_multi(h)
                self._h = None
            raise AlignError(rc, msg)
        self.n_devices = int(self.lib.ba_multi_devices(h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ba_destroy_multi(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != BA_OK:
            raise AlignError(rc, self.lib.ba_multi_last_error(self._h).decode())

    def set_option(self, key: str, value: int):
        self._check(self.lib.ba_multi_set_option(self._h, key.encode(), int(value)))

    def set_scoring(self, kind: int, affine: bool, la, let: int, alpha_len: int, gap_open: int, index):
        la = np.ascontiguousarray(la, dtype=np.int64).reshape(-1)
        index = np.ascontiguousarray(index, dtype=np.int32)
        self._check(self.lib.ba_multi_set_scoring(self._h, kind, int(bool(affine)), la.ctypes.data, let, alpha_len,
                                                  int(gap_open), index.ctypes.data))

    def align_packed(self, letters, stride: int, ref_off, ref_len, qry_off, qry_len, collect: bool = True):
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        ref_len = np.ascontiguousarray(ref_len, dtype=np.int32)
        qry_len = np.ascontiguousarray(qry_len, dtype=np.int32)
        n = len(ref_off)
        res = BaResult()
        self._check(self.lib.ba_align_batch_multi(self._h, letters.ctypes.data, stride, ref_off.ctypes.data,
                                                  ref_len.ctypes.data, qry_off.ctypes.data, qry_len.ctypes.data, n,
                                                  C.byref(res)))
        if collect:
            out = BatchResult()
            out.n = n
            out.status = _as_array(res.status, n, np.int32)
            out.err_pos = _as_array(res.err_pos, n, np.int64)
            out.seg_start = _as_array(res.seg_start, n, np.int64)
            out.seg_count = _as_array(res.seg_count, n, np.int32)
            ns = int(res.n_segs)
            out.segs = _as_array(res.segs, 5 * ns, np.int32).reshape(ns, 5)
            out.stats = _stats(res)
            self.lib.ba_multi_free_result(self._h, C.byref(res))
            return out
        stats = _stats(res)
        self.lib.ba_multi_free_result(self._h, C.byref(res))
        return stats


def matrix_names(lib_path: str | None = None):
    """Names of the scoring tables compiled into the library (align/matrix/matrices.go)."""
    lib = load(lib_path)
    return [lib.ba_matrix_name(k).decode() for k in range(lib.ba_matrix_count())]


def matrix_table(name: str, lib_path: str | None = None):
    """(letters in index order, table as a list of rows) of a named scoring table."""
    lib = load(lib_path)
    let, letters, tab = C.c_int(), C.c_char_p(), C.POINTER(C.c_int16)()
    if lib.ba_matrix_get(name.encode(), C.byref(let), C.byref(letters), C.byref(tab)) != BA_OK:
        raise KeyError(name)
    n = let.value
    flat = [int(tab[k]) for k in range(n * n)]
    return (letters.value or b"").decode(), [flat[r * n:(r + 1) * n] for r in range(n)]
