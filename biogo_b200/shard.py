"""One batch over several GPUs of a box.

Pairs are independent, so there is no collective: pairs are dealt to devices largest-first so that
every device receives the same number of DP cells within about a percent, each device's share is
packed contiguously (``ba_gather_pairs``) and aligned from its own host thread on its own
``ba_ctx``, and the results are scattered back by original pair index on the host.
"""
from __future__ import annotations

import threading

import numpy as np

from . import _lib


def partition(ref_len, qry_len, n_dev: int):
    """Deal pair ids to ``n_dev`` devices: sorted by cells (descending), dealt in snake order
    (0..n-1, n-1..0, ...), which is LPT for near-uniform sizes and keeps the sums within ~1 %.
    Returns a list of int64 id arrays (each ascending, so a share keeps the batch order)."""
    ref_len = np.asarray(ref_len, np.int64)
    qry_len = np.asarray(qry_len, np.int64)
    n = len(ref_len)
    order = np.argsort(-(ref_len * qry_len), kind="stable")
    k = np.arange(n) % (2 * n_dev)
    dev = np.where(k < n_dev, k, 2 * n_dev - 1 - k)
    return [np.sort(order[dev == d]).astype(np.int64) for d in range(n_dev)]


class MultiDevice:
    """One ``_lib.Context`` per device; ``align_packed`` spreads a batch over all of them."""

    def __init__(self, devices=(0,), lib_path: str | None = None):
        self.ctxs = [_lib.Context(d, lib_path) for d in devices]

    def set_scoring(self, *args):
        for c in self.ctxs:
            c.set_scoring(*args)

    def align_packed(self, letters, stride: int, ref_off, ref_len, qry_off, qry_len) -> _lib.BatchResult:
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        ref_len = np.ascontiguousarray(ref_len, dtype=np.int32)
        qry_len = np.ascontiguousarray(qry_len, dtype=np.int32)
        n = len(ref_off)
        shares = partition(ref_len, qry_len, len(self.ctxs))
        parts, errs = [None] * len(self.ctxs), []

        def work(d):
            try:
                ids, ctx = shares[d], self.ctxs[d]
                k = len(ids)
                nbytes = int(stride * (ref_len[ids].astype(np.int64).sum() + qry_len[ids].astype(np.int64).sum()))
                dst = np.empty(max(nbytes, 1), np.uint8)
                nro, nqo = np.empty(max(k, 1), np.int64), np.empty(max(k, 1), np.int64)
                rc = ctx.lib.ba_gather_pairs(letters.ctypes.data, stride, ref_off.ctypes.data, ref_len.ctypes.data,
                                             qry_off.ctypes.data, qry_len.ctypes.data, ids.ctypes.data, k,
                                             dst.ctypes.data, nbytes, nro.ctypes.data, nqo.ctypes.data)
                if rc != _lib.BA_OK:
                    raise _lib.AlignError(rc, "ba_gather_pairs failed")
                parts[d] = ctx.align_packed(dst[:nbytes], stride, nro[:k], ref_len[ids], nqo[:k], qry_len[ids])
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)

        threads = [threading.Thread(target=work, args=(d,)) for d in range(len(self.ctxs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        # ---- gather by original pair index
        out = _lib.BatchResult()
        out.n = n
        out.status = np.zeros(n, np.int32)
        out.err_pos = np.zeros(n, np.int64)
        out.seg_count = np.zeros(n, np.int32)
        for d, ids in enumerate(shares):
            out.status[ids] = parts[d].status
            out.err_pos[ids] = parts[d].err_pos
            out.seg_count[ids] = parts[d].seg_count
        out.seg_start = np.zeros(n, np.int64)
        if n:
            out.seg_start[1:] = np.cumsum(out.seg_count.astype(np.int64))[:-1]
        total = int(out.seg_count.astype(np.int64).sum())
        out.segs = np.zeros((total, 5), np.int32)
        for d, ids in enumerate(shares):
            p = parts[d]
            cnt = p.seg_count.astype(np.int64)
            if cnt.sum() == 0:
                continue
            # segment rows of the share, in share order, and where each goes in the merged array
            src = np.repeat(p.seg_start, cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
            dst = np.repeat(out.seg_start[ids], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
            out.segs[dst] = p.segs[src]
        out.stats = {"devices": len(self.ctxs), "cells_per_device": [int((ref_len[i].astype(np.int64) *
                                                                          qry_len[i]).sum()) for i in shares]}
        for key in ("gpu_launches", "pairs_fast16", "cells", "n_segs"):
            out.stats[key] = sum(p.stats[key] for p in parts)
        return out
