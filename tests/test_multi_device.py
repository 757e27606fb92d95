"""One batch over several devices: deal pairs largest-first, pack each share, align per
device from its own host thread, gather by pair index.  CPU tier: two emulator contexts stand in
for two devices; GPU tier: every visible device (two contexts on one GPU when there is only one)."""
import random

import numpy as np
import pytest

import golden_vectors as gv
import helpers
import randcases
from biogo_b200 import shard
from oracle import oracle as O

DNA_IDX = O.build_index("-acgt")


def test_partition_balances_cells_and_keeps_every_pair():
    rng = np.random.default_rng(5)
    for n, d in ((1, 4), (7, 2), (1000, 8), (5000, 3)):
        rl = rng.integers(1, 5000, n)
        ql = rng.integers(1, 3000, n)
        shares = shard.partition(rl, ql, d)
        allids = np.sort(np.concatenate(shares))
        assert np.array_equal(allids, np.arange(n))
        assert all(np.all(np.diff(s) > 0) for s in shares if len(s) > 1)
        if n >= 1000:
            cells = np.array([(rl[s] * ql[s]).sum() for s in shares], dtype=np.float64)
            assert cells.max() / cells.mean() < 1.02
    uniform = shard.partition(np.full(800, 1000), np.full(800, 250), 8)
    assert [len(s) for s in uniform] == [100] * 8


def _check(md, npairs, seed, stride=1):
    rng = random.Random(seed)
    pairs = randcases.rand_pairs(rng, npairs, [0, 1, 20, 90, 150, 300], [0, 1, 16, 64, 100, 257], related=0.8, dirty=0.15)
    pk = helpers.pack(pairs, stride)
    M = gv.match(-1, 2, -1)
    for kind, affine, go in ((0, True, -4), (1, True, -4), (2, False, 0)):
        md.set_scoring(kind, affine, np.array(M, np.int64).reshape(-1), 5, 5, go, DNA_IDX)
        res = md.align_packed(pk[0], stride, pk[1], pk[2], pk[3], pk[4])
        helpers.compare(res, helpers.run_oracle(kind, affine, M, go, DNA_IDX, pk, stride), f"multi {kind} {affine}")
        assert res.stats["devices"] == len(md.ctxs) and res.stats["gpu_launches"] > 0


def test_two_emulated_devices(built):
    md = shard.MultiDevice((0, 0), helpers.EMU_LIB)
    _check(md, 37, 1)
    _check(md, 21, 2, stride=2)
    md1 = shard.MultiDevice((0,), helpers.EMU_LIB)
    _check(md1, 9, 3)


def _gpu_devices():
    """Every visible GPU.  BA_TEST_GPUS=N demands at least N physical devices and FAILS otherwise (a
    multi-GPU run must not silently degrade); without it a one-GPU box stands in with two contexts on
    device 0, which exercises the host logic but not two physical devices."""
    import os

    import torch

    n = torch.cuda.device_count()
    want = int(os.environ.get("BA_TEST_GPUS", "0"))
    assert n >= want, f"BA_TEST_GPUS={want} but only {n} GPU(s) are visible"
    return tuple(range(n)) if n >= 2 else (0, 0)


@pytest.mark.gpu
def test_all_visible_gpus():
    md = shard.MultiDevice(_gpu_devices())
    _check(md, 4000, 4)
    cells = md.align_packed(*_uniform()).stats["cells_per_device"]
    assert max(cells) == min(cells)


# ---- the same split UNDER the C ABI: ba_create_multi / ba_align_batch_multi ----------------------

def _check_multi(mc, npairs, seed, stride=1, kinds=((0, True, -4), (1, True, -4), (2, False, 0))):
    rng = random.Random(seed)
    pairs = randcases.rand_pairs(rng, npairs, [0, 1, 20, 90, 150, 300], [0, 1, 16, 64, 100, 257, 300], related=0.8,
                                 dirty=0.15)
    pk = helpers.pack(pairs, stride)
    M = gv.match(-1, 2, -1)
    for kind, affine, go in kinds:
        mc.set_scoring(kind, affine, np.array(M, np.int64).reshape(-1), 5, 5, go, DNA_IDX)
        res = mc.align_packed(pk[0], stride, pk[1], pk[2], pk[3], pk[4])
        helpers.compare(res, helpers.run_oracle(kind, affine, M, go, DNA_IDX, pk, stride), f"abi multi {kind} {affine}")
        assert res.stats["gpu_launches"] > 0 and res.stats["cells"] == sum(len(r) * len(q) for r, q in pairs)


def test_c_abi_multi_device_on_two_emulated_devices(built):
    from biogo_b200 import _lib

    mc = _lib.MultiContext((0, 1), helpers.EMU_LIB)
    assert mc.n_devices == 2
    _check_multi(mc, 41, 11, kinds=[(k, a, -3 if a else 0) for k, a in helpers.KINDS])
    _check_multi(mc, 23, 12, stride=2)
    # equal-shaped pairs: contiguous ranges, no gather; then the same batch dealt by buckets
    rng = random.Random(13)
    pairs = randcases.rand_pairs(rng, 33, [120], [60], related=0.9, dirty=0.0)
    pk = helpers.pack(pairs)
    M = gv.match(-1, 1, -1)
    mc.set_scoring(0, True, np.array(M, np.int64).reshape(-1), 5, 5, -5, DNA_IDX)
    orc = helpers.run_oracle(0, True, M, -5, DNA_IDX, pk)
    helpers.compare(mc.align_packed(pk[0], 1, pk[1], pk[2], pk[3], pk[4]), orc, "ranges")
    mc.set_option("contiguous_shares", 0)
    helpers.compare(mc.align_packed(pk[0], 1, pk[1], pk[2], pk[3], pk[4]), orc, "dealt")
    mc.set_option("contiguous_shares", 1)
    # pairs that overflow their slots go through the exact path on their device: host-resident segments
    # of a share are merged behind its device-resident ones
    mc.set_option("slot_cap", 2)
    try:
        _check_multi(mc, 30, 14, kinds=((0, True, -2),))
    finally:
        mc.set_option("slot_cap", 0)
    # a dealt share uploaded in several groups: a feeder thread gathers and enqueues them while the first
    # chunks are already being filled
    for ng, cp in ((3, 8), (7, 64)):
        mc.set_option("upload_groups", ng)
        mc.set_option("chunk_pairs", cp)
        try:
            _check_multi(mc, 90, 16 + ng, kinds=((0, True, -2), (2, False, 0)))
            _check_multi(mc, 40, 17 + ng, stride=2, kinds=((1, True, -3),))
        finally:
            mc.set_option("upload_groups", 0)
            mc.set_option("chunk_pairs", 8192)
    # one device, and an empty batch
    m1 = _lib.MultiContext((0,), helpers.EMU_LIB)
    _check_multi(m1, 9, 15)
    e = m1.align_packed(np.zeros(0, np.uint8), 1, np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int64),
                        np.zeros(0, np.int32))
    assert e.n == 0 and len(e.segs) == 0


def test_c_abi_multi_device_dealing_balances_cells(built):
    """The dealing itself, observed through the per-device cells of a batch of mixed lengths."""
    from biogo_b200 import _lib

    mc = _lib.MultiContext((0, 1, 0, 1), helpers.EMU_LIB)
    rng = random.Random(21)
    pairs = randcases.rand_pairs(rng, 400, [5, 10, 20, 40, 80], [5, 10, 20, 40, 80], related=0.5, dirty=0.0)
    pk = helpers.pack(pairs)
    M = gv.match(-1, 1, -1)
    mc.set_scoring(0, False, np.array(M, np.int64).reshape(-1), 5, 5, 0, DNA_IDX)
    res = mc.align_packed(pk[0], 1, pk[1], pk[2], pk[3], pk[4])
    helpers.compare(res, helpers.run_oracle(0, False, M, 0, DNA_IDX, pk), "four shares")


@pytest.mark.gpu
def test_c_abi_multi_device_on_all_visible_gpus():
    from biogo_b200 import _lib, synth

    mc = _lib.MultiContext(_gpu_devices())
    _check_multi(mc, 4000, 31)
    _check_multi(mc, 600, 32, stride=2, kinds=[(k, a, -3 if a else 0) for k, a in helpers.KINDS])
    # config 2 shape: contiguous shares; every pair equals the single-device result
    w = synth.config(2, 6000)
    la = np.array(w.matrix, np.int64).reshape(-1)
    mc.set_scoring(w.kind, True, la, 5, 5, w.gap_open, w.alpha.LetterIndex())
    multi = mc.align_packed(w.letters, 1, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    one = _lib.Context(0)
    one.set_scoring(w.kind, True, la, 5, 5, w.gap_open, w.alpha.LetterIndex())
    single = one.align_packed(w.letters, 1, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    assert np.array_equal(multi.status, single.status) and np.array_equal(multi.seg_count, single.seg_count)
    for p in range(w.n):
        assert np.array_equal(multi.pairs(p), single.pairs(p)), p
    # config 5 shape (mixed lengths, QLetters): dealt by length buckets
    w5 = synth.config(5, 600)
    la5 = np.array(w5.matrix, np.int64).reshape(-1)
    mc.set_scoring(w5.kind, True, la5, 5, 5, w5.gap_open, w5.alpha.LetterIndex())
    pk5 = (w5.letters, w5.ref_off, w5.ref_len, w5.qry_off, w5.qry_len)
    res5 = mc.align_packed(w5.letters, w5.stride, *pk5[1:])
    helpers.compare(res5, helpers.run_oracle(w5.kind, True, w5.matrix, w5.gap_open, w5.alpha.LetterIndex(), pk5,
                                             w5.stride, 16), "cfg5 multi")


def _uniform():
    rng = random.Random(9)
    pairs = randcases.rand_pairs(rng, 64, [200], [100], related=1.0, dirty=0.0)
    pk = helpers.pack(pairs)
    return pk[0], 1, pk[1], pk[2], pk[3], pk[4]


def test_two_host_threads_two_contexts(built):
    """bíogo's aligners are called from many goroutines (concurrent/processor.go:48-74): two host
    threads, one context each, interleaved batches -- the per-device compute token serialises the
    kernel phases, results stay per context."""
    import threading

    from biogo_b200 import _lib

    rng = random.Random(77)
    M = gv.match(-1, 1, -1)
    jobs = []
    for kind, affine, go in ((0, True, -5), (1, False, 0)):
        pairs = randcases.rand_pairs(rng, 25, [10, 90, 200], [10, 64, 257], related=0.8, dirty=0.1)
        pk = helpers.pack(pairs)
        jobs.append((kind, affine, go, pk, helpers.run_oracle(kind, affine, M, go, DNA_IDX, pk)))
    errs = []

    def worker(job):
        try:
            kind, affine, go, pk, orc = job
            ctx = _lib.Context(0, helpers.EMU_LIB)
            for _ in range(4):
                res = helpers.run_lib(ctx, kind, affine, M, go, DNA_IDX, 5, pk)
                helpers.compare(res, orc, f"thread kind {kind}")
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    th = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[0]
