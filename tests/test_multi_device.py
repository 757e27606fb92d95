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


@pytest.mark.gpu
def test_all_visible_gpus():
    import torch

    n = torch.cuda.device_count()
    devices = tuple(range(n)) if n >= 2 else (0, 0)
    md = shard.MultiDevice(devices)
    _check(md, 4000, 4)
    cells = md.align_packed(*_uniform()).stats["cells_per_device"]
    assert max(cells) == min(cells)


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
