"""align.Format for a batch on the device (ba_format_batch) against the Format rows of
the reference's golden vectors (align/*_example_test.go) and the oracle's restatement of
align/align.go:146-170.  CPU tier: the kernels run under the SIMT emulator; GPU tier: on the B200."""
import random

import numpy as np
import pytest

import golden_vectors as gv
import helpers
import randcases
from biogo_b200 import _lib
from oracle import oracle as O

DNA_IDX = O.build_index("-acgt")


def _rows(off_a, off_b, row_a, row_b, p):
    return row_a[off_a[p]:off_a[p + 1]].tobytes().decode(), row_b[off_b[p]:off_b[p + 1]].tobytes().decode()


def _check_golden(ctx):
    for name, kind, affine, M, go, ref, qry, want, fmt in gv.GOLDEN:
        if fmt is None:
            continue
        pk = helpers.pack([(ref.encode(), qry.encode())])
        res = helpers.run_lib(ctx, kind, affine, M, go, DNA_IDX, 5, pk)
        for letters in (None, pk[0]):      # reuse the letters still on the device / upload them again
            out = ctx.format_packed(letters, 1, pk[1], pk[2], pk[3], pk[4], res)
            assert _rows(*out, 0) == fmt, name


def _check_random(ctx, npairs, stride, seed):
    rng = random.Random(seed)
    pairs = randcases.rand_pairs(rng, npairs, [0, 1, 7, 40, 130, 300], [0, 1, 9, 33, 100, 257], related=0.7, dirty=0.1)
    pk = helpers.pack(pairs, stride)
    M = gv.match(-1, 2, -1)
    for kind, affine, go in ((0, True, -3), (1, False, 0), (2, True, -3)):
        res = helpers.run_lib(ctx, kind, affine, M, go, DNA_IDX, 5, pk, stride)
        out = ctx.format_packed(pk[0], stride, pk[1], pk[2], pk[3], pk[4], res, ord("*"))
        for p, (r, q) in enumerate(pairs):
            segs = res.pairs(p)
            want = O.format_alignment(r, q, segs, "*") if len(segs) else ("", "")
            assert _rows(*out, p) == want, (kind, affine, p)


def test_format_golden_rows_emulated(emu_ctx):
    _check_golden(emu_ctx)


def test_format_random_batches_emulated(emu_ctx):
    _check_random(emu_ctx, 40, 1, 11)
    _check_random(emu_ctx, 25, 2, 12)


def test_format_rejects_bad_segments(emu_ctx):
    pk = helpers.pack([(b"ACGT", b"ACG")])
    res = helpers.run_lib(emu_ctx, 0, False, gv.match(-1, 2, -1), 0, DNA_IDX, 5, pk)
    res.segs = res.segs.copy()
    res.segs[0, 1] = 99   # a.end beyond the reference
    with pytest.raises(_lib.AlignError) as e:
        emu_ctx.format_packed(pk[0], 1, pk[1], pk[2], pk[3], pk[4], res)
    assert e.value.code == _lib.BA_ERR_ARG


@pytest.mark.gpu
def test_format_golden_rows_gpu(gpu_ctx):
    _check_golden(gpu_ctx)


@pytest.mark.gpu
def test_format_random_batches_gpu(gpu_ctx):
    _check_random(gpu_ctx, 3000, 1, 21)
    _check_random(gpu_ctx, 1500, 2, 22)
