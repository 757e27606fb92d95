"""Kernel LOGIC parity on the CPU: the unmodified kernel sources (biogo_b200/csrc/*.cuh) and
host scheduler (ba_api.cu) are compiled with g++ against tests/emu/cuda_emu.h and run through
the same C ABI, then compared bit-exactly with the oracle.  This is test infrastructure for the
container that has no GPU; the `-m gpu` tests repeat the comparison on the real sm_100a build."""
import random

import numpy as np
import pytest

import golden_vectors as gv
import helpers
import randcases
from biogo_b200 import matrix
from helpers import O

DNA_IDX = O.build_index("-acgt")
PROT_IDX = O.build_index("-abcdefghijklmnpqrstvwxyz*")


@pytest.mark.parametrize("vec", gv.GOLDEN, ids=[v[0] for v in gv.GOLDEN])
def test_golden_through_emulated_kernels(emu_ctx, vec):
    name, kind, affine, M, go, ref, qry, want, _ = vec
    pk = helpers.pack([(ref.encode(), qry.encode())])
    res = helpers.run_lib(emu_ctx, kind, affine, M, go, DNA_IDX, 5, pk)
    assert int(res.status[0]) == 0
    assert O.format_pairs(res.pairs(0)) == want


@pytest.mark.parametrize("vec", gv.DERIVED, ids=[v[0] for v in gv.DERIVED])
def test_derived_through_emulated_kernels(emu_ctx, vec):
    name, kind, affine, alpha, M, go, ref, qry, want = vec
    if alpha == "protein":
        M, idx, alen = matrix.with_gap(matrix.BLOSUM62, -1), PROT_IDX, 26
    else:
        idx, alen = DNA_IDX, 5
    res = helpers.run_lib(emu_ctx, kind, affine, M, go, idx, alen, helpers.pack([(ref.encode(), qry.encode())]))
    assert O.format_pairs(res.pairs(0)) == want


@pytest.mark.parametrize("kind,affine", helpers.KINDS, ids=[helpers.KIND_NAMES[k] for k in helpers.KINDS])
def test_random_small_batches(emu_ctx, kind, affine):
    rng = random.Random(100 + kind + 10 * affine)
    for trial in range(12):
        M = randcases.rand_matrix(rng)
        go = rng.choice([0, -1, -2, -5, 1])
        stride = rng.choice([1, 2])
        pairs = randcases.rand_pairs(rng, 24, [0, 1, 2, 3, 5, 8, 13, 30, 60], [0, 1, 2, 3, 5, 9, 17, 33, 40, 70, 100])
        pk = helpers.pack(pairs, stride)
        res = helpers.run_lib(emu_ctx, kind, affine, M, go, DNA_IDX, 5, pk, stride)
        helpers.compare(res, helpers.run_oracle(kind, affine, M, go, DNA_IDX, pk, stride),
                        f"{helpers.KIND_NAMES[(kind, affine)]} trial {trial} M={M} go={go} stride={stride}")


@pytest.mark.parametrize("kind,affine", helpers.KINDS, ids=[helpers.KIND_NAMES[k] for k in helpers.KINDS])
def test_multi_pass_and_wide_lanes(emu_ctx, kind, affine):
    """queries of 129..600 letters: 5..8 columns per lane and more than one 256-column pass."""
    rng = random.Random(500 + kind + 10 * affine)
    M = gv.match(-1, 2, -1) if not affine else gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 6, [40, 90, 300], [129, 200, 256, 257, 300, 600], related=0.7, dirty=0.0)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(emu_ctx, kind, affine, M, -3, DNA_IDX, 5, pk)
    helpers.compare(res, helpers.run_oracle(kind, affine, M, -3, DNA_IDX, pk),
                    f"{helpers.KIND_NAMES[(kind, affine)]} multipass")


def test_protein_blosum62(emu_ctx):
    rng = random.Random(9)
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    M = matrix.with_gap(matrix.BLOSUM62, -1)
    pairs = randcases.rand_pairs(rng, 16, [20, 45, 80], [20, 45, 80], related=0.8, dirty=0.15, alpha=aa,
                                 dirty_alpha=aa + b"BZX*J-acdu")
    pk = helpers.pack(pairs)
    for kind, affine in helpers.KINDS:
        res = helpers.run_lib(emu_ctx, kind, affine, M, -10, PROT_IDX, 26, pk)
        helpers.compare(res, helpers.run_oracle(kind, affine, M, -10, PROT_IDX, pk), f"protein {kind} {affine}")


def test_waves_when_the_arena_is_small(emu_ctx):
    """A tiny traceback arena forces several waves; results must not change."""
    rng = random.Random(3)
    pairs = randcases.rand_pairs(rng, 40, [30, 60], [20, 50, 90], dirty=0.0)
    pk = helpers.pack(pairs)
    M = gv.match(-1, 1, -1)
    emu_ctx.set_option("arena_bytes", 20000)
    try:
        res = helpers.run_lib(emu_ctx, 0, True, M, -5, DNA_IDX, 5, pk)
    finally:
        emu_ctx.set_option("arena_bytes", 0)
    helpers.compare(res, helpers.run_oracle(0, True, M, -5, DNA_IDX, pk), "waves")


def test_empty_batch_and_range_error(emu_ctx):
    from biogo_b200 import _lib

    M = gv.match(-1, 1, -1)
    res = helpers.run_lib(emu_ctx, 0, True, M, -5, DNA_IDX, 5, helpers.pack([]))
    assert res.n == 0 and len(res.segs) == 0
    big = [[1 << 22] * 5 for _ in range(5)]
    with pytest.raises(_lib.AlignError) as e:
        emu_ctx.set_scoring(0, False, np.array(big, np.int64).reshape(-1), 5, 5, 0, DNA_IDX)
    assert e.value.code == _lib.BA_ERR_RANGE
    with pytest.raises(_lib.AlignError) as e:   # SW/NW reject a matrix smaller than the alphabet
        emu_ctx.set_scoring(0, False, np.zeros(16, np.int64), 4, 5, 0, DNA_IDX)
    assert e.value.code == _lib.BA_ERR_MATRIX_SIZE
    emu_ctx.set_scoring(2, False, np.zeros(16, np.int64), 4, 5, 0, DNA_IDX)  # Fitted has no such check


def _fast_batch(rng, count, lens_r, lens_q):
    pairs = randcases.rand_pairs(rng, count, lens_r, lens_q, related=0.85, dirty=0.0)
    return [(r, q) for r, q in pairs]


@pytest.mark.parametrize("params", [(-1, 1, -1, -5), (-1, 2, -1, -3), (-2, 5, -4, -10), (0, 1, -1, -2)],
                         ids=["m1", "m2", "nuc4like", "gap0"])
def test_fast16_path_swaffine(emu_ctx, params):
    """The packed 16-bit DPX fill + checkpointed tile traceback (ba_fill16 / ba_trace16)."""
    gap, ma, mi, go = params
    rng = random.Random(hash(params) & 0xffff)
    M = gv.match(gap, ma, mi)
    pairs = _fast_batch(rng, 40, [1, 2, 7, 15, 16, 17, 31, 33, 48, 64, 90, 130],
                        [1, 2, 5, 8, 9, 16, 31, 32, 33, 40, 64, 65, 95, 100, 160])
    pk = helpers.pack(pairs)
    res = helpers.run_lib(emu_ctx, 0, True, M, go, DNA_IDX, 5, pk)
    assert res.stats["pairs_fast16"] == len(pairs)
    helpers.compare(res, helpers.run_oracle(0, True, M, go, DNA_IDX, pk), f"fast16 {params}")


def test_fast16_multi_pass_mixed_and_overflow(emu_ctx):
    rng = random.Random(77)
    M = gv.match(-1, 1, -1)
    pairs = _fast_batch(rng, 10, [60, 200, 300], [257, 300, 520])                 # > 256 columns: several passes
    pairs += randcases.rand_pairs(rng, 6, [20, 50], [20, 50], dirty=1.0)           # N: illegal-letter errors
    pairs += randcases.rand_pairs(rng, 8, [20, 50], [20, 50], dirty=1.0, dirty_alpha=b"ACGT-acgt")  # gap letters -> exact path
    pairs += _fast_batch(rng, 12, [30, 80], [30, 80])
    pk = helpers.pack(pairs, 2)
    res = helpers.run_lib(emu_ctx, 0, True, M, -5, DNA_IDX, 5, pk, 2)
    assert 0 < res.stats["pairs_fast16"] < len(pairs)
    orc = helpers.run_oracle(0, True, M, -5, DNA_IDX, pk, 2)
    helpers.compare(res, orc, "fast16 mixed")
    emu_ctx.set_option("slot_cap", 2)   # almost every pair overflows its segment slots -> exact path
    try:
        res2 = helpers.run_lib(emu_ctx, 0, True, M, -5, DNA_IDX, 5, pk, 2)
    finally:
        emu_ctx.set_option("slot_cap", 0)
    helpers.compare(res2, orc, "fast16 overflow fallback")
    emu_ctx.set_option("force_path", 1)
    try:
        res3 = helpers.run_lib(emu_ctx, 0, True, M, -5, DNA_IDX, 5, pk, 2)
    finally:
        emu_ctx.set_option("force_path", 0)
    assert res3.stats["pairs_fast16"] == 0
    helpers.compare(res3, orc, "forced 32-bit path")


def test_uniform_batch_closed_form_plan_and_waves(emu_ctx):
    """All pairs of one shape take the planner's closed-form path; a small arena forces waves."""
    rng = random.Random(21)
    pairs = randcases.rand_pairs(rng, 30, [70], [40], related=0.8, dirty=0.0)
    pk = helpers.pack(pairs)
    M = gv.match(-1, 1, -1)
    for kind, affine in ((0, True), (1, True), (2, False)):
        emu_ctx.set_option("arena_bytes", 60000)
        try:
            res = helpers.run_lib(emu_ctx, kind, affine, M, -5, DNA_IDX, 5, pk)
        finally:
            emu_ctx.set_option("arena_bytes", 0)
        helpers.compare(res, helpers.run_oracle(kind, affine, M, -5, DNA_IDX, pk), f"uniform waves {kind} {affine}")


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SWAffine", "NWAffine", "FittedAffine"])
def test_fast16_all_affine_kinds_dna_and_protein(emu_ctx, kind):
    """The packed 16-bit path for NWAffine / FittedAffine (borders, -inf, start cells) and for
    a generic alphabet (Protein/BLOSUM62 through the shared-memory score tables)."""
    rng = random.Random(900 + kind)
    M = gv.match(-1, 2, -1)
    pairs = randcases.rand_pairs(rng, 30, [1, 5, 17, 40, 90, 150], [1, 4, 16, 33, 64, 100, 257, 300], related=0.85,
                                 dirty=0.0)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(emu_ctx, kind, True, M, -4, DNA_IDX, 5, pk)
    assert res.stats["pairs_fast16"] >= len(pairs) - 3
    helpers.compare(res, helpers.run_oracle(kind, True, M, -4, DNA_IDX, pk), f"fast16 dna kind {kind}")
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    P = matrix.with_gap(matrix.BLOSUM62, -1)
    ppairs = randcases.rand_pairs(rng, 24, [3, 20, 45, 80, 300], [2, 20, 45, 80, 300], related=0.8, dirty=0.15,
                                  alpha=aa, dirty_alpha=aa + b"BZX*J-acdu")
    ppk = helpers.pack(ppairs, 2)
    res = helpers.run_lib(emu_ctx, kind, True, P, -12, PROT_IDX, 26, ppk, 2)
    assert res.stats["pairs_fast16"] > 0
    helpers.compare(res, helpers.run_oracle(kind, True, P, -12, PROT_IDX, ppk, 2), f"fast16 protein kind {kind}")


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SWAffine", "NWAffine", "FittedAffine"])
def test_fast16_wavefront_over_passes(emu_ctx, kind):
    """Long queries, few pairs: the passes of a pair run as a wavefront on different warps
    (LONG mode of ba_fill16, boundary column handed over through global memory)."""
    rng = random.Random(4100 + kind)
    M = gv.match(-1, 2, -1)
    # 4..7 passes (column classes 7 and 8), odd pair count, very different reference lengths
    pairs = randcases.rand_pairs(rng, 7, [9, 40, 130, 260], [800, 1025, 1300, 1700], related=0.85, dirty=0.0)
    pk = helpers.pack(pairs)
    emu_ctx.set_option("long_mode", 2)
    try:
        res = helpers.run_lib(emu_ctx, kind, True, M, -4, DNA_IDX, 5, pk)
    finally:
        emu_ctx.set_option("long_mode", 1)
    assert res.stats["wavefront_launches"] >= 1
    assert res.stats["pairs_fast16"] >= len(pairs) - 1
    helpers.compare(res, helpers.run_oracle(kind, True, M, -4, DNA_IDX, pk), f"wavefront kind {kind}")
    if kind == 0:
        aa = b"ACDEFGHIKLMNPQRSTVWY"
        P = matrix.with_gap(matrix.BLOSUM62, -1)
        ppairs = randcases.rand_pairs(rng, 3, [50, 120], [1030, 1290], related=0.8, dirty=0.0, alpha=aa)
        ppk = helpers.pack(ppairs)
        emu_ctx.set_option("long_mode", 2)
        try:
            res = helpers.run_lib(emu_ctx, kind, True, P, -12, PROT_IDX, 26, ppk)
        finally:
            emu_ctx.set_option("long_mode", 1)
        assert res.stats["wavefront_launches"] >= 1
        helpers.compare(res, helpers.run_oracle(kind, True, P, -12, PROT_IDX, ppk), "wavefront protein")


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SWAffine", "NWAffine", "FittedAffine"])
def test_fast16_warp_per_pair_traceback(emu_ctx, kind):
    """ba_trace16w_kernel: lanes split the start-cell search and rebuild a window of tiles."""
    rng = random.Random(4300 + kind)
    M = gv.match(-1, 2, -1)
    pairs = randcases.rand_pairs(rng, 18, [1, 5, 17, 40, 90, 150, 260], [1, 4, 16, 33, 64, 100, 257, 300, 600],
                                 related=0.85, dirty=0.0)
    pairs += randcases.rand_pairs(rng, 4, [30, 200], [30, 200], related=0.0, dirty=0.0)   # unrelated: gappy paths
    pk = helpers.pack(pairs)
    orc = helpers.run_oracle(kind, True, M, -4, DNA_IDX, pk)
    # a warp per pair (4 x 8 tile window), a CTA per pair (diagonal band of tiles), warps for the head of
    # the work list and threads for the rest
    for mode in (2, 3, 4, 5):
        emu_ctx.set_option("trace_mode", mode)
        try:
            res = helpers.run_lib(emu_ctx, kind, True, M, -4, DNA_IDX, 5, pk)
            assert (res.stats["warp_trace_launches"] >= 1) == (mode != 5)   # 5: persistent thread-per-pair kernel
            helpers.compare(res, orc, f"team trace mode {mode} kind {kind}")
            emu_ctx.set_option("slot_cap", 2)
            res2 = helpers.run_lib(emu_ctx, kind, True, M, -4, DNA_IDX, 5, pk)
            helpers.compare(res2, orc, f"team trace overflow mode {mode} kind {kind}")
        finally:
            emu_ctx.set_option("slot_cap", 0)
            emu_ctx.set_option("trace_mode", 1)
    emu_ctx.set_option("trace_mode", 0)
    try:
        res0 = helpers.run_lib(emu_ctx, kind, True, M, -4, DNA_IDX, 5, pk)
    finally:
        emu_ctx.set_option("trace_mode", 1)
    assert res0.stats["warp_trace_launches"] == 0
    helpers.compare(res0, orc, f"thread trace kind {kind}")


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SW", "NW", "Fitted"])
def test_fast16_linear_kinds(emu_ctx, kind):
    """Linear-gap aligners on the packed path (LIN mode of ba_fill16 + linear tiles of ba_trace16):
    DNA through PRMT profiles, Protein through the shared-memory tables, several passes, both
    traceback kernels and the wavefront over passes."""
    rng = random.Random(4700 + kind)
    M = gv.match(-1, 2, -1)
    pairs = randcases.rand_pairs(rng, 36, [1, 2, 5, 17, 40, 90, 150, 260], [1, 3, 4, 16, 33, 64, 100, 257, 300, 520],
                                 related=0.85, dirty=0.0)
    pairs += randcases.rand_pairs(rng, 6, [30, 120], [30, 120], related=0.0, dirty=0.0)
    pk = helpers.pack(pairs)
    orc = helpers.run_oracle(kind, False, M, 0, DNA_IDX, pk)
    for mode in (0, 2, 3, 5):
        emu_ctx.set_option("trace_mode", mode)
        try:
            res = helpers.run_lib(emu_ctx, kind, False, M, 0, DNA_IDX, 5, pk)
        finally:
            emu_ctx.set_option("trace_mode", 1)
        assert res.stats["pairs_fast16"] >= len(pairs) // 2   # segment-slot overflow re-runs on the exact path
        helpers.compare(res, orc, f"fast16 linear kind {kind} trace_mode {mode}")
    # ties everywhere: match 1 / mismatch -1 / gap -1 makes diag, up and left collide often
    T = gv.match(-1, 1, -1)
    tp = randcases.rand_pairs(rng, 30, [3, 20, 64, 130], [3, 20, 64, 130, 300], related=0.6, dirty=0.0, alpha=b"AC")
    tpk = helpers.pack(tp, 2)
    res = helpers.run_lib(emu_ctx, kind, False, T, 0, DNA_IDX, 5, tpk, 2)
    assert res.stats["pairs_fast16"] >= len(tp) // 2
    helpers.compare(res, helpers.run_oracle(kind, False, T, 0, DNA_IDX, tpk, 2), f"fast16 linear ties kind {kind}")
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    P = matrix.with_gap(matrix.BLOSUM62, -4)
    ppairs = randcases.rand_pairs(rng, 20, [3, 20, 45, 80, 300], [2, 20, 45, 80, 300], related=0.8, dirty=0.15,
                                  alpha=aa, dirty_alpha=aa + b"BZX*J-acdu")
    ppk = helpers.pack(ppairs)
    res = helpers.run_lib(emu_ctx, kind, False, P, 0, PROT_IDX, 26, ppk)
    assert res.stats["pairs_fast16"] > 0
    helpers.compare(res, helpers.run_oracle(kind, False, P, 0, PROT_IDX, ppk), f"fast16 linear protein kind {kind}")
    long_pairs = randcases.rand_pairs(rng, 5, [9, 60, 200], [800, 1025, 1300], related=0.85, dirty=0.0)
    lpk = helpers.pack(long_pairs)
    emu_ctx.set_option("long_mode", 2)
    try:
        res = helpers.run_lib(emu_ctx, kind, False, M, 0, DNA_IDX, 5, lpk)
    finally:
        emu_ctx.set_option("long_mode", 1)
    assert res.stats["wavefront_launches"] >= 1
    helpers.compare(res, helpers.run_oracle(kind, False, M, 0, DNA_IDX, lpk), f"fast16 linear wavefront kind {kind}")


def test_linear_sw_with_zero_gap_scores_takes_the_exact_path(emu_ctx):
    """er = eq = 0 breaks 'the maximum is always reached through the diagonal': not eligible."""
    rng = random.Random(4800)
    M = gv.match(0, 1, -1)
    pairs = randcases.rand_pairs(rng, 12, [5, 30, 70], [5, 30, 70], related=0.7, dirty=0.0)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(emu_ctx, 0, False, M, 0, DNA_IDX, 5, pk)
    assert res.stats["pairs_fast16"] == 0
    helpers.compare(res, helpers.run_oracle(0, False, M, 0, DNA_IDX, pk), "linear SW gap 0")


def test_plan_cache_reuses_work_lists_for_same_shaped_batches(emu_ctx):
    """Same length arrays, different letters: the cached plan must not leak results between
    calls; a new scoring, new lengths or a changed option drop it."""
    rng = random.Random(4900)
    lens_r, lens_q = [40, 90, 150, 33], [30, 64, 100, 257]
    shape = [(rng.choice(lens_r), rng.choice(lens_q)) for _ in range(24)]

    def batch(seed):
        r = random.Random(seed)
        out = []
        for n, m in shape:
            ref = bytes(r.choice(b"ACGT") for _ in range(n))
            qry = bytearray(ref[:m].ljust(m, b"A"))
            for _ in range(max(1, m // 10)):
                qry[r.randrange(m)] = r.choice(b"ACGT")
            out.append((ref, bytes(qry)))
        return out

    M = gv.match(-1, 2, -1)
    for kind, affine, go in ((0, True, -4), (1, True, -4), (0, False, 0)):
        for seed in (1, 2, 2, 3):
            pk = helpers.pack(batch(seed))
            res = helpers.run_lib(emu_ctx, kind, affine, M, go, DNA_IDX, 5, pk)
            helpers.compare(res, helpers.run_oracle(kind, affine, M, go, DNA_IDX, pk), f"cache {kind} {affine} seed {seed}")
    # an overflowing pair in between (exact path reuses the work-list buffers) and a different shape
    emu_ctx.set_option("slot_cap", 2)
    try:
        pk = helpers.pack(batch(5))
        res = helpers.run_lib(emu_ctx, 0, True, M, -4, DNA_IDX, 5, pk)
    finally:
        emu_ctx.set_option("slot_cap", 0)
    helpers.compare(res, helpers.run_oracle(0, True, M, -4, DNA_IDX, pk), "cache after overflow")
    pk = helpers.pack(batch(6))
    res = helpers.run_lib(emu_ctx, 0, True, M, -4, DNA_IDX, 5, pk)
    helpers.compare(res, helpers.run_oracle(0, True, M, -4, DNA_IDX, pk), "cache rebuilt")
    pk = helpers.pack(batch(6)[:11])
    res = helpers.run_lib(emu_ctx, 0, True, M, -4, DNA_IDX, 5, pk)
    helpers.compare(res, helpers.run_oracle(0, True, M, -4, DNA_IDX, pk), "cache other n")
    emu_ctx.set_option("plan_cache", 0)
    try:
        res = helpers.run_lib(emu_ctx, 0, True, M, -4, DNA_IDX, 5, pk)
    finally:
        emu_ctx.set_option("plan_cache", 1)
    helpers.compare(res, helpers.run_oracle(0, True, M, -4, DNA_IDX, pk), "cache off")


@pytest.mark.parametrize("affine", [True, False], ids=["SWAffine", "SW"])
def test_sw_best_cell_ties_across_passes(emu_ctx, affine):
    """Unrelated sequences with a harsh mismatch score: the table maximum is small and attained in
    many cells, often in different 256-column passes.  The reference keeps the LAST one in row-major
    order (align/sw_affine_letters.go:126-134, sw_letters.go:110-114) whatever pass it lies in."""
    rng = random.Random(5900 + affine)
    M = gv.match(-2, 1, -4)
    pairs = randcases.rand_pairs(rng, 10, [150, 330], [520, 700], related=0.0, dirty=0.0)
    pk = helpers.pack(pairs)
    go = -2 if affine else 0
    orc = helpers.run_oracle(0, affine, M, go, DNA_IDX, pk)
    for force in (1, 0):
        emu_ctx.set_option("force_path", force)
        try:
            res = helpers.run_lib(emu_ctx, 0, affine, M, go, DNA_IDX, 5, pk)
        finally:
            emu_ctx.set_option("force_path", 0)
        helpers.compare(res, orc, f"ties across passes, force_path {force}")


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SWAffine", "NWAffine", "FittedAffine"])
def test_fast16_general_and_symmetric_gap_kernels(emu_ctx, kind):
    """The SYM fill variant (er == eq: M is kept as M + go + e) and the general one must agree with
    the oracle: symmetric scores with the variant switched off, and asymmetric gap scores (which
    can only take the general kernels)."""
    rng = random.Random(6100 + kind)
    pairs = randcases.rand_pairs(rng, 26, [1, 9, 40, 130, 260], [1, 16, 64, 100, 257, 300, 600], related=0.85, dirty=0.0)
    pk = helpers.pack(pairs)
    M = gv.match(-1, 2, -1)
    orc = helpers.run_oracle(kind, True, M, -4, DNA_IDX, pk)
    for sym in (1, 0):
        emu_ctx.set_option("sym_kernels", sym)
        try:
            res = helpers.run_lib(emu_ctx, kind, True, M, -4, DNA_IDX, 5, pk)
        finally:
            emu_ctx.set_option("sym_kernels", 1)
        assert res.stats["pairs_fast16"] >= len(pairs) - 3
        helpers.compare(res, orc, f"sym_kernels {sym} kind {kind}")
    A = [row[:] for row in M]
    for a in range(1, 5):
        A[a][0] = -1      # reference letter against a gap
        A[0][a] = -2      # gap against a query letter
    res = helpers.run_lib(emu_ctx, kind, True, A, -4, DNA_IDX, 5, pk)
    assert res.stats["pairs_fast16"] >= len(pairs) - 3
    helpers.compare(res, helpers.run_oracle(kind, True, A, -4, DNA_IDX, pk), f"asymmetric gaps kind {kind}")


@pytest.mark.parametrize("kind,affine", helpers.KINDS, ids=[helpers.KIND_NAMES[k] for k in helpers.KINDS])
@pytest.mark.parametrize("setup", [
    dict(waves=3),
    dict(waves=5, fill_streams=1, fill_claims=0),
    dict(waves=4, arena_bytes=40000, fill_claims=2, trace_cta=64),      # fewer checkpoint regions than waves
    dict(waves=3, upload_groups=4, trace_cta=128, fill_cap=1),
    dict(waves=1, upload_groups=5, chunk_pairs=8),                       # one wave, fill in chunks as the upload arrives
    dict(waves=1, upload_groups=3, chunk_pairs=8, fill_streams=1),
], ids=["w3", "w5-one-stream-persistent", "w4-region-reuse", "w3-upload-groups", "chunked", "chunked-one-stream"])
def test_overlapped_waves_do_not_change_results(emu_ctx, kind, affine, setup):
    """Batches cut into several waves (traceback of wave w beside the fill of wave w+1), with the
    per-wave device bookkeeping, region reuse and the grouped upload: same results as one wave."""
    rng = random.Random(4100 + kind + 10 * affine)
    M = gv.match(-1, 2, -1) if not affine else gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 70, [1, 20, 45, 80, 130], [3, 17, 40, 64, 100, 300], related=0.7, dirty=0.05)
    pk = helpers.pack(pairs)
    for k, v in setup.items():
        emu_ctx.set_option(k, v)
    try:
        res = helpers.run_lib(emu_ctx, kind, affine, M, -3, DNA_IDX, 5, pk)
        assert res.stats["waves"] >= min(setup["waves"], 2)
        if "chunk_pairs" in setup:
            assert res.stats["fill_launches"] > res.stats["waves"] + 1
    finally:
        for k in setup:
            emu_ctx.set_option(k, {"fill_streams": 2, "fill_claims": 1}.get(k, 0))
    helpers.compare(res, helpers.run_oracle(kind, affine, M, -3, DNA_IDX, pk),
                    f"{helpers.KIND_NAMES[(kind, affine)]} {setup}")


def test_overlapped_waves_with_slot_overflow_and_mixed_alphabet(emu_ctx):
    """Pairs that overflow their segment slots or leave the 4-letter class are collected on the device
    across all waves and re-run on the exact path; the rest keep their fast-path results."""
    rng = random.Random(77)
    M = gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 60, [30, 60, 90], [20, 50, 90], related=0.6, dirty=0.0)
    pairs += [(b"ACGTNNACGT", b"ACGTACGT"), (b"ACG-TACGT", b"ACGTACGT")]   # illegal letter; gap letter: exact path
    rng.shuffle(pairs)
    pk = helpers.pack(pairs, 2)
    emu_ctx.set_option("waves", 4)
    emu_ctx.set_option("slot_cap", 3)
    try:
        res = helpers.run_lib(emu_ctx, 0, True, M, -2, DNA_IDX, 5, pk, 2)
        assert res.stats["waves"] >= 4
        assert 0 < res.stats["pairs_fast16"] < len(pairs)
    finally:
        emu_ctx.set_option("waves", 0)
        emu_ctx.set_option("slot_cap", 0)
    helpers.compare(res, helpers.run_oracle(0, True, M, -2, DNA_IDX, pk, 2), "overflow across waves")


def test_wavefront_timeout_falls_back_to_the_pass_by_pass_kernels(emu_ctx):
    """A wavefront pass that gives up waiting (never observed; injected here) must not fail the call: the
    batch is re-run with the pass-by-pass kernels and the results are the oracle's."""
    rng = random.Random(4242)
    M = gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 5, [40, 130], [1025, 1300, 1700], related=0.85, dirty=0.0)
    pk = helpers.pack(pairs)
    emu_ctx.set_option("long_mode", 2)
    emu_ctx.set_option("test_long_timeout", 1)
    try:
        res = helpers.run_lib(emu_ctx, 2, True, M, -4, DNA_IDX, 5, pk)
    finally:
        emu_ctx.set_option("long_mode", 1)
        emu_ctx.set_option("test_long_timeout", 0)
    assert res.stats["wavefront_launches"] >= 1          # the first attempt did use the wavefront
    helpers.compare(res, helpers.run_oracle(2, True, M, -4, DNA_IDX, pk), "wavefront retry")
    again = helpers.run_lib(emu_ctx, 2, True, M, -4, DNA_IDX, 5, pk)     # and the context is fine afterwards
    helpers.compare(again, helpers.run_oracle(2, True, M, -4, DNA_IDX, pk), "after retry")


@pytest.mark.parametrize("stride", [1, 2], ids=["Letters", "QLetters"])
def test_validation_word_scan_every_alignment(emu_ctx, stride):
    """Which error the reference raises first (and where) for stray illegal / gap letters at every alignment of the
    sequences inside the packed buffer -- ba_validate_kernel scans aligned 32-bit words."""
    rng = random.Random(4242 + stride)
    pairs = randcases.sparse_dirty_pairs(rng, 160)
    pk = helpers.pack(pairs, stride)
    M = gv.match(-1, 1, -1)
    for kind, affine in helpers.KINDS:
        res = helpers.run_lib(emu_ctx, kind, affine, M, -5 if affine else 0, DNA_IDX, 5, pk, stride)
        helpers.compare(res, helpers.run_oracle(kind, affine, M, -5 if affine else 0, DNA_IDX, pk, stride),
                        f"validation scan kind {kind} affine {affine} stride {stride}")


def test_region_reuse_under_a_chunked_fill(emu_ctx):
    """Arena-limited waves + grouped upload + two fill streams (the stream race the GPU fuzzer found)."""
    randcases.region_reuse_chunked_case(emu_ctx, helpers, gv.match, DNA_IDX, 2)


@pytest.mark.parametrize("go,gap", [(-1900, -100), (-1990, -11), (-40, -1), (0, 0)],
                         ids=["bias2000", "bias2001-general", "bias41", "bias0"])
def test_biased_sym_sw_kernel_at_its_bounds(emu_ctx, go, gap):
    """The SYM SW fill keeps every value biased by vb = -(go + e) (one 32-bit subtraction gives M + go + e for both
    packed pairs): the largest admitted bias (2000), one beyond it (general kernels), a small one, and zero gap
    scores (vb = 0, where SW's exactness condition sends affine pairs to the exact path) -- each with scores near
    the top of the 16-bit range (long identical stretches, match 10)."""
    randcases.biased_sym_sw_case(emu_ctx, helpers, gv.match, DNA_IDX, go, gap)
