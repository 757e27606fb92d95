"""Parity tests proper: the sm_100a CUDA path, through the C ABI, against the oracle."""
import random

import numpy as np
import pytest

import golden_vectors as gv
import helpers
import randcases
from biogo_b200 import align, alphabet, feat, matrix, seq, synth
from helpers import O

pytestmark = pytest.mark.gpu

DNA_IDX = O.build_index("-acgt")
PROT_IDX = O.build_index("-abcdefghijklmnpqrstvwxyz*")
CLS = {(0, False): align.SW, (1, False): align.NW, (2, False): align.Fitted,
       (0, True): align.SWAffine, (1, True): align.NWAffine, (2, True): align.FittedAffine}


@pytest.mark.parametrize("vec", gv.GOLDEN, ids=[v[0] for v in gv.GOLDEN])
def test_reference_golden_vectors_like_the_go_examples(gpu_ctx, vec):
    """Reads like align/sw_example_test.go: build Seqs, Align, print, Format."""
    name, kind, affine, M, go, ref, qry, want, fmt = vec
    a = CLS[(kind, affine)](Matrix=M, GapOpen=go) if affine else CLS[(kind, affine)](M)
    sa = seq.Seq(Seq=alphabet.BytesToLetters(ref.encode()), Alpha=alphabet.DNAgapped)
    sb = seq.Seq(Seq=alphabet.BytesToLetters(qry.encode()), Alpha=alphabet.DNAgapped)
    aln, err = a.Align(sa, sb)
    assert err is None
    assert feat.Sprint(aln) == want
    if fmt is not None:
        fa = align.Format(sa, sb, aln, "-")
        assert (str(fa[0]), str(fa[1])) == fmt
    # the same through QLetters (quality ignored)
    qa = seq.QSeq(Seq=alphabet.QLetters(ref.encode(), 40), Alpha=alphabet.DNAgapped)
    qb = seq.QSeq(Seq=alphabet.QLetters(qry.encode(), 40), Alpha=alphabet.DNAgapped)
    aln2, err = a.Align(qa, qb)
    assert err is None and feat.Sprint(aln2) == want


@pytest.mark.parametrize("vec", gv.DERIVED, ids=[v[0] for v in gv.DERIVED])
def test_derived_anchors(gpu_ctx, vec):
    name, kind, affine, alpha, M, go, ref, qry, want = vec
    if alpha == "protein":
        M, idx, alen = matrix.with_gap(matrix.BLOSUM62, -1), PROT_IDX, 26
    else:
        idx, alen = DNA_IDX, 5
    res = helpers.run_lib(gpu_ctx, kind, affine, M, go, idx, alen, helpers.pack([(ref.encode(), qry.encode())]))
    assert O.format_pairs(res.pairs(0)) == want


@pytest.mark.parametrize("kind,affine", helpers.KINDS, ids=[helpers.KIND_NAMES[k] for k in helpers.KINDS])
def test_random_batches_all_kinds(gpu_ctx, kind, affine):
    rng = random.Random(1000 + kind + 10 * affine)
    for trial in range(30):
        M = randcases.rand_matrix(rng)
        go = rng.choice([0, -1, -2, -5, 1])
        stride = rng.choice([1, 2])
        pairs = randcases.rand_pairs(rng, 200, [0, 1, 2, 3, 5, 8, 13, 30, 60, 150],
                                     [0, 1, 2, 3, 5, 9, 17, 31, 32, 33, 64, 65, 100, 129, 200, 256])
        pk = helpers.pack(pairs, stride)
        res = helpers.run_lib(gpu_ctx, kind, affine, M, go, DNA_IDX, 5, pk, stride)
        assert res.stats["gpu_launches"] > 0
        helpers.compare(res, helpers.run_oracle(kind, affine, M, go, DNA_IDX, pk, stride, 16),
                        f"{helpers.KIND_NAMES[(kind, affine)]} trial {trial} M={M} go={go} stride={stride}")


@pytest.mark.parametrize("kind,affine", helpers.KINDS, ids=[helpers.KIND_NAMES[k] for k in helpers.KINDS])
def test_long_queries_multi_pass(gpu_ctx, kind, affine):
    rng = random.Random(2000 + kind + 10 * affine)
    M = gv.match(-1, 2, -1) if not affine else gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 60, [100, 700, 1500, 3000], [257, 300, 512, 513, 1000, 2500], related=0.8,
                                 dirty=0.0)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(gpu_ctx, kind, affine, M, -4, DNA_IDX, 5, pk)
    helpers.compare(res, helpers.run_oracle(kind, affine, M, -4, DNA_IDX, pk, 1, 8),
                    f"{helpers.KIND_NAMES[(kind, affine)]} long")


def test_protein_blosum62_all_kinds(gpu_ctx):
    rng = random.Random(9)
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    M = matrix.with_gap(matrix.BLOSUM62, -1)
    pairs = randcases.rand_pairs(rng, 400, [20, 45, 80, 300], [20, 45, 80, 300], related=0.8, dirty=0.1, alpha=aa,
                                 dirty_alpha=aa + b"BZX*J-acdu")
    pk = helpers.pack(pairs)
    for kind, affine in helpers.KINDS:
        res = helpers.run_lib(gpu_ctx, kind, affine, M, -10, PROT_IDX, 26, pk)
        helpers.compare(res, helpers.run_oracle(kind, affine, M, -10, PROT_IDX, pk, 1, 16), f"protein {kind} {affine}")


@pytest.mark.parametrize("num,n", [(1, 3000), (2, 2000), (3, 3000), (5, 300)])
def test_baseline_config_shapes_against_oracle(gpu_ctx, num, n):
    """Seeded sub-samples of BASELINE configs 1, 2, 3, 5 at their real shapes."""
    w = synth.config(num, n)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    go = w.gap_open if w.affine else 0
    res = helpers.run_lib(gpu_ctx, w.kind, w.affine, w.matrix, go, w.alpha.LetterIndex(), w.alpha.Len(), pk, w.stride)
    helpers.compare(res, helpers.run_oracle(w.kind, w.affine, w.matrix, go, w.alpha.LetterIndex(), pk, w.stride, 16),
                    w.name)


def test_config4_long_pair_full_size(gpu_ctx):
    """One 10 kb x 12 kb FittedAffine pair (120 M cells; the oracle needs 2.9 GB for it)."""
    w = synth.config(4, 1)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    res = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    helpers.compare(res, helpers.run_oracle(w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), pk, 1, 1),
                    w.name)


def _rescore(w, res, p):
    """Oracle-free property: segments tile the path contiguously and their scores re-add from
    the letters and the matrix."""
    r, q = w.pair(p)
    idx = w.alpha.LetterIndex()
    M, go = w.matrix, (w.gap_open if w.affine else 0)
    segs = res.pairs(p)
    total = 0
    for k, (a0, a1, b0, b1, sc) in enumerate(segs.tolist()):
        assert 0 <= a0 <= a1 <= len(r) and 0 <= b0 <= b1 <= len(q)
        if k:
            assert (a0, b0) == (segs[k - 1][1], segs[k - 1][3])
        if a1 - a0 == b1 - b0:
            s = sum(M[idx[r[a0 + x]]][idx[q[b0 + x]]] for x in range(a1 - a0))
        elif a1 == a0:
            s = sum(M[0][idx[q[x]]] for x in range(b0, b1)) + (go if b1 > b0 else 0)
        else:
            assert b1 == b0
            s = sum(M[idx[r[x]]][0] for x in range(a0, a1)) + (go if a1 > a0 else 0)
        assert s == sc, (p, k, segs.tolist())
        total += sc
    return total


def test_full_size_shard_properties(gpu_ctx):
    """BASELINE config 2 at bench size (one GPU's shard of 125 000 pairs): every pair OK,
    segment chains consistent, and a checksum that is stable from run to run."""
    w = synth.config(2, 125_000)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    res = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    assert res.n == w.n and int(np.count_nonzero(res.status)) == 0
    assert res.stats["cells"] == w.cells
    rng = random.Random(5)
    for p in rng.sample(range(w.n), 300):
        assert _rescore(w, res, p) > 50  # a 250 bp read with 5 % errors aligns with a clearly positive score
    res2 = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    for p in rng.sample(range(w.n), 2000):
        assert np.array_equal(res.pairs(p), res2.pairs(p))
    # oracle on a seeded subsample of the same shard
    sub = np.array(sorted(rng.sample(range(w.n), 1500)))
    ws = w.subset(sub)
    orc = helpers.run_oracle(w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(),
                             (ws.letters, ws.ref_off, ws.ref_len, ws.qry_off, ws.qry_len), 1, 16)
    for k, p in enumerate(sub):
        want = orc[3][orc[2][k]:orc[2][k + 1]]
        assert np.array_equal(res.pairs(int(p)).astype(np.int64), want), p


def test_small_arena_waves_match(gpu_ctx):
    w = synth.config(2, 600)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    full = helpers.run_lib(gpu_ctx, 0, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    gpu_ctx.set_option("arena_bytes", 40 << 20)
    try:
        waves = helpers.run_lib(gpu_ctx, 0, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    finally:
        gpu_ctx.set_option("arena_bytes", 0)
    for p in range(w.n):
        assert np.array_equal(full.pairs(p), waves.pairs(p))


def test_align_batch_host_mirror(gpu_ctx):
    rng = random.Random(11)
    pairs = randcases.rand_pairs(rng, 50, [10, 40, 90], [10, 40, 90])
    a = align.NWAffine(Matrix=gv.match(-1, 1, -1), GapOpen=-3)
    refs = [seq.Seq(Seq=r, Alpha=alphabet.DNAgapped) for r, _ in pairs]
    qrys = [seq.Seq(Seq=q, Alpha=alphabet.DNAgapped) for _, q in pairs]
    sc = O.Scoring(1, True, gv.match(-1, 1, -1), -3, DNA_IDX)
    try:
        alns, errs = a.AlignBatch(refs, qrys)
    except align.GoPanic:
        alns = None
    n_ok = 0
    for k, (r, q) in enumerate(pairs):
        st, segs, ep = O.align(sc, r, q)
        if alns is None:
            continue
        if st == 0:
            assert errs[k] is None and [p.astuple() for p in alns[k]] == segs
            n_ok += 1
    # single Align calls agree and panics surface as GoPanic
    for r, q in pairs[:10]:
        st, segs, ep = O.align(sc, r, q)
        sa, sb = seq.Seq(Seq=r, Alpha=alphabet.DNAgapped), seq.Seq(Seq=q, Alpha=alphabet.DNAgapped)
        if st == 3:
            with pytest.raises(align.GoPanic):
                a.Align(sa, sb)
        else:
            aln, err = a.Align(sa, sb)
            assert err is None and [p.astuple() for p in aln] == segs


@pytest.mark.parametrize("params", [(-1, 1, -1, -5), (-1, 2, -1, -3), (-2, 5, -4, -10), (0, 1, -1, -2)],
                         ids=["m1", "m2", "nuc4like", "gap0"])
def test_fast16_path_random(gpu_ctx, params):
    """Packed 16-bit DPX fill + checkpointed tile traceback against the oracle."""
    gap, ma, mi, go = params
    rng = random.Random(hash(params) & 0xffff)
    M = gv.match(gap, ma, mi)
    pairs = randcases.rand_pairs(rng, 1500, [1, 2, 7, 15, 16, 17, 31, 33, 48, 64, 90, 130, 400, 1000],
                                 [1, 2, 5, 8, 9, 16, 31, 32, 33, 40, 64, 65, 95, 100, 160, 250, 256, 257, 300, 700],
                                 related=0.85, dirty=0.02)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(gpu_ctx, 0, True, M, go, DNA_IDX, 5, pk)
    assert res.stats["pairs_fast16"] > 0.9 * len(pairs)
    helpers.compare(res, helpers.run_oracle(0, True, M, go, DNA_IDX, pk, 1, 16), f"fast16 {params}")


def test_fast16_and_exact_paths_agree_on_config2(gpu_ctx):
    w = synth.config(2, 4000)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    fast = helpers.run_lib(gpu_ctx, 0, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    assert fast.stats["pairs_fast16"] == w.n
    gpu_ctx.set_option("force_path", 1)
    try:
        exact = helpers.run_lib(gpu_ctx, 0, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    finally:
        gpu_ctx.set_option("force_path", 0)
    assert exact.stats["pairs_fast16"] == 0
    for p in range(w.n):
        assert np.array_equal(fast.pairs(p), exact.pairs(p)), p
    gpu_ctx.set_option("slot_cap", 3)
    try:
        ovf = helpers.run_lib(gpu_ctx, 0, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    finally:
        gpu_ctx.set_option("slot_cap", 0)
    assert 0 < ovf.stats["pairs_fast16"] < w.n
    for p in range(w.n):
        assert np.array_equal(ovf.pairs(p), exact.pairs(p)), p


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SWAffine", "NWAffine", "FittedAffine"])
def test_fast16_all_affine_kinds(gpu_ctx, kind):
    rng = random.Random(3000 + kind)
    M = gv.match(-1, 2, -1)
    pairs = randcases.rand_pairs(rng, 1200, [1, 5, 17, 40, 90, 150, 400, 1000],
                                 [1, 4, 16, 33, 64, 100, 160, 250, 257, 300, 700], related=0.85, dirty=0.02)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(gpu_ctx, kind, True, M, -4, DNA_IDX, 5, pk)
    assert res.stats["pairs_fast16"] > 0.85 * len(pairs)
    helpers.compare(res, helpers.run_oracle(kind, True, M, -4, DNA_IDX, pk, 1, 16), f"fast16 dna kind {kind}")
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    P = matrix.with_gap(matrix.BLOSUM62, -1)
    ppairs = randcases.rand_pairs(rng, 800, [3, 20, 45, 80, 300, 600], [2, 20, 45, 80, 300, 500], related=0.8,
                                  dirty=0.1, alpha=aa, dirty_alpha=aa + b"BZX*J-acdu")
    ppk = helpers.pack(ppairs, 2)
    res = helpers.run_lib(gpu_ctx, kind, True, P, -12, PROT_IDX, 26, ppk, 2)
    assert res.stats["pairs_fast16"] > 0.5 * len(ppairs)
    helpers.compare(res, helpers.run_oracle(kind, True, P, -12, PROT_IDX, ppk, 2, 16), f"fast16 protein kind {kind}")


def test_fitted_affine_long_pairs_fast_path(gpu_ctx):
    """FittedAffine on kb-scale pairs (BASELINE config 4 scaled down): many passes per pair."""
    rng = random.Random(41)
    M = gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 24, [1500, 3000, 4000], [1200, 2500, 3300], related=1.0, dirty=0.0)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(gpu_ctx, 2, True, M, -5, DNA_IDX, 5, pk)
    assert res.stats["pairs_fast16"] == len(pairs)
    helpers.compare(res, helpers.run_oracle(2, True, M, -5, DNA_IDX, pk, 1, 8), "fitted long")


def test_graft_entry_smoke(gpu_ctx):
    """The driver's smoke(): one small invocation of every aligner kind, checked against the oracle."""
    import __graft_entry__ as g

    g.smoke()


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SWAffine", "NWAffine", "FittedAffine"])
def test_wavefront_over_passes(gpu_ctx, kind):
    """Few long pairs: the passes of each pair run concurrently on different warps, handing the
    boundary column over through global memory (LONG mode of ba_fill16)."""
    rng = random.Random(5100 + kind)
    M = gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 13, [300, 1500, 3100], [900, 1025, 2500, 4100], related=0.9, dirty=0.0)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(gpu_ctx, kind, True, M, -5, DNA_IDX, 5, pk)
    assert res.stats["wavefront_launches"] >= 1
    orc = helpers.run_oracle(kind, True, M, -5, DNA_IDX, pk, 1, 8)
    helpers.compare(res, orc, f"wavefront kind {kind}")
    gpu_ctx.set_option("long_mode", 0)
    try:
        res0 = helpers.run_lib(gpu_ctx, kind, True, M, -5, DNA_IDX, 5, pk)
    finally:
        gpu_ctx.set_option("long_mode", 1)
    assert res0.stats["wavefront_launches"] == 0
    helpers.compare(res0, orc, f"no wavefront kind {kind}")


def test_wavefront_with_more_tasks_than_resident_warps(gpu_ctx):
    """Forced wavefront over a batch far larger than the GPU holds at once: passes are claimed in
    pass-major order, so a waiting warp's left neighbour is always already running."""
    rng = random.Random(5200)
    M = gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 6000, [60, 150], [800, 1030, 1300], related=0.9, dirty=0.0)
    pk = helpers.pack(pairs)
    gpu_ctx.set_option("long_mode", 2)
    try:
        res = helpers.run_lib(gpu_ctx, 0, True, M, -5, DNA_IDX, 5, pk)
    finally:
        gpu_ctx.set_option("long_mode", 1)
    assert res.stats["wavefront_launches"] >= 1
    helpers.compare(res, helpers.run_oracle(0, True, M, -5, DNA_IDX, pk, 1, 16), "wavefront oversubscribed")


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SWAffine", "NWAffine", "FittedAffine"])
def test_warp_per_pair_traceback_matches_thread_per_pair(gpu_ctx, kind):
    rng = random.Random(5300 + kind)
    M = gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 300, [1, 30, 250, 700, 1500], [1, 20, 150, 257, 600, 1300], related=0.85,
                                 dirty=0.0)
    pairs += randcases.rand_pairs(rng, 40, [100, 400], [100, 400], related=0.0, dirty=0.0)
    pk = helpers.pack(pairs)
    orc = helpers.run_oracle(kind, True, M, -5, DNA_IDX, pk, 1, 16)
    for mode, want in ((2, True), (3, True), (0, False), (5, False)):
        gpu_ctx.set_option("trace_mode", mode)
        try:
            res = helpers.run_lib(gpu_ctx, kind, True, M, -5, DNA_IDX, 5, pk)
        finally:
            gpu_ctx.set_option("trace_mode", 1)
        assert (res.stats["warp_trace_launches"] >= 1) == want
        helpers.compare(res, orc, f"trace_mode {mode} kind {kind}")


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["SW", "NW", "Fitted"])
def test_fast16_linear_kinds(gpu_ctx, kind):
    """Linear-gap aligners on the packed DPX path: DNA, Protein, long queries, both trace kernels."""
    rng = random.Random(5400 + kind)
    M = gv.match(-1, 2, -1)
    pairs = randcases.rand_pairs(rng, 1500, [1, 5, 17, 40, 150, 500, 900], [1, 4, 16, 64, 150, 257, 300, 1100],
                                 related=0.9, dirty=0.0)
    pk = helpers.pack(pairs)
    orc = helpers.run_oracle(kind, False, M, 0, DNA_IDX, pk, 1, 16)
    for mode in (0, 2, 3, 5):
        gpu_ctx.set_option("trace_mode", mode)
        try:
            res = helpers.run_lib(gpu_ctx, kind, False, M, 0, DNA_IDX, 5, pk)
        finally:
            gpu_ctx.set_option("trace_mode", 1)
        assert res.stats["pairs_fast16"] > 0.7 * len(pairs)
        helpers.compare(res, orc, f"fast16 linear kind {kind} trace_mode {mode}")
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    P = matrix.with_gap(matrix.BLOSUM62, -4)
    ppairs = randcases.rand_pairs(rng, 600, [20, 45, 80, 300], [20, 45, 80, 300], related=0.8, dirty=0.1, alpha=aa,
                                  dirty_alpha=aa + b"BZX*J-acdu")
    ppk = helpers.pack(ppairs, 2)
    res = helpers.run_lib(gpu_ctx, kind, False, P, 0, PROT_IDX, 26, ppk, 2)
    assert res.stats["pairs_fast16"] > 0.4 * len(ppairs)
    helpers.compare(res, helpers.run_oracle(kind, False, P, 0, PROT_IDX, ppk, 2, 16), f"fast16 linear protein {kind}")
    lp = randcases.rand_pairs(rng, 9, [700, 2500], [1100, 3000], related=0.9, dirty=0.0)
    lpk = helpers.pack(lp)
    res = helpers.run_lib(gpu_ctx, kind, False, M, 0, DNA_IDX, 5, lpk)
    assert res.stats["wavefront_launches"] >= 1
    helpers.compare(res, helpers.run_oracle(kind, False, M, 0, DNA_IDX, lpk, 1, 8), f"fast16 linear wavefront {kind}")


def test_config1_fast_and_exact_paths_agree(gpu_ctx):
    w = synth.config(1, 20000)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    idx = w.alpha.LetterIndex()
    fast = helpers.run_lib(gpu_ctx, w.kind, False, w.matrix, 0, idx, 5, pk)
    assert fast.stats["pairs_fast16"] > 0.99 * w.n
    gpu_ctx.set_option("force_path", 1)
    try:
        exact = helpers.run_lib(gpu_ctx, w.kind, False, w.matrix, 0, idx, 5, pk)
    finally:
        gpu_ctx.set_option("force_path", 0)
    assert exact.stats["pairs_fast16"] == 0
    assert np.array_equal(fast.status, exact.status)
    for p in range(w.n):
        assert np.array_equal(fast.pairs(p), exact.pairs(p)), p


def test_config4_nuc4_scores_promote_to_the_32bit_path(gpu_ctx):
    """BASELINE config 4, second run: +5/-4 scores (NUC_4-like, gaps -2, GapOpen -10) on a pair long
    enough that the best score leaves the int16 range: the pair must be promoted to the 32-bit
    exact-code kernels and still match the oracle."""
    rng = random.Random(5600)
    M = gv.match(-2, 5, -4)
    ref = bytes(rng.choice(b"ACGT") for _ in range(7600))
    qry = bytearray(ref[150:7350])
    for _ in range(len(qry) // 40):                   # 2.5 % substitutions: best score ~ 7200 * 4.8
        qry[rng.randrange(len(qry))] = rng.choice(b"ACGT")
    pairs = [(ref, bytes(qry))]
    pairs += randcases.rand_pairs(rng, 3, [700, 900], [600], related=1.0, dirty=0.0)   # these stay on the packed path
    pk = helpers.pack(pairs)
    res = helpers.run_lib(gpu_ctx, 2, True, M, -10, DNA_IDX, 5, pk)
    assert res.stats["pairs_fast16"] == 3
    best = int(res.pairs(0)[:, 4].sum())
    assert best > 32767, best
    helpers.compare(res, helpers.run_oracle(2, True, M, -10, DNA_IDX, pk, 1, 2), "cfg4 NUC_4-like")


def test_mixed_batch_gives_warps_to_its_longest_pairs(gpu_ctx):
    """Between ~4.7 k and ~75 k pairs of very different lengths the traceback runs the head of the
    size-sorted work list with one warp per pair and the rest with one thread per pair."""
    rng = random.Random(5700)
    M = gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 300, [2500, 3500], [1300, 2000], related=0.9, dirty=0.0)
    pairs += randcases.rand_pairs(rng, 6000, [60, 150, 300], [40, 100, 200], related=0.9, dirty=0.0)
    rng.shuffle(pairs)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(gpu_ctx, 0, True, M, -5, DNA_IDX, 5, pk)
    assert res.stats["warp_trace_launches"] >= 1 and res.stats["pairs_fast16"] > 0.95 * len(pairs)
    helpers.compare(res, helpers.run_oracle(0, True, M, -5, DNA_IDX, pk, 1, 16), "mixed trace")


@pytest.mark.parametrize("num,n", [(3, 20000), (5, 3000), (4, 6)])
def test_fast16_and_exact_paths_agree_at_config_shapes(gpu_ctx, num, n):
    """Oracle-free at sizes the CPU oracle cannot reach quickly: the packed DPX path against the
    32-bit exact-code path (itself pinned to the oracle) on BASELINE configs 3, 5 and 4."""
    w = synth.config(num, n)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    idx, al = w.alpha.LetterIndex(), w.alpha.Len()
    fast = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, idx, al, pk, w.stride)
    assert fast.stats["pairs_fast16"] > 0.9 * w.n
    gpu_ctx.set_option("force_path", 1)
    try:
        exact = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, idx, al, pk, w.stride)
    finally:
        gpu_ctx.set_option("force_path", 0)
    assert exact.stats["pairs_fast16"] == 0
    assert np.array_equal(fast.status, exact.status) and np.array_equal(fast.seg_count, exact.seg_count)
    # (the two paths lay their segments out in their own work order: compare pair by pair)
    for p in range(w.n):
        assert np.array_equal(fast.pairs(p), exact.pairs(p)), p


@pytest.mark.parametrize("affine", [True, False], ids=["SWAffine", "SW"])
def test_sw_best_cell_ties_across_passes(gpu_ctx, affine):
    """Many cells share the (small) table maximum of unrelated sequences, in different 256-column
    passes; the reference keeps the last one in row-major order.  Both device paths vs the oracle."""
    rng = random.Random(5950 + affine)
    M = gv.match(-2, 1, -4)
    pairs = randcases.rand_pairs(rng, 60, [300, 900], [520, 1030, 1500], related=0.0, dirty=0.0)
    pk = helpers.pack(pairs)
    go = -2 if affine else 0
    orc = helpers.run_oracle(0, affine, M, go, DNA_IDX, pk, 1, 16)
    for force in (1, 0):
        gpu_ctx.set_option("force_path", force)
        try:
            res = helpers.run_lib(gpu_ctx, 0, affine, M, go, DNA_IDX, 5, pk)
        finally:
            gpu_ctx.set_option("force_path", 0)
        helpers.compare(res, orc, f"ties across passes, force_path {force}")


OVERLAP_SETUPS = [
    dict(waves=3),
    dict(waves=6, fill_streams=1, fill_claims=0),
    dict(waves=5, arena_bytes=96 << 20, fill_claims=2, trace_cta=64),   # fewer checkpoint regions than waves
    dict(waves=4, upload_groups=5, trace_cta=128, fill_cap=4),
]


@pytest.mark.gpu
@pytest.mark.parametrize("num,n", [(1, 3000), (2, 2000), (3, 3000), (5, 300)])
@pytest.mark.parametrize("setup", OVERLAP_SETUPS, ids=["w3", "w6-one-stream-persistent", "w5-region-reuse",
                                                       "w4-upload-groups"])
def test_overlapped_waves_against_oracle(gpu_ctx, num, n, setup):
    """Traceback of wave w beside the fill of wave w+1 (separate streams, non-persistent fill
    grids, per-wave device bookkeeping, grouped upload): bit-exact against the oracle."""
    w = synth.config(num, n)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    go = w.gap_open if w.affine else 0
    for k, v in setup.items():
        gpu_ctx.set_option(k, v)
    try:
        res = helpers.run_lib(gpu_ctx, w.kind, w.affine, w.matrix, go, w.alpha.LetterIndex(), w.alpha.Len(), pk,
                              w.stride)
        assert res.stats["waves"] >= 2
    finally:
        for k in setup:
            gpu_ctx.set_option(k, {"fill_streams": 2, "fill_claims": 1}.get(k, 0))
    helpers.compare(res, helpers.run_oracle(w.kind, w.affine, w.matrix, go, w.alpha.LetterIndex(), pk, w.stride, 16),
                    f"{w.name} {setup}")


@pytest.mark.gpu
def test_waves_and_chunked_upload_equal_one_wave_on_a_large_batch(gpu_ctx):
    """40 000 pairs of config 2 (10 G cells, 51 MB of letters): the batch uploads in groups and the fill
    starts chunk by chunk on what has arrived; forced waves overlap traceback and fill.  Every pair's
    segments equal those of the plain single-piece, single-wave run."""
    w = synth.config(2, 40_000)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    gpu_ctx.set_option("upload_groups", 1)
    try:
        one = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    finally:
        gpu_ctx.set_option("upload_groups", 0)
    assert one.stats["waves"] == 1 and one.stats["fill_launches"] == 1
    chunked = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    assert chunked.stats["waves"] == 1 and chunked.stats["fill_launches"] >= 3
    gpu_ctx.set_option("waves", 4)
    try:
        waves = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    finally:
        gpu_ctx.set_option("waves", 0)
    assert waves.stats["waves"] >= 4
    for other in (chunked, waves):
        assert np.array_equal(one.status, other.status) and np.array_equal(one.seg_count, other.seg_count)
        for p in range(0, w.n, 7):
            assert np.array_equal(one.pairs(p), other.pairs(p)), p
    # twice in a row through the cached plan
    again = helpers.run_lib(gpu_ctx, w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    for p in range(0, w.n, 11):
        assert np.array_equal(one.pairs(p), again.pairs(p)), p


def test_wavefront_timeout_falls_back_to_the_pass_by_pass_kernels(gpu_ctx):
    """A wavefront pass that gives up waiting (injected) re-runs the batch with the pass-by-pass kernels."""
    rng = random.Random(4242)
    M = gv.match(-1, 1, -1)
    pairs = randcases.rand_pairs(rng, 9, [400, 1300], [1025, 1300, 2100], related=0.85, dirty=0.0)
    pk = helpers.pack(pairs)
    gpu_ctx.set_option("long_mode", 2)
    gpu_ctx.set_option("test_long_timeout", 1)
    try:
        res = helpers.run_lib(gpu_ctx, 2, True, M, -4, DNA_IDX, 5, pk)
    finally:
        gpu_ctx.set_option("long_mode", 1)
        gpu_ctx.set_option("test_long_timeout", 0)
    assert res.stats["wavefront_launches"] >= 1
    helpers.compare(res, helpers.run_oracle(2, True, M, -4, DNA_IDX, pk, 1, 8), "wavefront retry")


def test_named_scoring_on_the_gpu(gpu_ctx):
    """ba_set_scoring_named (align/matrix by name) against the explicit table and the oracle, and the resident
    scoring cache when alternating between two aligners."""
    rng = random.Random(8)
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    pairs = randcases.rand_pairs(rng, 200, [60, 150, 300], [60, 150, 300], related=0.8, dirty=0.0, alpha=aa)
    pk = helpers.pack(pairs)
    for name, gap, go in (("BLOSUM62", -1, -10), ("PAM250", -2, -8), ("BLOSUM80", -1, -12)):
        M = matrix.with_gap(matrix.by_name(name), gap)
        want = helpers.run_oracle(1, True, M, go, PROT_IDX, pk, 1, 16)
        for _ in range(2):                                  # second round: both scorings come from the cache
            gpu_ctx.set_scoring_named(1, True, name, gap, 26, go, PROT_IDX)
            helpers.compare(gpu_ctx.align_packed(pk[0], 1, pk[1], pk[2], pk[3], pk[4]), want, f"named {name}")
            helpers.compare(helpers.run_lib(gpu_ctx, 0, False, gv.match(-1, 2, -1), 0, DNA_IDX, 5,
                                            helpers.pack([(b"ACGTACGTTT", b"ACGTCGTTT")])),
                            helpers.run_oracle(0, False, gv.match(-1, 2, -1), 0, DNA_IDX,
                                               helpers.pack([(b"ACGTACGTTT", b"ACGTCGTTT")])), "dna between")


@pytest.mark.gpu
@pytest.mark.parametrize("stride", [1, 2], ids=["Letters", "QLetters"])
def test_validation_word_scan_every_alignment(gpu_ctx, stride):
    """Twin of the emulator-tier test: stray illegal / gap letters at every word alignment, all six aligners."""
    rng = random.Random(4242 + stride)
    pairs = randcases.sparse_dirty_pairs(rng, 600)
    pk = helpers.pack(pairs, stride)
    M = gv.match(-1, 1, -1)
    for kind, affine in helpers.KINDS:
        res = helpers.run_lib(gpu_ctx, kind, affine, M, -5 if affine else 0, DNA_IDX, 5, pk, stride)
        helpers.compare(res, helpers.run_oracle(kind, affine, M, -5 if affine else 0, DNA_IDX, pk, stride),
                        f"validation scan kind {kind} affine {affine} stride {stride}")


@pytest.mark.gpu
def test_region_reuse_under_a_chunked_fill(gpu_ctx):
    """Arena-limited waves + grouped upload + two fill streams (the stream race the GPU fuzzer found)."""
    randcases.region_reuse_chunked_case(gpu_ctx, helpers, gv.match, DNA_IDX, 40)


@pytest.mark.gpu
@pytest.mark.parametrize("go,gap", [(-1900, -100), (-1990, -11), (-40, -1), (0, 0)],
                         ids=["bias2000", "bias2001-general", "bias41", "bias0"])
def test_biased_sym_sw_kernel_at_its_bounds(gpu_ctx, go, gap):
    """Twin of the emulator-tier test: the biased SYM SW fill at the bounds of its bias, scores near the 16-bit top."""
    randcases.biased_sym_sw_case(gpu_ctx, helpers, gv.match, DNA_IDX, go, gap)
