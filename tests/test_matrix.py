"""align/matrix as a library: the 78 scoring tables by name.

Pins: (1) every table the library serves equals the table parsed from the reference's
align/matrix/matrices.go when that tree is present (build container); (2) its SHA-256 equals the
committed digest (tests/golden/matrices.json, written by tools/gen_matrices.py) everywhere;
(3) the hand-restated NUC_4 / BLOSUM62 of biogo_b200/matrix.py and Match equal the reference's
(matrices.go:28-35, :842-870; generated.go:10-31)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import golden_vectors as gv
import helpers
from biogo_b200 import _lib, alphabet, matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "matrices.json")))


def _digest(rows):
    return hashlib.sha256(",".join(str(v) for r in rows for v in r).encode()).hexdigest()


def test_library_serves_all_78_tables_with_the_committed_digests(built):
    names = matrix.names()
    assert len(names) == 78 and sorted(names) == sorted(GOLD)
    for name in names:
        rows = matrix.by_name(name)
        assert len(rows) == GOLD[name]["let"] and all(len(r) == len(rows) for r in rows)
        assert _digest(rows) == GOLD[name]["sha256"], name
        assert all(v == 0 for v in rows[0]) and all(r[0] == 0 for r in rows)   # no gap penalty (matrices.go:10-13)
    assert matrix.letters("NUC_4") == "-ACGT" and matrix.letters("BLOSUM62") == "-ABCDEFGHIJKLMNPQRSTVWXYZ*"
    assert matrix.letters("NUC_4_4") == "-ACMGRSVTWYHKDBN"
    with pytest.raises(KeyError):
        matrix.by_name("BLOSUM63")


def test_tables_equal_the_reference_file(built):
    ref = "/root/reference/align/matrix/matrices.go"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this box (digests are checked instead)")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_matrices

    parsed = gen_matrices.parse(ref)
    assert len(parsed) == 78
    for name, header, rows in parsed:
        assert matrix.by_name(name) == rows, name
        if header and len(header) == len(rows):
            assert matrix.letters(name) == "".join(header)


def test_restated_tables_and_match(built):
    assert matrix.NUC_4 == matrix.by_name("NUC_4")
    assert matrix.BLOSUM62 == matrix.by_name("BLOSUM62")
    # alphabets index the tables directly (alphabet/alphabet.go:26-79)
    assert alphabet.DNAgapped.Letters().upper().startswith("-ACGT") or alphabet.DNAgapped.Len() == 5
    assert alphabet.Protein.Len() == 26 == len(matrix.BLOSUM62)
    # generated.go:10-31
    m = matrix.Match(alphabet.DNAgapped, -1, 2, -1)
    g = gv.match(-1, 2, -1)            # the examples' hand-written matrix: the same except [0][0] = 0
    assert m[0][0] == -1 and g[0][0] == 0
    g[0][0] = -1
    assert m == g
    assert matrix.with_gap(matrix.NUC_4, -2)[0] == [0, -2, -2, -2, -2]


def test_named_scoring_equals_explicit_scoring_through_the_emulated_kernels(emu_ctx):
    """ba_set_scoring_named(BLOSUM62, gap -1) aligns exactly like ba_set_scoring with the hand-built table
    (and both like the oracle); switching back and forth reuses the resident device copies."""
    import random

    import randcases

    idx = helpers.O.build_index("-abcdefghijklmnpqrstvwxyz*")
    rng = random.Random(5)
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    pairs = randcases.rand_pairs(rng, 12, [20, 45], [20, 45, 70], related=0.8, dirty=0.0, alpha=aa)
    pk = helpers.pack(pairs)
    M = matrix.with_gap(matrix.BLOSUM62, -1)
    for kind, affine in ((1, True), (0, False)):
        want = helpers.run_oracle(kind, affine, M, -10 if affine else 0, idx, pk)
        emu_ctx.set_scoring_named(kind, affine, "BLOSUM62", -1, 26, -10 if affine else 0, idx)
        helpers.compare(emu_ctx.align_packed(pk[0], 1, pk[1], pk[2], pk[3], pk[4]), want, "named")
        helpers.compare(helpers.run_lib(emu_ctx, kind, affine, M, -10 if affine else 0, idx, 26, pk), want, "explicit")
    didx = helpers.O.build_index("-acgt")
    dpk = helpers.pack([(b"ACGTTGCA", b"ACGTGCA")])
    emu_ctx.set_scoring_named(0, True, "NUC_4", -2, 5, -10, didx)
    helpers.compare(emu_ctx.align_packed(dpk[0], 1, dpk[1], dpk[2], dpk[3], dpk[4]),
                    helpers.run_oracle(0, True, matrix.with_gap(matrix.NUC_4, -2), -10, didx, dpk), "NUC_4 named")
    with pytest.raises(_lib.AlignError):
        emu_ctx.set_scoring_named(0, True, "NO_SUCH_TABLE", -1, 5, -5, didx)
