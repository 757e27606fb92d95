"""The oracle against every golden vector the reference holds for the path (G1..G8) plus the
survey's derived anchors, an independent pure-Python restatement, and boundary behaviour."""
import random

import numpy as np
import pytest

import golden_vectors as gv
from biogo_b200 import alphabet, matrix
from oracle import oracle as O
from oracle import py_oracle as P

DNA_IDX = O.build_index("-acgt")
PROT_IDX = O.build_index("-abcdefghijklmnpqrstvwxyz*")


@pytest.mark.parametrize("vec", gv.GOLDEN, ids=[v[0] for v in gv.GOLDEN])
def test_reference_golden_vectors(vec):
    name, kind, affine, M, go, ref, qry, want, fmt = vec
    st, segs, _ = O.align(O.Scoring(kind, affine, M, go, DNA_IDX), ref.encode(), qry.encode())
    assert st == O.ORC_OK
    assert O.format_pairs(segs) == want
    if fmt is not None:
        assert O.format_alignment(ref.encode(), qry.encode(), segs) == fmt


@pytest.mark.parametrize("vec", gv.DERIVED, ids=[v[0] for v in gv.DERIVED])
def test_derived_anchors(vec):
    name, kind, affine, alpha, M, go, ref, qry, want = vec
    if alpha == "protein":
        M, idx = matrix.with_gap(matrix.BLOSUM62, -1), PROT_IDX
    else:
        idx = DNA_IDX
    st, segs, _ = O.align(O.Scoring(kind, affine, M, go, idx), ref.encode(), qry.encode())
    assert st == O.ORC_OK
    assert O.format_pairs(segs) == want


def test_index_matches_python_alphabets():
    for a, letters in ((alphabet.DNAgapped, "-acgt"), (alphabet.Protein, "-abcdefghijklmnpqrstvwxyz*"),
                       (alphabet.DNAredundant, "-acmgrsvtwyhkdbn"), (alphabet.DNA, "acgt")):
        assert np.array_equal(O.build_index(letters), a.LetterIndex())
    assert alphabet.DNA.IndexOf(alphabet.DNA.Gap()) == -1  # why every aligner rejects alphabet.DNA
    assert alphabet.DNAgapped.IndexOf(ord("-")) == 0


def test_two_independent_restatements_agree():
    """C oracle vs pure-Python oracle on random small cases (all kinds, odd matrices,
    illegal letters, empty inputs, QLetters stride)."""
    rng = random.Random(7)
    n_cmp = 0
    for trial in range(1500):
        kind, affine = rng.randrange(3), rng.random() < 0.6
        if rng.random() < 0.5:
            M = gv.match(rng.choice([0, -1, -2]), rng.choice([1, 2]), rng.choice([0, -1, -2]))
        else:
            M = [[rng.randint(-3, 3) for _ in range(5)] for _ in range(5)]
        go = rng.choice([0, -1, -3, 2])
        al = b"ACGT" if rng.random() < 0.85 else b"ACGTN-a"
        ref = bytes(rng.choice(al) for _ in range(rng.randrange(0, 12)))
        qry = bytes(rng.choice(al) for _ in range(rng.randrange(0, 12)))
        stride = rng.choice([1, 2])
        if stride == 2:
            r2 = bytes(b for ch in ref for b in (ch, 33))
            q2 = bytes(b for ch in qry for b in (ch, 33))
        else:
            r2, q2 = ref, qry
        st, segs, ep = O.align(O.Scoring(kind, affine, M, go, DNA_IDX), r2, q2, stride)
        pst, psegs, pep = P.align(kind, affine, M, go, list(DNA_IDX), ref, qry)
        assert (st, ep if st in (1, 2) else -1) == (pst, pep if pst in (1, 2) else -1), (kind, affine, M, go, ref, qry)
        if st == 0:
            assert segs == psegs, (kind, affine, M, go, ref, qry)
            n_cmp += 1
    assert n_cmp > 800


def test_error_and_panic_outcomes():
    M = gv.match(-1, 1, -1)
    sw = O.Scoring(0, False, M, 0, DNA_IDX)
    assert O.align(sw, b"NACG", b"ACGN")[0::2] == (O.ORC_ILLEGAL_REF, 0)      # ref[0] is checked first
    assert O.align(sw, b"ACGN", b"ANGT")[0::2] == (O.ORC_ILLEGAL_QRY, 1)      # then the first bad query letter
    assert O.align(sw, b"ACNN", b"ACGT")[0::2] == (O.ORC_ILLEGAL_REF, 2)      # then the first bad ref letter
    assert O.align(sw, b"", b"NNN")[0] == O.ORC_OK                            # nothing inspected
    assert O.align(sw, b"", b"")[1] == [(0, 0, 0, 0, 0)]
    nw = O.Scoring(1, False, M, 0, DNA_IDX)
    assert O.align(nw, b"ACGT", b"ANGT")[0] == O.ORC_PANIC                    # la[index[..]] out of range
    assert O.align(nw, b"", b"ACG")[1] == [(0, 0, 0, 3, -3), (0, 0, 3, 3, 0)]
    fit = O.Scoring(2, False, M, 0, DNA_IDX)
    assert O.align(fit, b"ACGT", b"")[0] == O.ORC_PANIC                       # qSeq[-1]
    assert O.align(fit, b"ACNT", b"ACG")[0::2] == (O.ORC_ILLEGAL_REF, 2)
    assert O.align(fit, b"", b"ACG")[1] == [(0, 0, 3, 3, 0)]
    for kind in (1, 2):
        aff = O.Scoring(kind, True, M, -2, DNA_IDX)
        assert O.align(aff, b"", b"ACG")[0] == O.ORC_PANIC
        assert O.align(aff, b"ACG", b"")[0] == O.ORC_PANIC
    assert O.align(O.Scoring(2, True, M, -2, DNA_IDX), b"ACNT", b"ACG")[0::2] == (O.ORC_ILLEGAL_REF, 2)
    assert O.align(O.Scoring(1, True, M, -2, DNA_IDX), b"ACNT", b"ACG")[0] == O.ORC_PANIC


def test_case_insensitive_index():
    M = gv.match(-1, 2, -1)
    sc = O.Scoring(0, False, M, 0, DNA_IDX)
    assert O.align(sc, b"acacacta", b"AGCACACA")[1] == O.align(sc, b"ACACACTA", b"agcacaca")[1]
