"""Host-side mirror of the Go surface: validation order, sentinel errors, panics, Format.
The DP behind it runs in the emulated kernels here (no GPU in this tier)."""
import pytest

import golden_vectors as gv
import helpers
from biogo_b200 import align, alphabet, feat, seq


@pytest.fixture(autouse=True)
def _use_emulated_library(built):
    align.set_library_path(helpers.EMU_LIB)
    yield
    align.set_library_path(None)


def S(b, alpha=alphabet.DNAgapped):
    return seq.Seq(Seq=alphabet.BytesToLetters(b), Alpha=alpha)


def Q(b, alpha=alphabet.DNAgapped):
    return seq.QSeq(Seq=alphabet.QLetters(b, 40), Alpha=alpha)


M = gv.match(-1, 1, -1)
ALL = [align.SW(M), align.NW(M), align.Fitted(M), align.SWAffine(Matrix=M, GapOpen=-5),
       align.NWAffine(Matrix=M, GapOpen=-5), align.FittedAffine(Matrix=M, GapOpen=-5)]


@pytest.mark.parametrize("a", ALL, ids=lambda a: type(a).__name__)
def test_wrapper_validation_order(a):
    # align/sw.go:23-48: no alphabet -> mismatched alphabets -> not gapped -> types
    assert a.Align(S(b"ACGT", None), S(b"ACGT"))[1] is align.ErrNoAlphabet
    assert a.Align(S(b"ACGT"), S(b"ACGT", alphabet.DNAredundant))[1] is align.ErrMismatchedAlphabets
    assert a.Align(S(b"ACGT", alphabet.DNA), S(b"ACGT", alphabet.DNA))[1] is align.ErrNotGappedAlphabet
    assert a.Align(S(b"ACGT"), Q(b"ACGT"))[1] is align.ErrMismatchedTypes
    assert a.Align(Q(b"ACGT"), S(b"ACGT"))[1] is align.ErrMismatchedTypes

    class Other:
        def Alphabet(self):
            return alphabet.DNAgapped

        def Slice(self):
            return [1, 2, 3]

    assert a.Align(Other(), S(b"ACGT"))[1] is align.ErrTypeNotHandled


def test_matrix_checks():
    small = [[0, -1], [-1, 1]]
    for cls in (align.SW, align.NW):
        assert cls(small).Align(S(b"ACGT"), S(b"ACGT"))[1] == align.ErrMatrixWrongSize(Size=2, Len=5)
    assert align.SWAffine(Matrix=small, GapOpen=-1).Align(S(b"A"), S(b"A"))[1] == align.ErrMatrixWrongSize(2, 5)
    ragged = [row[:] for row in M]
    ragged[2] = ragged[2][:4]
    for a in (align.SW(ragged), align.Fitted(ragged), align.NWAffine(Matrix=ragged, GapOpen=-1)):
        assert a.Align(S(b"ACGT"), S(b"ACGT"))[1] is align.ErrMatrixNotSquare
    # Fitted has no size check (align/fitted_letters.go:71-78): a larger matrix is fine everywhere
    big = gv.match(-1, 1, -1, 7)
    assert align.Fitted(big).Align(S(b"ACGT"), S(b"ACG"))[1] is None


def test_illegal_letters_and_panics():
    aln, err = align.SW(M).Align(S(b"ACGN"), S(b"AXGT"))
    assert aln is None and str(err) == "align: illegal letter 'X' at position 1 in qSeq"
    aln, err = align.SWAffine(Matrix=M, GapOpen=-2).Align(S(b"NCGT"), S(b"AXGT"))
    assert str(err) == "align: illegal letter 'N' at position 0 in rSeq"
    aln, err = align.FittedAffine(Matrix=M, GapOpen=-2).Align(S(b"ACNT"), S(b"ACG"))
    assert str(err) == "align: illegal letter 'N' at position 2 in rSeq"
    with pytest.raises(align.GoPanic):
        align.NW(M).Align(S(b"ACNT"), S(b"ACG"))
    with pytest.raises(align.GoPanic):
        align.NWAffine(Matrix=M, GapOpen=-2).Align(S(b""), S(b"ACG"))
    with pytest.raises(align.GoPanic):
        align.Fitted(M).Align(S(b"ACG"), S(b""))
    aln, err = align.SW(M).Align(S(b""), S(b""))
    assert err is None and feat.Sprint(aln) == "[-/[0,0)=0]"


@pytest.mark.parametrize("vec", gv.GOLDEN, ids=[v[0] for v in gv.GOLDEN])
def test_examples_and_format(vec):
    name, kind, affine, Mx, go, ref, qry, want, fmt = vec
    cls = {(0, False): align.SW, (1, False): align.NW, (2, False): align.Fitted, (0, True): align.SWAffine,
           (1, True): align.NWAffine, (2, True): align.FittedAffine}[(kind, affine)]
    a = cls(Matrix=Mx, GapOpen=go) if affine else cls(Mx)
    for mk in (S, Q):
        sa, sb = mk(ref.encode()), mk(qry.encode())
        aln, err = a.Align(sa, sb)
        assert err is None and feat.Sprint(aln) == want
        if fmt:
            fa = align.Format(sa, sb, aln, "-")
            assert (str(fa[0]), str(fa[1])) == fmt
            # the batched device Format agrees with the per-alignment one, qualities included
            fb = align.FormatBatch([sa, sa], [sb, sb], [aln, None], "-")
            assert (str(fb[0][0]), str(fb[0][1])) == fmt and (str(fb[1][0]), str(fb[1][1])) == ("", "")
            assert type(fb[0][0]) is type(fa[0]) and getattr(fb[0][0], "raw", None) == getattr(fa[0], "raw", None)
            assert getattr(fb[0][1], "raw", None) == getattr(fa[1], "raw", None)


def test_align_batch_mixes_errors_and_results():
    a = align.SWAffine(Matrix=M, GapOpen=-5)
    refs = [S(b"ATAGGAAG"), S(b"ACGT", None), S(b"ACGN"), S(b"")]
    qrys = [S(b"ATTGGCAATG"), S(b"ACGT"), S(b"ACGT"), S(b"ACG")]
    alns, errs = a.AlignBatch(refs, qrys)
    assert feat.Sprint(alns[0]) == "[[0,7)/[0,7)=3]" and errs[0] is None
    assert errs[1] is align.ErrNoAlphabet and alns[1] is None
    assert str(errs[2]) == "align: illegal letter 'N' at position 3 in rSeq"
    assert feat.Sprint(alns[3]) == "[-/[0,0)=0]"


def test_pair_type():
    p = feat.Pair(0, 3, 2, 5, 7)
    assert str(p) == "[0,3)/[2,5)=7" and p.Score() == 7
    assert p.Features()[0].Len() == 3 and p.Features()[0].Name() == "" and p.Features()[0].Location() is None
    p.Invert()
    assert str(p) == "[2,5)/[0,3)=7"
    assert str(feat.Pair(1, 1, 0, 2, -3)) == "-/[0,2)=-3"
    assert str(feat.Pair(1, 4, 2, 2, -3)) == "[1,4)/-=-3"
