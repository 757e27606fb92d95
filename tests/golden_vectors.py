"""Golden vectors of the reference's own tests (G1..G8) and the derived
regression anchors of SURVEY.md appendix B (Q1, Q2, D1..D8; not reference-pinned).

G1/G2  align/sw_example_test.go:14-70      G3 align/sw_example_test.go:72-108
G4/G5  align/nw_example_test.go:14-83      G6/G7 align/fitted_example_test.go:14-83
G8     align/align_test.go:28-51 (TestIssue58)
"""


def match(gap, m, x, n=5, zero00=True):
    M = [[(gap if (i == 0 or j == 0) else (m if i == j else x)) for j in range(n)] for i in range(n)]
    if zero00:
        M[0][0] = 0
    return M


_NW4 = [[0, -5, -5, -5, -5], [-5, 10, -3, -1, -4], [-5, -3, 9, -5, 0], [-5, -1, -5, 7, -3], [-5, -4, 0, -3, 8]]
_PM1 = match(-1, 1, -1)

# (name, kind, affine, matrix, gap_open, ref, query, fmt.Sprint(aln), Format rows or None)
GOLDEN = [
    ("G1", 0, False, match(-1, 2, -1), 0, "ACACACTA", "AGCACACA",
     "[[0,1)/[0,1)=2 -/[1,2)=-1 [1,6)/[2,7)=10 [6,7)/-=-1 [7,8)/[7,8)=2]", ("A-CACACTA", "AGCACAC-A")),
    ("G2", 0, False, match(0, 2, -1), 0, "AAAATTTAAAA", "AAAAGGGAAAA",
     "[[0,4)/[0,4)=8 -/[4,7)=0 [4,7)/-=0 [7,11)/[7,11)=8]", ("AAAA---TTTAAAA", "AAAAGGG---AAAA")),
    ("G3", 0, True, _PM1, -5, "ATAGGAAG", "ATTGGCAATG", "[[0,7)/[0,7)=3]", ("ATAGGAA", "ATTGGCA")),
    ("G4", 1, False, _NW4, 0, "AGACTAGTTA", "GACAGACG",
     "[[0,1)/-=-5 [1,4)/[0,3)=26 [4,5)/-=-5 [5,10)/[3,8)=12]", ("AGACTAGTTA", "-GAC-AGACG")),
    ("G5", 1, True, _PM1, -5, "ATAGGAAG", "ATTGGCAATG",
     "[[0,7)/[0,7)=3 -/[7,9)=-7 [7,8)/[9,10)=1]", ("ATAGGAA--G", "ATTGGCAATG")),
    ("G6", 2, False, _NW4, 0, "GTTGACAGACTAGATTCACG", "GACAGACGA",
     "[[3,10)/[0,7)=62 [10,12)/-=-10 [12,14)/[7,9)=17]", ("GACAGACTAGA", "GACAGAC--GA")),
    ("G7", 2, True, _PM1, -5, "ATTGGCAATGA", "ATAGGAA", "[[0,7)/[0,7)=3]", ("ATTGGCA", "ATAGGAA")),
    ("G8", 1, True, _PM1, -1, "GCTCACTAAAAACACAATCTACAACAGACGTTGCACTAACACTGTAATTGCCTTTAGTCC", "ACTGCGTA",
     "[[0,4)/-=-5 [4,7)/[0,3)=3 [7,32)/-=-26 [32,34)/[3,5)=2 [34,43)/-=-10 [43,46)/[5,8)=3 [46,60)/-=-15]",
     None),
]

_P_REF = "EMTTEIYYVQWVSWRIAYDELERASMHPKNNPTDMDVVLLRFWLANNR"
_P_QRY = "EMTTTIYYVQWVSWRIAYDELERISRHPKNPTDMDVVKFRFWLANR"
_D_REF = "TTAGGTATGTCTTAGTGACTCTAAATACCAAGGCAGTCCTCGATCCGTTCCTAATAAGGAATGGTGATTC"
_D_QRY = "TCGACTTCAAATACAAGGCCAGTGCTTCGATCCTTCCTAAT"
_D5 = ("[[15,16)/[0,1)=2 -/[1,2)=-1 [16,19)/[2,5)=6 -/[5,6)=-1 [19,21)/[6,8)=4 [21,22)/-=-1 [22,27)/[8,13)=10 "
       "[27,28)/-=-1 [28,33)/[13,18)=10 -/[18,19)=-1 [33,37)/[19,23)=8 -/[23,24)=-1 [37,46)/[24,33)=15 "
       "[46,47)/-=-1 [47,55)/[33,41)=16]")

# (name, kind, affine, "dna"|"protein", matrix or None (protein: BLOSUM62 gap -1), gap_open, ref, qry, expected)
DERIVED = [
    ("Q1", 0, True, "dna", match(0, 2, -1), 0, "TGCA", "TACC", "[[0,3)/[0,3)=3 -/[3,4)=0 -/[4,4)=0]"),
    ("Q2", 2, True, "dna", match(0, 2, -1), 0, "CTCGCCC", "TGAA",
     "[[1,2)/[0,1)=2 [2,3)/-=0 [3,5)/[1,3)=1 -/[3,4)=0 -/[4,4)=0]"),
    ("D1", 1, True, "protein", None, -10, _P_REF, _P_QRY,
     "[[0,30)/[0,30)=149 [30,31)/-=-11 [31,46)/[30,45)=71 [46,47)/-=-11 [47,48)/[45,46)=5]"),
    ("D2", 0, True, "protein", None, -10, _P_REF, _P_QRY, "[[0,30)/[0,30)=149 [30,31)/-=-11 [31,47)/[30,46)=71]"),
    ("D3", 2, True, "dna", _PM1, -5, _D_REF, _D_QRY, "[[14,55)/[0,41)=13]"),
    ("D4", 0, True, "dna", _PM1, -5, _D_REF, _D_QRY, "[[16,55)/[2,41)=15]"),
    ("D5", 0, False, "dna", match(-1, 2, -1), 0, _D_REF, _D_QRY, _D5),
    ("D6", 2, False, "dna", match(-1, 2, -1), 0, _D_REF, _D_QRY, _D5),
    ("D8", 1, True, "dna", _PM1, -5, _D_REF, _D_QRY, "[[0,2)/[0,2)=0 [2,16)/-=-19 [16,55)/[2,41)=15 [55,70)/-=-20]"),
]
