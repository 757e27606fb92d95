"""Scoring tables of biogo's align/matrix package that the configs use.

All tables are indexed to match the alphabets of ``biogo_b200.alphabet`` (gap letter at
index 0).  As in the reference (align/matrix/matrices.go:10-13) the gap row/column of the
published tables is all zero and the protein letter J is present but undefined (zero).

* ``Match``    align/matrix/generated.go:10-31
* ``NUC_4``    align/matrix/matrices.go:28-35  (rows/cols: - A C G T)
* ``BLOSUM62`` align/matrix/matrices.go:842-870 (rows/cols: - A B C D E F G H I J K L M N P Q R S T V W X Y Z *)

BLOSUM62 is rebuilt here from the standard NCBI table (order ARNDCQEGHILKMFPSTWYVBZX*)
rather than restated cell by cell; tests/test_matrix.py pins a checksum of the result.
"""
from __future__ import annotations


def Match(alpha, gap: int, match: int, mismatch: int):
    """matrix.Match: ``gap`` in row/col of the gap letter (including [g][g]), ``match`` on the
    diagonal, ``mismatch`` elsewhere."""
    n = alpha.Len()
    g = alpha.IndexOf(alpha.Gap())
    out = []
    for i in range(n):
        row = []
        for j in range(n):
            if i == g or j == g:
                row.append(gap)
            elif i == j:
                row.append(match)
            else:
                row.append(mismatch)
        out.append(row)
    return out


NUC_4 = [
    #  -   A   C   G   T
    [0, 0, 0, 0, 0],
    [0, 5, -4, -4, -4],
    [0, -4, 5, -4, -4],
    [0, -4, -4, 5, -4],
    [0, -4, -4, -4, 5],
]

_NCBI_ORDER = "ARNDCQEGHILKMFPSTWYVBZX*"
_NCBI_BLOSUM62 = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""

_PROTEIN_ORDER = "-ABCDEFGHIJKLMNPQRSTVWXYZ*"


def _reorder(ncbi_text: str):
    rows = [[int(x) for x in line.split()] for line in ncbi_text.strip().splitlines()]
    pos = {ch: k for k, ch in enumerate(_NCBI_ORDER)}
    out = []
    for a in _PROTEIN_ORDER:
        row = []
        for b in _PROTEIN_ORDER:
            if a in pos and b in pos:
                row.append(rows[pos[a]][pos[b]])
            else:
                row.append(0)  # gap and J rows/columns are zero in align/matrix
        out.append(row)
    return out


BLOSUM62 = _reorder(_NCBI_BLOSUM62)


def with_gap(matrix, gap: int):
    """Copy of ``matrix`` with row 0 / column 0 set to ``gap`` and [0][0] = 0 -- how the
    BASELINE configs turn a published table (zero gap row) into a usable one."""
    out = [list(r) for r in matrix]
    for k in range(len(out)):
        out[0][k] = gap
        out[k][0] = gap
    out[0][0] = 0
    return out
