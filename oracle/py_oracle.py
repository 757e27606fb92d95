"""Second, independent restatement of the six aligners in pure Python -- TEST INFRASTRUCTURE.

Written separately from oracle/align_oracle.c (2-D tables, candidate lists instead of an
if-chain, Python's unbounded ints with an explicit minInt sentinel) so that the two
restatements pin each other where the reference has no test of its own (QLetters, Protein,
errors, empty segments).  Small cases only; it follows the same reference lines:
align/sw_letters.go:68-193, nw_letters.go:75-196, fitted_letters.go:70-201,
sw_affine_letters.go:85-292, nw_affine_letters.go:98-332, fitted_affine_letters.go:93-308.
"""
from __future__ import annotations

SW, NW, FIT = 0, 1, 2
OK, ILLEGAL_REF, ILLEGAL_QRY, PANIC, NO_PATH = 0, 1, 2, 3, 4
MININT = -(1 << 63)
M_, U_, L_ = 0, 1, 2


class _Panic(Exception):
    pass


class _Illegal(Exception):
    def __init__(self, which, pos):
        self.which, self.pos = which, pos


def _wrap(v: int) -> int:
    """Go's int arithmetic wraps at 64 bits."""
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _sat_add(a: int, b: int) -> int:
    return MININT if (a == MININT or b == MININT) else _wrap(a + b)  # align/align.go:92-97


class _Mat:
    def __init__(self, matrix):
        self.let = len(matrix)
        self.flat = [v for row in matrix for v in row]

    def at(self, k: int) -> int:
        if k < 0 or k >= len(self.flat):
            raise _Panic()
        return self.flat[k]


def align(kind, affine, matrix, gap_open, index, ref: bytes, qry: bytes):
    """-> (status, segments as 5-tuples in final order, err_pos)"""
    try:
        segs = (_affine if affine else _linear)(kind, _Mat(matrix), gap_open, index, ref, qry)
        return OK, segs, -1
    except _Illegal as e:
        return (ILLEGAL_REF if e.which == "r" else ILLEGAL_QRY), [], e.pos
    except _Panic:
        return PANIC, [], -1


def _letters(index, ref, qry, i, j):
    rv, qv = index[ref[i - 1]], index[qry[j - 1]]
    if rv < 0:
        raise _Illegal("r", i - 1)
    if qv < 0:
        raise _Illegal("q", j - 1)
    return rv, qv


def _linear(kind, mat, _go, index, ref, qry):
    let = mat.let
    rows, cols = len(ref) + 1, len(qry) + 1
    T = [[0] * cols for _ in range(rows)]
    if kind != SW:
        for j in range(1, cols):
            T[0][j] = _wrap(T[0][j - 1] + mat.at(index[qry[j - 1]]))
    if kind == NW:
        for i in range(1, rows):
            T[i][0] = _wrap(T[i - 1][0] + mat.at(index[ref[i - 1]] * let))
    best = (0, 0, 0)
    for i in range(1, rows):
        for j in range(1, cols):
            rv, qv = _letters(index, ref, qry, i, j)
            cand = (T[i - 1][j - 1] + mat.at(rv * let + qv), T[i - 1][j] + mat.at(rv * let), T[i][j - 1] + mat.at(qv))
            s = max(cand)
            if kind == SW:
                if s > 0:
                    if s >= best[0] and s == cand[0]:
                        best = (s, i, j)
                else:
                    s = 0
            T[i][j] = s
    if kind == SW:
        _, i, j = best
    elif kind == NW:
        i, j = rows - 1, cols - 1
    else:
        j = cols - 1
        if j - 1 < 0:
            raise _Panic()
        qend = index[qry[j - 1]]
        i, top = 0, MININT
        for y in range(1, rows):
            rv = index[ref[y - 1]]
            if rv < 0:
                continue
            if T[y][cols - 1] >= top and mat.at(rv * let + qend) >= 0:
                i, top = y, T[y][cols - 1]
    end_i, end_j, score, last, out = i, j, 0, "d", []
    while i > 0 and j > 0:
        rv, qv = index[ref[i - 1]], index[qry[j - 1]]
        here = T[i][j]
        if kind == SW and here == 0:
            break
        corner = (i == rows - 1 and j == cols - 1)
        options = [("d", T[i - 1][j - 1], mat.at(rv * let + qv)), ("u", T[i - 1][j], mat.at(rv * let)),
                   ("l", T[i][j - 1], mat.at(qv))]
        for move, prev, w in options:
            if here == prev + w:
                break
        else:
            raise _Panic()
        guarded = move != "d" and kind != SW
        if last != move and not (guarded and corner):
            out.append((i, end_i, j, end_j, score))
            end_i, end_j, score = i, j, 0
        score += here - prev
        if move != "l":
            i -= 1
        if move != "u":
            j -= 1
        last = move
    out.append((i, end_i, j, end_j, score))
    if kind == NW and i != j:
        out.append((0, i, 0, j, T[i][j]))
    return out[::-1]


def _affine(kind, mat, go, index, ref, qry):
    let = mat.let
    rows, cols = len(ref) + 1, len(qry) + 1
    T = [[[0, 0, 0] for _ in range(cols)] for _ in range(rows)]
    if kind != SW:
        T[0][0] = [0, MININT, MININT]
        if cols < 2:
            raise _Panic()  # qSeq[0]
        T[0][1] = [MININT, MININT, _wrap(go + mat.at(index[qry[0]]))]
        for j in range(2, cols):
            T[0][j] = [MININT, MININT, _wrap(T[0][j - 1][L_] + mat.at(index[qry[j - 1]]))]
        if rows < 2:
            raise _Panic()  # table[c] / rSeq[0]
        if kind == NW:
            T[1][0] = [MININT, _wrap(go + mat.at(index[ref[0]] * let)), MININT]
            for i in range(2, rows):
                T[i][0] = [MININT, _wrap(T[i - 1][0][U_] + mat.at(index[ref[i - 1]] * let)), MININT]
        else:
            for i in range(1, rows):
                T[i][0] = [MININT, 0, MININT]
    best = (0, 0, 0)
    for i in range(1, rows):
        for j in range(1, cols):
            rv, qv = _letters(index, ref, qry, i, j)
            sub, er, eq = mat.at(rv * let + qv), mat.at(rv * let), mat.at(qv)
            d = T[i - 1][j - 1]
            if kind == SW:
                s = max(d)
                matched = s == d[M_]
                s += sub
                if s > 0:
                    if s >= best[0] and matched:
                        best = (s, i, j)
                else:
                    s = 0
                T[i][j] = [s, max(0, T[i - 1][j][M_] + go + er, T[i - 1][j][U_] + er),
                           max(0, T[i][j - 1][M_] + go + eq, T[i][j - 1][L_] + eq)]
            else:
                T[i][j] = [_wrap(max(d) + sub),
                           max(_sat_add(T[i - 1][j][M_], go + er), _sat_add(T[i - 1][j][U_], er)),
                           max(_sat_add(T[i][j - 1][M_], go + eq), _sat_add(T[i][j - 1][L_], eq))]
    layer = M_
    if kind == SW:
        _, i, j = best
    elif kind == NW:
        i, j = rows - 1, cols - 1
        top = T[i][j][M_]
        for k in (U_, L_):
            if T[i][j][k] > top:
                top, layer = T[i][j][k], k
    else:
        j, i, top = cols - 1, 0, MININT
        for y in range(1, rows):
            if T[y][cols - 1][M_] >= top:
                i, top = y, T[y][cols - 1][M_]
    end_i, end_j, score, last, out = i, j, 0, "d", []
    while i > 0 and j > 0:
        rv, qv = index[ref[i - 1]], index[qry[j - 1]]
        sub, er, eq = mat.at(rv * let + qv), mat.at(rv * let), mat.at(qv)
        here = T[i][j][layer]
        if kind == SW and here == 0:
            break
        up, lf, dg = T[i - 1][j], T[i][j - 1], T[i - 1][j - 1]
        options = [("u", U_, up[U_], er), ("l", L_, lf[L_], eq), ("u", M_, up[M_], go + er), ("l", M_, lf[M_], go + eq)]
        diag = [("d", M_, dg[M_], sub), ("d", U_, dg[U_], sub), ("d", L_, dg[L_], sub)]
        options += diag if kind == SW else diag[1:] + diag[:1]
        for move, nl, prev, w in options:
            if here == _wrap(prev + w):
                break
        else:
            raise _Panic()
        corner = (i == rows - 1 and j == cols - 1)
        if last != move and (move == "d" or not corner):
            out.append((i, end_i, j, end_j, score))
            end_i, end_j, score = i, j, 0
        score += here - prev
        if move != "l":
            i -= 1
        if move != "u":
            j -= 1
        layer, last = nl, move
    out.append((i, end_i, j, end_j, score))
    if kind == NW and i != j:
        out.append((0, i, 0, j, T[i][j][L_ if i == 0 else U_]))
    return out[::-1]
