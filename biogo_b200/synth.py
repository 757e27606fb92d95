"""Deterministic synthetic workloads for the five BASELINE.json configs.

Everything derives from splitmix64 streams (seed = 0xB10600 + config number, SURVEY.md
section 8d), vectorised with numpy so a 125k-pair shard is generated in seconds.  The
generator is data only: it produces the packed batch layout of the C ABI
(one letters buffer + byte offsets + lengths) and the scoring parameters of each config.
"""
from __future__ import annotations

import numpy as np

from . import alphabet, matrix

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, n: int, start: int = 0) -> np.ndarray:
    """n outputs of splitmix64 seeded with ``seed`` (counter-based, so any slice of the
    stream can be produced independently: element k uses state seed + (k+1)*gamma)."""
    with np.errstate(over="ignore"):
        k = np.arange(start + 1, start + n + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


class _Stream:
    """Sequential consumer of one splitmix64 stream."""

    def __init__(self, seed: int):
        self.seed, self.pos = seed, 0

    def u64(self, n: int) -> np.ndarray:
        out = splitmix64(self.seed, n, self.pos)
        self.pos += n
        return out

    def uniform(self, n: int) -> np.ndarray:
        return (self.u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))

    def randint(self, n: int, hi: int) -> np.ndarray:
        return (self.u64(n) % np.uint64(hi)).astype(np.int64)


def _mutate(st: _Stream, src: np.ndarray, out_len: int, p_sub: float, p_ins: float, p_del: float,
            letters: np.ndarray) -> np.ndarray:
    """Row-wise mutation of ``src`` [pairs, L] (substitute / insert after / delete), then
    truncate or pad with fresh random letters to exactly ``out_len`` columns."""
    n, L = src.shape
    u = st.uniform(n * L).reshape(n, L)
    sub_mask = u < p_sub
    del_mask = (u >= p_sub) & (u < p_sub + p_del)
    ins_mask = st.uniform(n * L).reshape(n, L) < p_ins
    rnd = letters[st.randint(n * L, len(letters))].reshape(n, L)
    rnd2 = letters[st.randint(n * L, len(letters))].reshape(n, L)
    base = np.where(sub_mask, rnd, src)
    # interleave: position 2k = (possibly deleted) source letter, 2k+1 = optional insertion
    wide = np.empty((n, 2 * L), dtype=np.uint8)
    keep = np.empty((n, 2 * L), dtype=bool)
    wide[:, 0::2] = base
    wide[:, 1::2] = rnd2
    keep[:, 0::2] = ~del_mask
    keep[:, 1::2] = ins_mask
    dest = np.cumsum(keep, axis=1) - 1
    pad = letters[st.randint(n * out_len, len(letters))].reshape(n, out_len)
    out = pad.copy()
    ok = keep & (dest < out_len)
    rows = np.broadcast_to(np.arange(n)[:, None], wide.shape)
    out[rows[ok], dest[ok]] = wide[ok]
    return out


class Workload:
    """A packed batch plus the aligner parameters it is meant for."""

    def __init__(self, name, kind, affine, mat, gap_open, alpha, letters, stride, ref_off, ref_len, qry_off,
                 qry_len):
        self.name, self.kind, self.affine = name, kind, affine
        self.matrix, self.gap_open, self.alpha = mat, gap_open, alpha
        self.letters, self.stride = letters, stride
        self.ref_off, self.ref_len, self.qry_off, self.qry_len = ref_off, ref_len, qry_off, qry_len

    @property
    def n(self) -> int:
        return len(self.ref_len)

    @property
    def cells(self) -> int:
        return int((self.ref_len.astype(np.int64) * self.qry_len.astype(np.int64)).sum())

    def subset(self, idx) -> "Workload":
        idx = np.asarray(idx)
        return Workload(self.name, self.kind, self.affine, self.matrix, self.gap_open, self.alpha, self.letters,
                        self.stride, self.ref_off[idx], self.ref_len[idx], self.qry_off[idx], self.qry_len[idx])

    def pair(self, p: int):
        s = self.stride
        r = self.letters[self.ref_off[p]: self.ref_off[p] + self.ref_len[p] * s: s].tobytes()
        q = self.letters[self.qry_off[p]: self.qry_off[p] + self.qry_len[p] * s: s].tobytes()
        return r, q


_DNA = np.frombuffer(b"ACGT", dtype=np.uint8)
_AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)


def _fixed_shape(st, n_pairs, ref_n, qry_m, letters, p_sub, p_ins, p_del, substring, chunk=20000):
    """Pairs of one fixed shape, packed as [ref | qry] per pair, stride 1."""
    per = ref_n + qry_m
    buf = np.empty(n_pairs * per, dtype=np.uint8)
    view = buf.reshape(n_pairs, per)
    for lo in range(0, n_pairs, chunk):
        hi = min(n_pairs, lo + chunk)
        k = hi - lo
        ref = letters[st.randint(k * ref_n, len(letters))].reshape(k, ref_n)
        view[lo:hi, :ref_n] = ref
        if substring:
            L = min(ref_n, qry_m + qry_m // 8 + 8)
            off = st.randint(k, ref_n - L + 1)
            cols = off[:, None] + np.arange(L)[None, :]
            src = np.take_along_axis(ref, cols, axis=1)
        else:
            src = ref
        view[lo:hi, ref_n:] = _mutate(st, src, qry_m, p_sub, p_ins, p_del, letters)
    ref_off = np.arange(n_pairs, dtype=np.int64) * per
    return (buf, ref_off, np.full(n_pairs, ref_n, np.int32), ref_off + ref_n, np.full(n_pairs, qry_m, np.int32))


def config(num: int, n_pairs: int | None = None, shard: int = 0) -> Workload:
    """The synthetic batch of BASELINE.json config ``num`` (1..5).

    ``n_pairs`` bounds the batch (default: the config's full size); ``shard`` offsets the
    seed so that ranks of a multi-GPU run draw independent shards of the same distribution.
    """
    seed = 0xB10600 + num + (shard << 20)
    st = _Stream(seed)
    dna = alphabet.DNAgapped
    if num == 1:
        n = 100_000 if n_pairs is None else n_pairs
        mat = matrix.Match(dna, -1, 2, -1)
        b = _fixed_shape(st, n, 500, 150, _DNA, 0.04, 0.005, 0.005, True)
        return Workload("cfg1 SW DNA 150x500", 0, False, mat, 0, dna, b[0], 1, *b[1:])
    if num == 2:
        n = 1_000_000 if n_pairs is None else n_pairs
        mat = matrix.Match(dna, -1, 1, -1)
        b = _fixed_shape(st, n, 1000, 250, _DNA, 0.04, 0.005, 0.005, True)
        return Workload("cfg2 SWAffine DNA 250x1000", 0, True, mat, -5, dna, b[0], 1, *b[1:])
    if num == 3:
        n = 1_000_000 if n_pairs is None else n_pairs
        mat = matrix.with_gap(matrix.BLOSUM62, -1)
        b = _fixed_shape(st, n, 300, 300, _AA, 0.30, 0.02, 0.02, False)
        return Workload("cfg3 NWAffine Protein BLOSUM62 300x300", 1, True, mat, -10, alphabet.Protein, b[0], 1,
                        *b[1:])
    if num == 4:
        n = 128 if n_pairs is None else n_pairs
        mat = matrix.Match(dna, -1, 1, -1)
        b = _fixed_shape(st, n, 12000, 10000, _DNA, 0.08, 0.03, 0.03, True, chunk=16)
        return Workload("cfg4 FittedAffine DNA 10kx12k", 2, True, mat, -5, dna, b[0], 1, *b[1:])
    if num == 5:
        n = 200_000 if n_pairs is None else n_pairs
        mat = matrix.Match(dna, -1, 1, -1)
        # read length log-uniform in [100, 5000]; window = min(5000, 2L); QLetters (Q = 40)
        u = st.uniform(n)
        L = np.floor(100.0 * np.exp(u * np.log(50.0))).astype(np.int64).clip(100, 5000)
        Wn = np.minimum(5000, 2 * L)
        per = (Wn + L) * 2
        ends = np.cumsum(per)
        ref_off = ends - per
        buf = np.full(int(ends[-1]) if n else 0, 40, dtype=np.uint8)
        order = np.argsort(L, kind="stable")
        # generate per length-class so the mutation kernel stays vectorised
        bounds = np.flatnonzero(np.diff(L[order])) + 1
        for grp in np.split(order, bounds):
            l, w = int(L[grp[0]]), int(Wn[grp[0]])
            k = len(grp)
            ref = _DNA[st.randint(k * w, 4)].reshape(k, w)
            Ls = min(w, l + l // 8 + 8)
            off = st.randint(k, w - Ls + 1)
            src = np.take_along_axis(ref, off[:, None] + np.arange(Ls)[None, :], axis=1)
            q = _mutate(st, src, l, 0.04, 0.005, 0.005, _DNA)
            for row, p in enumerate(grp):
                o = int(ref_off[p])
                buf[o: o + 2 * w: 2] = ref[row]
                buf[o + 2 * w: o + 2 * (w + l): 2] = q[row]
        return Workload("cfg5 SWAffine QLetters 100bp-5kb", 0, True, mat, -5, dna, buf, 2, ref_off,
                        Wn.astype(np.int32), ref_off + 2 * Wn, L.astype(np.int32))
    raise ValueError("config number must be 1..5")
