"""The reference's benchmark input (align/crsp_test.go:7-70: two platypus crsp records, 1812 and
1797 nt, mixed case) through all six aligners.  The reference holds NO expected output for it (its
benchmarks only time it, align/align_test.go:53-140), so this is an input fixture: the CUDA path must
equal the oracle on it, and the oracle must equal the independent Python restatement on a prefix.
The benchmark scorings put the gap penalties in the LAST row/column (an older matrix layout) while
DNAgapped's gap letter has index 0 -- legal input with non-uniform "gap" entries, which the packed
16-bit path must refuse and the exact 32-bit path must handle."""
import os

import numpy as np
import pytest

import helpers
from biogo_b200 import alphabet, matrix, seqio
from oracle import oracle as O
from oracle import py_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
BENCH_SW = [[2, -1, -1, -1, -1], [-1, 2, -1, -1, -1], [-1, -1, 2, -1, -1], [-1, -1, -1, 2, -1], [-1, -1, -1, -1, 0]]
BENCH_NW = [[10, -3, -1, -4, -5], [-3, 9, -5, 0, -5], [-1, -5, 7, -3, -5], [-4, 0, -3, 8, -5], [-4, -4, -4, -4, 0]]
# (kind, affine, matrix, gap_open): the four reference benchmarks, then every kind with a Match matrix
CASES = [
    ("BenchmarkSWAlign", 0, False, BENCH_SW, 0),
    ("BenchmarkNWAlign", 1, False, BENCH_NW, 0),
    ("BenchmarkSWAffineAlign", 0, True, BENCH_SW, -5),
    ("BenchmarkNWAffineAlign", 1, True, BENCH_NW, -10),
] + [(helpers.KIND_NAMES[(k, a)], k, a, matrix.Match(alphabet.DNAgapped, -1, 2 if not a else 1, -1), -5 if a else 0)
     for k, a in helpers.KINDS]


def _records():
    f = seqio.parse_fasta(open(os.path.join(HERE, "golden", "crsp.fa"), "rb").read(), lib_path=helpers.EMU_LIB)
    assert len(f) == 2 and [int(x) for x in f.seq_len] == [1812, 1797]
    assert f.names[0] == "platypusX_curated" and f.descs[0].startswith("[gene=crspx]")
    return f.raw(0), f.raw(1)


IDX = alphabet.DNAgapped.LetterIndex()


def test_oracle_and_python_restatement_agree_on_a_prefix(built):
    r, q = _records()
    assert any(c in b"acgt" for c in r) and any(c in b"ACGT" for c in r)      # mixed case
    pk = helpers.pack([(r[:260], q[:240]), (q[100:330], r[90:300])])
    for name, kind, affine, M, go in CASES:
        orc = helpers.run_oracle(kind, affine, M, go, IDX, pk)
        for p in range(2):
            ref, qry = [(r[:260], q[:240]), (q[100:330], r[90:300])][p]
            st, want, _ = py_oracle.align(kind, affine, M, go, IDX, ref, qry)
            got = orc[3][orc[2][p]:orc[2][p + 1]]
            assert st == 0 == int(orc[0][p])
            assert [tuple(int(v) for v in s) for s in got] == [tuple(s) for s in want], (name, p)


@pytest.mark.parametrize("case", CASES[:4] + CASES[7:], ids=[c[0] for c in CASES[:4] + CASES[7:]])
def test_full_pair_through_the_emulated_kernels(emu_ctx, case):
    """CPU tier: the unmodified kernels under the SIMT emulator on the whole 1812 x 1797 pair (both
    orders), the benchmark scorings and the affine kinds."""
    name, kind, affine, M, go = case
    r, q = _records()
    pk = helpers.pack([(r, q), (q, r)])
    res = helpers.run_lib(emu_ctx, kind, affine, M, go, IDX, 5, pk)
    helpers.compare(res, helpers.run_oracle(kind, affine, M, go, IDX, pk), name)
    if "NW" in name and name.startswith("Benchmark"):
        assert res.stats["pairs_fast16"] == 0     # non-uniform gap entries: exact path only
    if "SW" in name and name.startswith("Benchmark"):
        assert res.stats["pairs_fast16"] == 2     # its gap entries happen to be uniform over a, c, g, t


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_full_pair_on_the_gpu(gpu_ctx, case):
    name, kind, affine, M, go = case
    r, q = _records()
    pairs = [(r, q), (q, r), (r, r[:900]), (q[500:], r)]
    for stride in (1, 2):
        pk = helpers.pack(pairs, stride)
        res = helpers.run_lib(gpu_ctx, kind, affine, M, go, IDX, 5, pk, stride)
        helpers.compare(res, helpers.run_oracle(kind, affine, M, go, IDX, pk, stride, 8), f"{name} stride {stride}")
