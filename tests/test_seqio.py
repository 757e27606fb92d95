"""FASTA / FASTQ ingest (ba_fasta_parse / ba_fastq_parse).

PINNED on the reference's own reader fixtures, restated under tests/golden/seqio_fixtures.json by
tests/golden/make_seqio_fixtures.py: io/seqio/fasta/fasta_test.go:30-93 (testaln0/testaln1 ->
expectN/expectS) and the nine fqTests of io/seqio/fastq/fastq_test.go:92-399 (ids, letters and Phred
values, including the '@'/'+' quality-start and the truncated-record cases).  Beyond those, a
pure-Python restatement of the readers (io/seqio/fasta/fasta.go:60-113, io/seqio/fastq/fastq.go:62-141)
checks hand-written and random texts.  CPU tier: the parsers are host code inside the library
(emulator build); the GPU tier feeds parsed records straight into ba_align_batch."""
import random

import numpy as np
import pytest

import golden_vectors as gv
import helpers
from biogo_b200 import alphabet, seqio
from oracle import oracle as O

SPACE = b" \t\n\v\f\r"


def ref_fasta(text: bytes):
    """fasta.go:60-113 restated: returns [(name, desc, letters)] or raises ValueError(msg)."""
    out, working = [], None
    for raw in text.split(b"\n"):
        line = raw.strip(SPACE)                       # bytes.TrimSpace after ReadLine
        if not line:
            continue
        if line.startswith(b">"):
            if working is not None:
                out.append(working)
            mark = min([i for i in (line.find(b" "), line.find(b"\t")) if i >= 0], default=-1)
            if mark < 0:
                working = [line[1:], b"", bytearray()]
            else:
                working = [line[1:max(mark, 1)], line[mark + 1:], bytearray()]
        else:
            if working is None:
                raise ValueError('fasta: badly formed line')
            working[2] += b"".join(line.split())      # bytes.Join(bytes.Fields(line), nil)
    if working is not None:
        out.append(working)
    return [(n.decode("latin-1"), d.decode("latin-1"), bytes(s)) for n, d, s in out]


def ref_fastq(text: bytes, off: int):
    """fastq.go:62-141 restated for a whole file: [(name, desc, letters, quals)] or ValueError."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    pos, out = 0, []
    while True:
        state, label, t, seqbuf, line = "id1", b"", None, b"", b""
        eof = False
        while True:
            if pos >= len(lines):
                eof = True
                break
            line = lines[pos].strip(SPACE)
            pos += 1
            if state == "id1" and line[:1] == b"@":
                state, label = "letters", line
                mark = min([i for i in (line.find(b" "), line.find(b"\t")) if i >= 0], default=-1)
                t = (line[1:], b"") if mark < 0 else (line[1:max(mark, 1)], line[mark + 1:])
            elif state == "id2" and line[:1] == b"+":
                state = "quality"
                if len(line) != 1 and label[1:] != line[1:]:
                    raise ValueError("fastq: quality header does not match sequence header")
            elif state == "letters" and len(line) > 0:
                if line[:1] == b"+" and (len(line) == 1 or label[1:] == line[1:]):
                    state = "quality"
                    continue
                state = "id2"
                seqbuf = bytes(c for c in line if c not in b"\t\n\v\f\r \x85\xa0")
            elif state == "quality":
                if len(line) == 0 and len(seqbuf) != 0:
                    continue
                break
        if eof:
            if t is not None and state == "quality":
                line = b""
            elif t is None and state == "id1":
                return out
            else:
                raise ValueError("fastq: unexpected end of file inside a record")
        q = b"".join(line.split())
        if len(q) != len(seqbuf):
            raise ValueError("fastq: sequence/quality length mismatch")
        out.append((t[0].decode("latin-1"), t[1].decode("latin-1"), seqbuf, bytes((c - off) & 0xff for c in q)))
        if eof:
            return out


FASTA_CASES = [
    b">a first record\nACGT\nAC GT\n\n>b\tsecond\r\n  TTTT  \r\n>empty\n>c\nA\n",
    b"",
    b"\n\n   \n",
    b">only\n",
    b"> spaced name\nACGT\n",
    b">x\nAC\tGT\x0bAA\n>y desc  with  blanks \nGG\n",
    b">no_newline_at_end\nACGTACGT",
]
FASTQ_CASES = [
    b"@r1 desc\nACGT\n+\nIIII\n@r2\nAC\n+r2\n!~\n",
    b"@r1\nACGT\n+\nIIII",
    b"",
    b"@r1\n+\n\n@r2\nGG\n+\nII\n",
    b"\n@r1\n  ACGT  \n+\n I I I I \n",
]


@pytest.fixture(scope="module")
def emu_lib(built):
    return helpers.EMU_LIB


@pytest.mark.parametrize("text", FASTA_CASES, ids=[str(k) for k in range(len(FASTA_CASES))])
def test_fasta_cases(emu_lib, text):
    want = ref_fasta(text)
    for q in (False, True):
        f = seqio.parse_fasta(text, qletters=q, lib_path=emu_lib)
        got = [(f.names[k], f.descs[k], f.raw(k)[::2] if q else f.raw(k)) for k in range(len(f))]
        assert got == want
        if q:
            assert all(set(f.raw(k)[1::2]) <= {0} for k in range(len(f)))


def test_fasta_errors_and_random(emu_lib):
    with pytest.raises(seqio.FormatError) as e:
        seqio.parse_fasta(b"\nACGT\n>late\nAC\n", lib_path=emu_lib)
    assert e.value.line == 2 and str(e.value).startswith("fasta: badly formed line")
    rng = random.Random(3)
    for _ in range(60):
        recs = []
        for k in range(rng.randrange(0, 9)):
            head = b">" + bytes(rng.choice(b"abcXYZ_/1") for _ in range(rng.randrange(0, 8)))
            if rng.random() < 0.6:
                head += rng.choice([b" ", b"\t"]) + bytes(rng.choice(b"de sc\t1") for _ in range(rng.randrange(0, 10)))
            body = b""
            for _ in range(rng.randrange(0, 4)):
                body += bytes(rng.choice(b"ACGTN- \t") for _ in range(rng.randrange(0, 30))) + rng.choice([b"\n", b"\r\n", b"\n\n"])
            recs.append(head + rng.choice([b"\n", b"\r\n"]) + body)
        text = b"".join(recs)
        f = seqio.parse_fasta(text, lib_path=emu_lib)
        assert [(f.names[k], f.descs[k], f.raw(k)) for k in range(len(f))] == ref_fasta(text)


@pytest.mark.parametrize("text", FASTQ_CASES, ids=[str(k) for k in range(len(FASTQ_CASES))])
def test_fastq_cases(emu_lib, text):
    for off in (seqio.SANGER, seqio.ILLUMINA_1_3):
        want = ref_fastq(text, off)
        f = seqio.parse_fastq(text, off, lib_path=emu_lib)
        got = [(f.names[k], f.descs[k], f.raw(k)[::2], f.raw(k)[1::2]) for k in range(len(f))]
        assert got == want


def test_fastq_errors(emu_lib):
    for text, msg in ((b"@r1\nACGT\n+\nIII\n", "fastq: sequence/quality length mismatch"),
                      (b"@r1\nACGT\n+other\nIIII\n", "fastq: quality header does not match sequence header"),
                      (b"@r1\nACGT\n", "fastq: unexpected end of file inside a record")):
        with pytest.raises(ValueError) as e1:
            ref_fastq(text, 33)
        with pytest.raises(seqio.FormatError) as e2:
            seqio.parse_fastq(text, lib_path=emu_lib)
        assert str(e1.value) == msg and str(e2.value) == msg


@pytest.mark.gpu
def test_fasta_records_feed_the_aligner(gpu_ctx):
    rng = random.Random(8)
    text, seqs = b"", []
    for k in range(40):
        s = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(20, 400)))
        seqs.append(s)
        text += b">s%d test\n" % k + b"\n".join(s[x:x + 60] for x in range(0, len(s), 60)) + b"\n"
    f = seqio.parse_fasta(text)
    assert [f.raw(k) for k in range(len(f))] == seqs
    pk = f.pairs(range(0, 40, 2), range(1, 40, 2))
    M = gv.match(-1, 1, -1)
    idx = O.build_index("-acgt")
    res = helpers.run_lib(gpu_ctx, 0, True, M, -5, idx, 5, pk)
    helpers.compare(res, helpers.run_oracle(0, True, M, -5, idx, pk), "fasta -> align")
    seqobjs = f.sequences(alphabet.DNAgapped)
    assert seqobjs[3].ID == "s3" and seqobjs[3].Desc == "test" and bytes(seqobjs[3].Slice()) == seqs[3]


# ---- the reference's own fixtures ---------------------------------------------------------------

def _fixtures():
    import json
    import os

    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seqio_fixtures.json")))


def test_fasta_reader_fixtures_of_the_reference(built):
    """io/seqio/fasta/fasta_test.go:95-126 TestReadFasta: both layouts of the alignment give expectN / expectS."""
    fx = _fixtures()["fasta"]
    assert len(fx["headers"]) == 11
    for text in fx["inputs"]:
        for stride in (1, 2):
            f = seqio.parse_fasta(text.encode(), qletters=(stride == 2), lib_path=helpers.EMU_LIB)
            assert len(f) == 11
            for k in range(11):
                header = f.names[k] + ((" " + f.descs[k]) if f.descs[k] else "")
                assert header == fx["headers"][k]
                assert f.raw(k)[::stride] == fx["sequences"][k].encode()
                assert int(f.seq_len[k]) == len(fx["sequences"][k])
        assert ref_fasta(text.encode()) == [(n, d, bytes(f.raw(k)[::2])) for k, (n, d) in enumerate(zip(f.names, f.descs))]


def test_fastq_reader_fixtures_of_the_reference(built):
    """io/seqio/fastq/fastq_test.go:401-424 TestReadFastq over all nine fqTests (Sanger)."""
    fx = _fixtures()["fastq"]
    assert len(fx["tests"]) == 9
    for t in fx["tests"]:
        f = seqio.parse_fastq(t["text"].encode(), seqio.SANGER, lib_path=helpers.EMU_LIB)
        assert len(f) == len(fx["ids"]) == 5
        for k, want in enumerate(t["seqs"]):
            header = f.names[k] + ((" " + f.descs[k]) if f.descs[k] else "")
            assert header == fx["ids"][k]
            raw = f.raw(k)
            if want is None:                       # the reference's nil QLetters: an empty record
                assert int(f.seq_len[k]) == 0 and raw == b""
                continue
            assert raw[0::2] == want["letters"].encode()
            assert list(raw[1::2]) == want["quals"]
