"""Extracts the reference's own reader fixtures into tests/golden/seqio_fixtures.json (run in the build
container, where /root/reference exists):
  io/seqio/fasta/fasta_test.go:30-93 (expectN, expectS) with the inputs testaln0_test.go / testaln1_test.go
  io/seqio/fastq/fastq_test.go:43-399 (expectedIds, expectedQLetters, fqTests)
so that ba_fasta_parse / ba_fastq_parse are pinned on the vectors the reference's tests hold."""
import json
import os
import re

REF = os.path.join(os.environ.get("BIOGO_REF", "/root/reference"), "io", "seqio")
HERE = os.path.dirname(os.path.abspath(__file__))


def raw(path, var):
    src = open(path).read()
    return re.search(r"var %s = `(.*?)`" % var, src, flags=re.S).group(1)


fa_test = open(os.path.join(REF, "fasta", "fasta_test.go")).read()
blockN = re.search(r"expectN = \[\]string\{(.*?)\n\t\}", fa_test, flags=re.S).group(1)
expectN = re.findall(r'"((?:[^"\\]|\\.)*)"', blockN)
blockS = re.search(r"expectS = \[\]\[\]alphabet\.Letter\{(.*?)\n\t\}", fa_test, flags=re.S).group(1)
expectS = re.findall(r'\[\]alphabet\.Letter\("([^"]*)"\)', blockS)
assert len(expectN) == 11 and len(expectS) == 11

fq_test = open(os.path.join(REF, "fastq", "fastq_test.go")).read()
ids = re.findall(r'"([^"]+)"', re.search(r"expectedIds = \[\]string\{(.*?)\n\t\}", fq_test, flags=re.S).group(1))
blk = re.search(r"expectedQLetters = constructQL\((.*?)\n\t\)", fq_test, flags=re.S).group(1)
letters = re.findall(r'\[\]alphabet\.Letter\("([^"]*)"\)', blk)
quals = [[int(x) for x in re.findall(r"\d+", q)] for q in re.findall(r"\n\t\t\t\{([\d, ]+)\},", blk)]
assert len(ids) == 5 and len(letters) == 5 and len(quals) == 5 and all(len(l) == len(q) for l, q in zip(letters, quals))
def ql(var):
    b = re.search(r"%s = constructQL\((.*?)\n\t\)" % var, fq_test, flags=re.S).group(1)
    ls = re.findall(r'\[\]alphabet\.Letter\("([^"]*)"\)', b)
    qs = [[int(x) for x in re.findall(r"\d+", q)] for q in re.findall(r"\n\t\t\t\{([\d, ]+)\},", b)]
    return ls, qs


records = {"expectedQLetters": (letters, quals), "plusStart": ql("plusStart"), "atStart": ql("atStart")}
tests = []
for m in re.finditer(r"fq: `(.*?)`,\n\t\t\tverbatim: (true|false),\n\t\t\tids:\s+expectedIds,\n\t\t\tseqs: \[\]alphabet\.QLetters\{(.*?)\n\t\t\t\},",
                     fq_test, flags=re.S):
    seqs = []
    for item in re.findall(r"(\w+)\[(\d+)\]|(nil)", m.group(3)):
        if item[2]:
            seqs.append(None)
        else:
            ls, qs = records[item[0]]
            seqs.append({"letters": ls[int(item[1])], "quals": qs[int(item[1])]})
    assert len(seqs) == 5
    tests.append({"text": m.group(1), "verbatim": m.group(2) == "true", "seqs": seqs})
assert len(tests) == 9, len(tests)
fixtures = {
    "source": "biogo v1.0.1 io/seqio/fasta/fasta_test.go:30-93, testaln0_test.go, testaln1_test.go; io/seqio/fastq/fastq_test.go:43-399",
    "fasta": {
        "alphabet": "Protein",
        "inputs": [raw(os.path.join(REF, "fasta", "testaln0_test.go"), "testaln0"),
                   raw(os.path.join(REF, "fasta", "testaln1_test.go"), "testaln1")],
        "headers": expectN,          # name + " " + description when a description is present
        "sequences": expectS,
    },
    "fastq": {
        "encoding": "Sanger (offset 33)",
        "ids": ids,
        # fqTests: the text, and per record the expected letters + Phred values (null = the reference's nil slice)
        "tests": tests,
    },
}
json.dump(fixtures, open(os.path.join(HERE, "seqio_fixtures.json"), "w"), indent=1)
print("fasta", len(expectN), "records; fastq", len(tests), "texts")
