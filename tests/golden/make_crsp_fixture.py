"""Copies the reference's 1812 x 1797 nt benchmark input (align/crsp_test.go:7-70, the FASTA text `crspFa`
with two mixed-case records) to tests/golden/crsp.fa, and records the four benchmark scorings of
align/align_test.go:53-140 (input only -- the reference holds no expected output for them)."""
import os
import re

REF = os.path.join(os.environ.get("BIOGO_REF", "/root/reference"), "align")
src = open(os.path.join(REF, "crsp_test.go")).read()
text = re.search(r"var crspFa = `(.*?)`", src, flags=re.S).group(1)
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "crsp.fa"), "w").write(text)
print(len(text), "bytes")
