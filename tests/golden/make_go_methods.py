"""Regenerates tests/golden/go_align_methods.txt from the reference tree (run in the build container,
where /root/reference exists): the receiver/method/signature of every unexported method the overlay
replaces -- `grep -h "^func (a " align/*_letters.go align/*_qletters.go`."""
import glob
import os
import re

ref = os.environ.get("BIOGO_REF", "/root/reference")
out = []
for f in sorted(glob.glob(os.path.join(ref, "align", "*_letters.go")) + glob.glob(os.path.join(ref, "align", "*_qletters.go"))):
    for ln in open(f):
        m = re.match(r"^func \(a (\w+)\) (\w+)\((.*)\) \((.*)\) \{", ln)
        if m:
            out.append(f"{os.path.relpath(f, ref)}\t{m.group(1)}\t{m.group(2)}\t{m.group(3)}\t{m.group(4)}")
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "go_align_methods.txt"), "w").write("\n".join(out) + "\n")
print(len(out), "methods")
