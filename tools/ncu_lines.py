#!/usr/bin/env python
"""Per-source-line totals from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`:
stall samples, warp instructions executed, the dominant stall reasons.  Usage: ncu_lines.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
hdr = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "" and r[0].isdigit():
        if len(r) > len(hdr):   # a source line with unescaped quotes and commas (inline asm): re-join its pieces
            extra = len(r) - len(hdr)
            r = r[:1] + [",".join(r[1:2 + extra])] + r[2 + extra:]
        lines.append((cur_file, int(r[0]), r[1], r))
ix = {n: i for i, n in enumerate(hdr)}
si = ix["# Samples"]; ii = ix["Instructions Executed"]; ti = ix["Thread Instructions Executed"]
stall_cols = [(n, i) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n]
def num(x):
    try: return float(x)
    except: return 0.0
tot_s = sum(num(l[3][si]) for l in lines); tot_i = sum(num(l[3][ii]) for l in lines)
print(f"total samples {tot_s:.0f}  warp instructions {tot_i:.0f}")
agg = {}
for f, ln, src, r in lines:
    k = (f, ln)
    a = agg.setdefault(k, [0.0, 0.0, 0.0, src, {}])
    a[0] += num(r[si]); a[1] += num(r[ii]); a[2] += num(r[ti])
    for n, i in stall_cols:
        a[4][n] = a[4].get(n, 0.0) + num(r[i])
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(a[4].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{n[6:]}={v:.0f}" for n, v in st if v > 0)
    print(f"{100*a[0]/tot_s:5.1f}% smp {100*a[1]/tot_i:5.1f}% ins thr/ins {a[2]/max(a[1],1):4.1f} {f}:{ln:<5d} {a[3].strip()[:90]}  [{sts}]")
