#!/bin/bash
# launch list (per-kernel device time) of one bench step of config $1 -> gpurun_out/$2_launches_cfg$1.csv
C=${1:-2}; T=${2:-ll}
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs --no-onebatch --steps 1 --warmup 1 --config $C"
$B > gpurun_out/${T}_plain_cfg$C.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_launches_cfg$C.csv $B > gpurun_out/${T}_ncu_cfg$C.log 2>&1
python - <<P
import csv,collections
rows=list(csv.reader(l for l in open("gpurun_out/${T}_launches_cfg$C.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e6 if u in("ns","nsecond") else v/1e3 if u in ("us","usecond") else v
    a=agg.setdefault(r[ki][:110],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print(f"{t:9.3f} ms {n:5d}x  {k}")
P
