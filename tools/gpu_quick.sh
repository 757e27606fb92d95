#!/bin/bash
# quick GPU check of a kernel change: parity suite, then the device-resident bench of all configs (no CPU legs)
T=${1:-quick}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${T}_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-onebatch --no-e2e --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
b=d["breakdown_ms_per_step"]
print("cfg2", round(d["value"]), "ms", round(d["ms_per_step"],3), "fill", round(b["fill"],3), "trace", round(b["trace"],3))
for k,v in d.get("other_configs",{}).items():
    print(k, round(v["value"]), "ms", round(v["ms_per_step"],3), "fill", round(v["fill_ms_per_step"],3), "trace", round(v["trace_ms_per_step"],3))
P
