#!/bin/bash
# bench config $1 with the default library and with every variant under biogo_b200/lib/variants (kernel experiments)
C=${1:-2}
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs --no-onebatch --steps 10 --warmup 3 --config $C"
run() { $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown_ms_per_step']
print('$1', 'cfg$C', round(d['value']), 'ms', round(d['ms_per_step'],3), 'fill', round(b['fill'],3), 'trace', round(b['trace'],3))"; }
run default
for f in biogo_b200/lib/variants/libbiogoalign_*.so; do BA_LIB_PATH=$PWD/$f run $(basename $f .so | sed 's/libbiogoalign_//'); done
