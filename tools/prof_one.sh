#!/bin/bash
# one ncu --set full capture (source-level): [BASE=mangled] tools/prof_one.sh <config> <kernel regex> <tag>
C=${1:-2}; K=${2:-ba_trace16_kernel}; T=${3:-prof}
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs --no-onebatch --steps 1 --warmup 1 --config $C"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base ${BASE:-function} -k regex:$K -s 1 -c 1 -f -o gpurun_out/${T} $B > gpurun_out/${T}.log 2>&1
ls -la gpurun_out/${T}.ncu-rep
