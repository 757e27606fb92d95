#!/bin/bash
# ncu --set full captures (source-level) of the traceback kernels of configs 2 and 3 and the config-3 fill kernel.
# Run on the GPU box: bash tools/prof_trace.sh <tag>
T=${1:-prof}
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs --no-onebatch --steps 1 --warmup 1"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ba_trace16_kernel -s 1 -c 1 -f -o gpurun_out/${T}_trace_cfg2 $B --config 2 > gpurun_out/${T}_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ba_trace16_kernel -s 1 -c 1 -f -o gpurun_out/${T}_trace_cfg3 $B --config 3 > gpurun_out/${T}_ncu2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ba_fill16_kernel -s 1 -c 1 -f -o gpurun_out/${T}_fill_cfg3 $B --config 3 > gpurun_out/${T}_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
