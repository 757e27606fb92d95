#!/bin/bash
# Evidence run on one B200 (results under gpurun_out/$1_*): the default bench line, the launch list of config 2,
# ncu --set full captures of the fill and traceback kernels of configs 2 and 3.  gpurun brings back at most 64 MiB,
# so the reports are reduced on the box to their raw-metric summary and per-source-line CSV (gzip); only the
# config-2 fill report itself travels.
T=${1:-final}
if [ -z "$SKIP_BENCH" ]; then
timeout 400 python bench.py > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err; echo "bench rc=$?"
bash tools/launches.sh 2 $T > gpurun_out/${T}_launches_summary.txt 2>&1
fi
cap() {  # config, mangled-name regex, tag, keep report?
  BASE=mangled bash tools/prof_one.sh $1 $2 ${T}_$3
  python tools/ncu_summary.py gpurun_out/${T}_$3.ncu-rep > gpurun_out/${T}_$3_summary.txt
  ncu -i gpurun_out/${T}_$3.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/${T}_$3_lines.csv.gz
  [ "$4" = keep ] || rm -f gpurun_out/${T}_$3.ncu-rep
}
cap 2 ba_fill16_kernelILi0ELi8ELb0ELb0ELb0ELb0ELb1E fill16_cfg2 keep
cap 2 ba_trace16_kernelILi0ELb0E trace16_cfg2
cap 3 ba_fill16_kernelILi1ELi10E fill16_cfg3
cap 3 ba_trace16_kernelILi1ELb0E trace16_cfg3
du -sh gpurun_out
