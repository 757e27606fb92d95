#!/bin/bash
# local helper: run a command on the GPU box through gpurun, retrying only while the pod answers "transient"
# usage: tools/gp.sh <timeout> '<command>'
T=$1; shift
for k in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gp.log 2>&1
  if grep -q "status=transient" /tmp/gp.log; then sleep 75; continue; fi
  break
done
tail -${GP_TAIL:-40} /tmp/gp.log
