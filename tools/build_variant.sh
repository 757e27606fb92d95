#!/bin/bash
# Build a variant of libbiogoalign.so with extra -D flags (kernel experiments): tools/build_variant.sh NAME -DFOO=1 ...
# The result is biogo_b200/lib/variants/libbiogoalign_NAME.so; run bench.py with BA_LIB_PATH pointing at it.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
OUT=$ROOT/biogo_b200/lib/variants; OBJ=$OUT/obj_$NAME
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $*"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC $FLAGS -c "$ROOT/biogo_b200/csrc/ba_api.cu" -o "$OBJ/ba_api.o" &
for k in 0 1 2; do for l in 0 1 2; do
  $NVCC $FLAGS -DBA_INST_KIND=$k -DBA_INST_LIN=$l -c "$ROOT/biogo_b200/csrc/ba_fill16_inst.cu" -o "$OBJ/f_${k}_${l}.o" &
done; done
wait
$NVCC $FLAGS -shared -o "$OUT/libbiogoalign_$NAME.so" "$OBJ"/ba_api.o "$OBJ"/f_*.o
rm -rf "$OBJ"
echo "$OUT/libbiogoalign_$NAME.so"
