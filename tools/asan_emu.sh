#!/bin/bash
# Memory check of the kernels without a GPU: build the CPU SIMT emulation of the CUDA sources with
# AddressSanitizer (device allocations are plain malloc there, so an out-of-bounds checkpoint,
# boundary-row or slot access shows up as a heap-buffer-overflow) and run the emulator test tier.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/ba_asan
mkdir -p "$OUT"
FLAGS="-O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -DBA_EMULATE -I$ROOT/tests/emu -I$ROOT/biogo_b200/csrc -fPIC -x c++"
for k in 0 1 2; do for l in 0 1 2; do
  g++ $FLAGS -DBA_INST_KIND=$k -DBA_INST_LIN=$l -c "$ROOT/biogo_b200/csrc/ba_fill16_inst.cu" -o "$OUT/f_${k}_${l}.o" &
done; done
g++ $FLAGS -c "$ROOT/biogo_b200/csrc/ba_api.cu" -o "$OUT/api.o" &
wait
g++ -shared -fsanitize=address -o "$OUT/libemu_asan.so" "$OUT"/api.o "$OUT"/f_*.o
EMU="$ROOT/tests/emu/_build/libbiogoalign_emu.so"
cp "$EMU" "$OUT/emu_backup.so"
trap 'cp "$OUT/emu_backup.so" "$EMU"; touch "$EMU"' EXIT
cp "$OUT/libemu_asan.so" "$EMU"; touch "$EMU"
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
  python -m pytest "$ROOT/tests/test_emu_parity.py" "$ROOT/tests/test_format.py" "$ROOT/tests/test_multi_device.py" -x -q -m "not gpu" -p no:cacheprovider
