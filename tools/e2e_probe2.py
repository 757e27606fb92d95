"""Where a single blocking ba_align_batch call spends its time (builder tool): spans reported by the library
for host-buffer calls under several upload/chunk settings."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biogo_b200 import _lib, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
w = synth.config(cfg, n)
c = _lib.Context(0)
c.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
              w.gap_open if w.affine else 0, w.alpha.LetterIndex())
nb = w.letters.nbytes
hp = c.lib.ba_alloc_host(nb)
h = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nb,))
h[:] = w.letters


def run(label, **opts):
    for k, v in opts.items():
        c.set_option(k, v)
    for _ in range(3):
        c.align_packed(h, w.stride, w.ref_off, w.ref_len, w.qry_off, w.qry_len, collect=False)
    t0 = time.perf_counter()
    k = 8
    for _ in range(k):
        s = c.align_packed(h, w.stride, w.ref_off, w.ref_len, w.qry_off, w.qry_len, collect=False)
    dt = (time.perf_counter() - t0) / k * 1e3
    print(label, "wall %.2f ms" % dt, {x: round(s[x], 2) for x in ("ms_h2d", "ms_kernels", "ms_fill", "ms_trace", "ms_d2h", "ms_host")},
          "fill_launches", s["fill_launches"], flush=True)


run("default")
run("one piece", upload_groups=1)
run("4 groups", upload_groups=4)
run("8 groups", upload_groups=8)
run("16 groups, chunk 4096", upload_groups=16, chunk_pairs=4096)
run("16 groups, one fill stream", upload_groups=16, chunk_pairs=8192, fill_streams=1)
run("32 groups", upload_groups=32, fill_streams=2)
