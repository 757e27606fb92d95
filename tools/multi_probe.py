"""Timing breakdown of ba_align_batch_multi (builder tool): BA_TIMING=1 python tools/multi_probe.py [config] [pairs] [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401  (device count)
from biogo_b200 import _lib, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
w = synth.config(cfg, n)
nd = torch.cuda.device_count()
mc = _lib.MultiContext(tuple(range(nd)))
nb = w.letters.nbytes
hp = mc.lib.ba_alloc_host(nb * reps)
hl = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nb * reps,))
ro, qo = [], []
for r in range(reps):
    hl[r * nb:(r + 1) * nb] = w.letters
    ro.append(w.ref_off + r * nb)
    qo.append(w.qry_off + r * nb)
ro, qo = np.concatenate(ro), np.concatenate(qo)
rl, ql = np.tile(w.ref_len, reps), np.tile(w.qry_len, reps)
mc.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
               w.gap_open if w.affine else 0, w.alpha.LetterIndex())
for _ in range(3):
    mc.align_packed(hl, w.stride, ro, rl, qo, ql, collect=False)
print("---- timed call", flush=True)
t0 = time.perf_counter()
st = mc.align_packed(hl, w.stride, ro, rl, qo, ql, collect=False)
print("wall ms", (time.perf_counter() - t0) * 1e3, {k: round(st[k], 2) for k in ("ms_h2d", "ms_kernels", "ms_fill", "ms_trace", "ms_d2h", "ms_host")})
