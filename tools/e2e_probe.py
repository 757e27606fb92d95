"""Probe: where the end-to-end time of ba_align_batch goes, one call at a time vs two host threads."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biogo_b200 import _lib, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
w = synth.config(cfg, n)


def mk():
    c = _lib.Context(0)
    c.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
                  w.gap_open if w.affine else 0, w.alpha.LetterIndex())
    return c


a, b = mk(), mk()
nb = w.letters.nbytes
hp = a.lib.ba_alloc_host(nb)
h = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nb,))
h[:] = w.letters


def step(cx, log=None, tag=""):
    t0 = time.perf_counter()
    s = cx.align_packed(h, w.stride, w.ref_off, w.ref_len, w.qry_off, w.qry_len, collect=False)
    t1 = time.perf_counter()
    if log is not None:
        log.append((tag, round(t0 * 1e3 % 100000, 2), round((t1 - t0) * 1e3, 2), round(s["ms_h2d"], 2),
                    round(s["ms_kernels"], 2), round(s["ms_d2h"], 2), round(s["ms_host"], 2)))


for cx in (a, b):
    step(cx)
    step(cx)
log = []
t0 = time.perf_counter()
for _ in range(6):
    step(a, log, "serial")
print("serial ms/step", (time.perf_counter() - t0) / 6 * 1e3)
for l in log:
    print(l)
log = []


def worker(cx, tag):
    for _ in range(4):
        step(cx, log, tag)


t0 = time.perf_counter()
th = [threading.Thread(target=worker, args=(a, "A")), threading.Thread(target=worker, args=(b, "B"))]
[x.start() for x in th]
[x.join() for x in th]
print("2 threads ms/step", (time.perf_counter() - t0) / 8 * 1e3)
for l in sorted(log, key=lambda x: x[1]):
    print(l)
