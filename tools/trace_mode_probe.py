"""Probe: traceback time of the static and the persistent thread-per-pair kernels per config."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biogo_b200 import _lib, synth  # noqa: E402

for cfg, n in ((2, 125000), (1, 100000), (3, 125000)):
    w = synth.config(cfg, n)
    ctx = _lib.Context(0)
    ctx.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
                    w.gap_open if w.affine else 0, w.alpha.LetterIndex())
    for mode in (1, 5):
        ctx.set_option("trace_mode", mode)
        ts = []
        for _ in range(6):
            s = ctx.align_packed(w.letters, w.stride, w.ref_off, w.ref_len, w.qry_off, w.qry_len, collect=False)
            ts.append(s["ms_trace"])
        print(f"cfg{cfg} trace_mode {mode}: trace {np.median(ts[2:]):.3f} ms")
    del ctx
