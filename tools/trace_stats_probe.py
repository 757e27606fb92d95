"""Probe (variant build with -DBA_TRACE_STATS): where the team traceback of config 4 spends its cycles."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from biogo_b200 import _lib, synth  # noqa: E402

path = os.environ.get("BA_PROBE_LIB") or os.path.join(ROOT, "biogo_b200", "lib", "variants", "libbiogoalign_tstats.so")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
w = synth.config(cfg, n)
ctx = _lib.Context(0, path)
ctx.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
                w.gap_open if w.affine else 0, w.alpha.LetterIndex())
for mode in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1"])]:
    ctx.set_option("trace_mode", mode)
    out = (C.c_ulonglong * 8)()
    have = hasattr(ctx.lib, "ba_debug_trace_stats")
    for rep in range(3):
        if have:
            ctx.lib.ba_debug_trace_stats(out, 1)
        s = ctx.align_packed(w.letters, w.stride, w.ref_off, w.ref_len, w.qry_off, w.qry_len, collect=False)
        if have:
            ctx.lib.ba_debug_trace_stats(out, 1)
    r, steps, c_search, c_rebuild, c_walk = [int(out[k]) for k in range(5)]
    print(f"cfg{cfg} n={n} trace_mode {mode}: trace {s['ms_trace']:.3f} ms fill {s['ms_fill']:.3f} ms; thread-0 teams: rounds {r}, "
          f"steps {steps} ({steps / max(r, 1):.1f}/round), cycles search {c_search / 1e6:.2f} M rebuild {c_rebuild / 1e6:.2f} M "
          f"({c_rebuild / max(r, 1):.0f}/round) walk {c_walk / 1e6:.2f} M ({c_walk / max(steps, 1):.1f}/step)")
