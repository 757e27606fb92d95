"""Sweep of the overlap-scheduling options of the 16-bit path on one B200 (builder tool).

    python tools/overlap_sweep.py [config] [pairs] [out.jsonl]

For every option set: device-resident ms/step (host wall clock around blocking calls, after
warm-up), the spans the library reports, and -- for the sets marked e2e -- host-buffer calls,
one at a time and from two threads with one context each.
"""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biogo_b200 import _lib, synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
out_path = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/sweep.jsonl"
steps = int(os.environ.get("SWEEP_STEPS", "8"))
w = synth.config(cfg, n)
dev = torch.device("cuda", 0)

DEFAULTS = dict(waves=0, fill_streams=2, fill_claims=1, fill_cap=0, trace_cta=0, trace_mode=1, upload_groups=0,
                plan_cache=1)

SETS = [
    ("one-wave", dict(waves=1), True),
    ("auto", dict(), True),
    ("w2", dict(waves=2), False),
    ("w4", dict(waves=4), True),
    ("w6", dict(waves=6), False),
    ("w8", dict(waves=8), True),
    ("w12", dict(waves=12), False),
    ("w16", dict(waves=16), False),
    ("w8-claims2", dict(waves=8, fill_claims=2), False),
    ("w8-claims4", dict(waves=8, fill_claims=4), False),
    ("w8-cta128", dict(waves=8, trace_cta=128), False),
    ("w4-cta128", dict(waves=4, trace_cta=128), False),
    ("w8-one-fill-stream", dict(waves=8, fill_streams=1), False),
    ("w8-persistent", dict(waves=8, fill_claims=0), False),
    ("w8-persistent-cap5", dict(waves=8, fill_claims=0, fill_cap=5, trace_cta=128), False),
    ("w8-persistent-cap4", dict(waves=8, fill_claims=0, fill_cap=4, trace_cta=128), False),
    ("w4-persistent-cap5", dict(waves=4, fill_claims=0, fill_cap=5, trace_cta=128), False),
    ("w8-persistent-cap5-cta64", dict(waves=8, fill_claims=0, fill_cap=5, trace_cta=64), False),
    ("w8-warp-trace", dict(waves=8, trace_mode=2), False),
    ("w8-no-plan-cache", dict(waves=8, plan_cache=0), True),
    ("one-wave-no-plan-cache", dict(waves=1, plan_cache=0), False),
]
if os.environ.get("SWEEP_SETS"):   # custom sets: JSON {"name": {option: value, ...}, ...}
    SETS = [(k, {a: b for a, b in v.items() if a != "_e2e"}, bool(v.get("_e2e"))) for k, v in
            json.loads(os.environ["SWEEP_SETS"]).items()]   # "_e2e": 1 adds the host-buffer measurements
    DEFAULTS = {}
if os.environ.get("SWEEP_ONLY"):
    keep = set(os.environ["SWEEP_ONLY"].split(","))
    SETS = [s for s in SETS if s[0] in keep]


def mk():
    c = _lib.Context(0)
    c.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
                  w.gap_open if w.affine else 0, w.alpha.LetterIndex())
    return c


a, b = mk(), mk()
d_letters = torch.from_numpy(w.letters).to(dev)
d_ro = torch.from_numpy(w.ref_off).to(dev)
d_rl = torch.from_numpy(w.ref_len).to(dev)
d_qo = torch.from_numpy(w.qry_off).to(dev)
d_ql = torch.from_numpy(w.qry_len).to(dev)
nb = w.letters.nbytes
hp = a.lib.ba_alloc_host(nb)
h = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nb,))
h[:] = w.letters


def dev_step(cx):
    return cx.align_device(d_letters.data_ptr(), d_letters.numel(), w.stride, d_ro.data_ptr(), d_rl.data_ptr(),
                           d_qo.data_ptr(), d_ql.data_ptr(), w.ref_len, w.qry_len, stream=0, collect=False)


def host_step(cx):
    return cx.align_packed(h, w.stride, w.ref_off, w.ref_len, w.qry_off, w.qry_len, collect=False)


def timed(fn, k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = None
    for _ in range(k):
        last = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3, last


def two_threads(step, k):
    def worker(cx, kk):
        for _ in range(kk):
            step(cx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(a, (k + 1) // 2)), threading.Thread(target=worker, args=(b, k // 2))]
    [x.start() for x in th]
    [x.join() for x in th]
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3


ref_segs = None
with open(out_path, "a") as out:
    for name, opts, e2e in SETS:
        full = dict(DEFAULTS)
        full.update(opts)
        for cx in (a, b):
            for k, v in full.items():
                cx.set_option(k, v)
        try:
            for cx in (a, b):
                dev_step(cx)
                dev_step(cx)
            ms, s = timed(lambda: dev_step(a), steps)
            rec = {"cfg": cfg, "pairs": n, "set": name, "opts": opts, "dev_ms": round(ms, 3),
                   "dev_gcups": round(w.cells / ms / 1e6, 1), "waves": s["waves"], "fill_span": round(s["ms_fill"], 3),
                   "trace_span": round(s["ms_trace"], 3), "kernels_span": round(s["ms_kernels"], 3),
                   "host_ms": round(s["ms_host"], 3), "n_segs": s["n_segs"]}
            if ref_segs is None:
                ref_segs = s["n_segs"]
            rec["segs_ok"] = (s["n_segs"] == ref_segs)
            rec["dev2_ms"] = round(two_threads(dev_step, steps), 3)
            if e2e:
                for cx in (a, b):
                    host_step(cx)
                ms1, s1 = timed(lambda: host_step(a), steps)
                rec["e2e1_ms"] = round(ms1, 3)
                rec["e2e1_h2d_span"] = round(s1["ms_h2d"], 3)
                rec["e2e2_ms"] = round(two_threads(host_step, steps), 3)
        except Exception as ex:  # noqa: BLE001
            rec = {"cfg": cfg, "set": name, "error": str(ex)[:300]}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
        out.flush()
