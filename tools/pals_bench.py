"""align/pals/dp batch on the GPU against the C oracle on the host cores (builder tool).

    python tools/pals_bench.py [problems] [out.json]
"""
import json
import os
import random
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_pals as tp  # noqa: E402
from biogo_b200 import _lib, alphabet, pals  # noqa: E402
from oracle import oracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/pals_bench.json"
rng = random.Random(99)
base = [tp._random_problem(rng, rng.randint(4000, 10000), rng.randint(3000, 8000), rng.randint(4, 12)) for _ in range(64)]
problems = [base[k % len(base)] for k in range(n)]
ctx = _lib.Context(0)
aligners, tzs = [], []
for tgt, qry, traps in problems:
    a = pals.NewAligner(tgt, qry, 6, 50, 0.8, ctx, alphabet.DNA)
    a.Costs = pals.Costs(*tp.COSTS)
    aligners.append(a)
    tzs.append(traps)
pals.AlignTrapsBatch(aligners[:64], tzs[:64], ctx)
t0 = time.perf_counter()
got = pals.AlignTrapsBatch(aligners, tzs, ctx)
gpu_s = time.perf_counter() - t0
# the same batch through the ctypes call alone (packed once): library time without the Python packing
chunks, to, tl, qo, ql, troff, trs, off = [], [], [], [], [], [0], [], 0
for tgt, qry, traps in problems:
    to.append(off); tl.append(len(tgt)); off += len(tgt)
    qo.append(off); ql.append(len(qry)); off += len(qry)
    chunks += [tgt, qry]
    for z in traps:
        trs += list(z)
    troff.append(len(trs) // 4)
letters = np.frombuffer(b"".join(chunks), dtype=np.uint8)
ctx.pals_dp_batch(letters, to, tl, qo, ql, tp.DNA_IDX, troff, trs, 6, 50, 0.8, tp.COSTS)
t0 = time.perf_counter()
_, kernel_ms = ctx.pals_dp_batch(letters, to, tl, qo, ql, tp.DNA_IDX, troff, trs, 6, 50, 0.8, tp.COSTS)
lib_s = time.perf_counter() - t0
threads = os.cpu_count() or 1
sample = problems[:max(64, 8 * threads)]


def one(p):
    return O.pals_align_traps(p[0], p[1], tp.DNA_IDX, p[2], 6, 50, 0.8, tp.COSTS)


t0 = time.perf_counter()
with ThreadPoolExecutor(threads) as ex:
    want = list(ex.map(one, sample))
cpu_s = time.perf_counter() - t0
ok = all([(h.Abpos, h.Bbpos, h.Aepos, h.Bepos, h.LowDiagonal, h.HighDiagonal, h.Score, h.Error) for h in g] == w
         for g, w in zip(got, want))
rec = {"problems": n, "hits": sum(len(g) for g in got), "gpu_s_end_to_end": gpu_s, "gpu_problems_per_s": n / gpu_s,
       "library_call_s": lib_s, "kernel_ms": kernel_ms, "kernel_problems_per_s": n / (kernel_ms * 1e-3),
       "cpu_threads": threads, "cpu_sample": len(sample), "cpu_problems_per_s": len(sample) / cpu_s,
       "speedup": (n / gpu_s) / (len(sample) / cpu_s), "sample_equal_to_oracle": ok,
       "note": "GPU: host buffers in, de-duplicated hits out (python packing included); CPU: oracle/pals_oracle.c, one "
               "problem per thread (ctypes releases the GIL)"}
print(json.dumps(rec))
open(out, "w").write(json.dumps(rec) + "\n")
