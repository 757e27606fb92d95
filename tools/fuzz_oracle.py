"""GPU fuzz against the CPU oracle: random scorings, kinds, lengths and strides, both device paths.

    python tools/fuzz_oracle.py [seed] [seconds]

FUZZ_LIB=<path> runs another build of the library (e.g. the CPU emulation, plain or AddressSanitizer);
FUZZ_SHAPES=short,edge restricts the length classes (the emulator is slow on long pairs); FUZZ_LOG=1 prints every
round's parameters before it runs (the last line names the round that crashed).
"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import randcases  # noqa: E402
from biogo_b200 import _lib, matrix  # noqa: E402
from oracle import oracle as O  # noqa: E402

DNA_IDX = O.build_index("-acgt")
PROT_IDX = O.build_index("-abcdefghijklmnpqrstvwxyz*")
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
rng = random.Random(seed)
ctx = _lib.Context(0, os.environ.get("FUZZ_LIB") or None)
SHAPES = (os.environ.get("FUZZ_SHAPES") or "short,mid,long,mixed,edge,edge").split(",")
t_end = time.time() + budget
rounds = mism = fastpairs = 0
while time.time() < t_end:
    kind = rng.randrange(3)
    affine = rng.random() < 0.6
    protein = rng.random() < 0.3
    if protein:
        M = matrix.with_gap(matrix.BLOSUM62, rng.choice([-1, -2, -4]))
        idx, al, alpha = PROT_IDX, 26, b"ACDEFGHIKLMNPQRSTVWY"
    else:
        gap, ma, mi = rng.choice([-1, -2, -3]), rng.choice([1, 2, 5]), rng.choice([-1, -3, -4])
        gap2 = gap if rng.random() < 0.7 else rng.choice([-1, -2, -3])   # sometimes er != eq
        M = [[0] + [gap2] * 4] + [[gap] + [ma if a == b else mi for b in range(4)] for a in range(4)]
        idx, al, alpha = DNA_IDX, 5, b"ACGT"
    go = rng.choice([-2, -5, -8, -12]) if affine else 0
    shape = rng.choice(SHAPES)
    lens_r = {"short": [1, 7, 33, 90], "mid": [150, 300, 700], "long": [900, 2200, 4000], "mixed": [5, 120, 900, 3000],
              "edge": [1, 2, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 513]}[shape]
    lens_q = {"short": [1, 9, 40, 100], "mid": [100, 257, 520], "long": [1030, 1800, 3100], "mixed": [3, 250, 1100, 2600],
              "edge": [1, 2, 15, 16, 17, 31, 32, 33, 64, 65, 159, 160, 161, 255, 256, 257, 319, 320, 321, 511, 512, 513,
                       767, 768, 769, 1024, 1025]}[shape]
    count = {"short": 200, "mid": 60, "long": 8, "mixed": 24, "edge": 80}[shape]
    pairs = randcases.rand_pairs(rng, count, lens_r, lens_q, related=rng.choice([0.0, 0.7, 0.95]),
                                 dirty=rng.choice([0.0, 0.05]), alpha=alpha)
    stride = rng.choice([1, 2])
    pk = helpers.pack(pairs, stride)
    opts = {"trace_mode": rng.choice([0, 1, 2, 3, 4, 5]), "long_mode": rng.choice([0, 1, 2]),
            "plan_cache": rng.choice([0, 1]), "sym_kernels": rng.choice([0, 1])}
    # round-2 scheduler: waves with overlapped traceback, region reuse, grouped upload + chunked fill
    opts.update({"waves": rng.choice([0, 0, 1, 2, 3, 5]), "upload_groups": rng.choice([0, 1, 3, 7]),
                 "chunk_pairs": rng.choice([8, 64, 8192]), "fill_claims": rng.choice([0, 1, 2]),
                 "fill_streams": rng.choice([1, 2]), "trace_cta": rng.choice([0, 64, 128])})
    if rng.random() < 0.1:
        opts["slot_cap"] = rng.choice([1, 3])
    if rng.random() < 0.15:
        opts["arena_bytes"] = rng.choice([200_000, 5_000_000])
    if os.environ.get("FUZZ_ONLY") and rounds + 1 != int(os.environ["FUZZ_ONLY"]):   # replay one round of a seed
        rounds += 1
        if rounds > int(os.environ["FUZZ_ONLY"]):
            break
        continue
    for kv in (os.environ.get("FUZZ_SET") or "").split(","):   # e.g. FUZZ_SET=trace_mode=1: override drawn options
        if "=" in kv:
            opts[kv.split("=")[0]] = int(kv.split("=")[1])
    if os.environ.get("FUZZ_LOG"):
        print("ROUND", dict(round=rounds + 1, kind=kind, affine=affine, protein=protein, go=go, M0=M[1][:3], shape=shape,
                            stride=stride, npairs=len(pairs), opts=opts), flush=True)
    for k, v in opts.items():
        ctx.set_option(k, v)
    orc = helpers.run_oracle(kind, affine, M, go, idx, pk, stride, 16)
    ok = True
    try:
        for force in (0, 1) * int(os.environ.get("FUZZ_REPEAT") or 1):   # (races: the same round many times)
            ctx.set_option("force_path", force)
            res = helpers.run_lib(ctx, kind, affine, M, go, idx, al, pk, stride)
            if force == 0:
                fast = res
            try:
                helpers.compare(res, orc, f"force_path {force}")
            except AssertionError as ex:
                ok = False
                print("  ", str(ex)[:400], flush=True)
    finally:
        ctx.set_option("force_path", 0)
        ctx.set_option("arena_bytes", 0)
        ctx.set_option("slot_cap", 0)
        ctx.set_option("trace_mode", 1)
        ctx.set_option("long_mode", 1)
        ctx.set_option("plan_cache", 1)
        ctx.set_option("sym_kernels", 1)
        for k, v in (("waves", 0), ("upload_groups", 0), ("chunk_pairs", 8192), ("fill_claims", 1), ("fill_streams", 2),
                     ("trace_cta", 0)):
            ctx.set_option(k, v)
    rounds += 1
    fastpairs += fast.stats["pairs_fast16"]
    if not ok:
        mism += 1
        if mism > 6:
            break
        print("MISMATCH", dict(seed=seed, round=rounds, kind=kind, affine=affine, protein=protein, go=go, M0=M[1][:3],
                               shape=shape, stride=stride, opts=opts), flush=True)
print(f"fuzz seed {seed}: {rounds} rounds, {fastpairs} pairs on the packed path, {mism} mismatches")
sys.exit(1 if mism else 0)
