"""A small deterministic slice of tools/fuzz_oracle.py on the emulated kernels: random scorings, kinds,
boundary lengths, strides and kernel options, both device paths against the oracle."""
import random

import numpy as np
import pytest

import helpers
import randcases
from biogo_b200 import matrix
from oracle import oracle as O

DNA_IDX = O.build_index("-acgt")
PROT_IDX = O.build_index("-abcdefghijklmnpqrstvwxyz*")
EDGE_R = [1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 64, 65, 130]
EDGE_Q = [1, 2, 15, 16, 17, 32, 33, 64, 65, 159, 160, 161, 256, 257, 320, 321, 513]


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_fuzz_slice(emu_ctx, seed):
    rng = random.Random(seed)
    for _ in range(7):
        kind, affine, protein = rng.randrange(3), rng.random() < 0.6, rng.random() < 0.3
        if protein:
            M, idx, al, alpha = matrix.with_gap(matrix.BLOSUM62, rng.choice([-1, -2, -4])), PROT_IDX, 26, b"ACDEFGHIKLMNPQRSTVWY"
        else:
            gap, ma, mi = rng.choice([-1, -2, -3]), rng.choice([1, 2, 5]), rng.choice([-1, -3, -4])
            M = [[0] + [gap] * 4] + [[gap] + [ma if a == b else mi for b in range(4)] for a in range(4)]
            idx, al, alpha = DNA_IDX, 5, b"ACGT"
        go = rng.choice([-2, -5, -12]) if affine else 0
        pairs = randcases.rand_pairs(rng, 14, EDGE_R, EDGE_Q, related=rng.choice([0.0, 0.8]), dirty=rng.choice([0.0, 0.1]),
                                     alpha=alpha)
        stride = rng.choice([1, 2])
        pk = helpers.pack(pairs, stride)
        orc = helpers.run_oracle(kind, affine, M, go, idx, pk, stride)
        opts = {"trace_mode": rng.choice([0, 1, 2, 3, 4, 5]), "long_mode": rng.choice([0, 1, 2]),
                "plan_cache": rng.choice([0, 1]), "slot_cap": rng.choice([0, 0, 0, 2]), "sym_kernels": rng.choice([0, 1])}
        try:
            for k, v in opts.items():
                emu_ctx.set_option(k, v)
            for force in (0, 1):
                emu_ctx.set_option("force_path", force)
                res = helpers.run_lib(emu_ctx, kind, affine, M, go, idx, al, pk, stride)
                helpers.compare(res, orc, f"seed {seed} kind {kind} affine {affine} protein {protein} opts {opts} force {force}")
        finally:
            for k, v in (("force_path", 0), ("trace_mode", 1), ("long_mode", 1), ("plan_cache", 1), ("slot_cap", 0), ("sym_kernels", 1)):
                emu_ctx.set_option(k, v)
