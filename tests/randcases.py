"""Seeded random batches shared by the emulator (CPU) and GPU parity tests."""
import random

import golden_vectors as gv

DNA_LETTERS = b"ACGT"


def rand_matrix(rng, n=5):
    style = rng.randrange(3)
    if style == 0:
        return gv.match(rng.choice([-1, -2, 0, -3]), rng.choice([1, 2, 3]), rng.choice([-1, -2, 0]), n)
    if style == 1:
        return [[rng.randint(-4, 4) for _ in range(n)] for _ in range(n)]
    M = gv.match(rng.choice([-2, -4]), 5, -4, n)
    return M


def rand_pairs(rng, count, lens_r, lens_q, related=0.5, dirty=0.2, alpha=DNA_LETTERS, dirty_alpha=b"ACGT-acgtN"):
    pairs = []
    for _ in range(count):
        n, m = rng.choice(lens_r), rng.choice(lens_q)
        al = dirty_alpha if rng.random() < dirty else alpha
        r = bytes(rng.choice(al) for _ in range(n))
        if rng.random() < related and n > 0 and m > 0:
            q = bytearray(r[:m]) if m <= n else bytearray(r + bytes(rng.choice(alpha) for _ in range(m - n)))
            for x in range(len(q)):
                u = rng.random()
                if u < 0.10:
                    q[x] = rng.choice(alpha)
            if len(q) > 4 and rng.random() < 0.5:
                cut = rng.randrange(1, len(q) - 1)
                q = q[:cut] + q[cut + rng.randrange(1, 3):] + bytes(rng.choice(alpha) for _ in range(2))
                q = q[:m]
            q = bytes(q)
        else:
            q = bytes(rng.choice(al) for _ in range(m))
        pairs.append((r, q))
    return pairs


def sparse_dirty_pairs(rng, count, max_len=45):
    """Pairs with at most a few stray letters -- one illegal letter ('N'), a gap letter ('-') or none -- at random
    positions of reference and/or query, lengths 0..max_len: packed back to back they give every word alignment
    of a sequence start, of its end and of the stray letter (the validation kernel scans aligned words)."""
    pairs = []
    for _ in range(count):
        seqs = []
        for _s in range(2):
            n = rng.randrange(0, max_len + 1)
            s = bytearray(rng.choice(b"ACGT") for _ in range(n))
            u = rng.random()
            if n and u < 0.35:
                s[rng.randrange(n)] = ord("N")
            elif n and u < 0.5:
                s[rng.randrange(n)] = ord("-")
            if n and rng.random() < 0.1:
                s[rng.randrange(n)] = ord("N")
            seqs.append(bytes(s))
        pairs.append((seqs[0], seqs[1]))
    return pairs


def region_reuse_chunked_case(ctx, helpers, match_matrix, dna_idx, repeats):
    """Arena-limited waves (tiny arena: every wave re-uses a checkpoint region) of a batch whose letters arrive in
    several upload groups, so that the fill is chunked and alternates between the two fill streams -- the
    combination in which tools/fuzz_oracle.py found a stream race (the second fill stream did not wait for the
    traceback of the wave that last used the region).  Timing-dependent, hence repeated."""
    import random

    rng = random.Random(1474)
    pairs = rand_pairs(rng, 200, [1, 7, 33, 90], [1, 9, 40, 100], related=0.7, dirty=0.0)
    pk = helpers.pack(pairs)
    M = match_matrix(-3, 2, -1)
    orc = helpers.run_oracle(0, True, M, -8, dna_idx, pk)
    opts = {"arena_bytes": 200000, "upload_groups": 7, "chunk_pairs": 8, "fill_streams": 2, "waves": 1}
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        for it in range(repeats):
            res = helpers.run_lib(ctx, 0, True, M, -8, dna_idx, 5, pk)
            assert res.stats["waves"] > 2
            helpers.compare(res, orc, f"region reuse under a chunked fill, run {it}")
    finally:
        for k, v in (("arena_bytes", 0), ("upload_groups", 0), ("chunk_pairs", 8192), ("fill_streams", 2), ("waves", 0)):
            ctx.set_option(k, v)


def biased_sym_sw_case(ctx, helpers, match_matrix, dna_idx, go, gap):
    """SWAffine with match 10 / mismatch -9 on long identical stretches (scores near the top of the 16-bit range)
    plus random related pairs, against the oracle; see test_biased_sym_sw_kernel_at_its_bounds."""
    import random

    rng = random.Random(20261020 + go)
    M = match_matrix(gap, 10, -9)
    base = bytes(rng.choice(b"ACGT") for _ in range(2900))
    pairs = [(base, base), (base[:700], base[100:600]), (base[:300] + base[400:1500], base[:1400])]
    pairs += rand_pairs(rng, 24, [1, 9, 60, 130, 300], [1, 16, 64, 257, 300], related=0.9, dirty=0.0)
    pk = helpers.pack(pairs)
    res = helpers.run_lib(ctx, 0, True, M, go, dna_idx, 5, pk)
    helpers.compare(res, helpers.run_oracle(0, True, M, go, dna_idx, pk), f"biased SYM SW go {go} gap {gap}")
    if gap < 0:
        assert res.stats["pairs_fast16"] == len(pairs)
