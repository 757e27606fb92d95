"""Multi-GPU host logic on the CPU: world_size-2 gloo run of the bench's sharding scheme (every
rank draws its own shard, no data-path collective, max-over-ranks timing) with the emulated
kernels standing in for the device.  Checks that shards are disjoint, deterministic, and that
each rank's results equal the oracle's."""
import os
import subprocess
import sys
import textwrap

import helpers

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch, torch.distributed as dist
    import helpers
    from biogo_b200 import _lib, synth
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    w = synth.config(2, 6, shard=rank)
    w = synth.Workload(w.name, w.kind, w.affine, w.matrix, w.gap_open, w.alpha, w.letters, w.stride,
                       w.ref_off, np.minimum(w.ref_len, 90).astype(np.int32), w.qry_off,
                       np.minimum(w.qry_len, 60).astype(np.int32))
    ctx = _lib.Context(0, helpers.EMU_LIB)
    pk = (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)
    res = helpers.run_lib(ctx, w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), 5, pk)
    helpers.compare(res, helpers.run_oracle(w.kind, True, w.matrix, w.gap_open, w.alpha.LetterIndex(), pk), "rank %d" % rank)
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
    digest = torch.tensor([float(int(w.letters[:64].astype(np.int64).sum()))], dtype=torch.float64)
    both = [torch.zeros_like(digest) for _ in range(2)]
    dist.all_gather(both, digest)
    if rank == 0:
        print(json.dumps({{"max": t.item(), "digests": [b.item() for b in both], "cells": w.cells}}))
    dist.destroy_process_group()
''')


def test_two_rank_sharding(built, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=helpers.ROOT, port=29517))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    import json

    line = json.loads(outs[0][0].strip().splitlines()[-1])
    assert line["max"] == 2.0
    assert line["digests"][0] != line["digests"][1]   # ranks drew different shards
