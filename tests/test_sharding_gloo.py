"""world_size-2 gloo test of the N>1 path's host logic: rows are sharded, each
rank's partial soft weights (from the CPU oracle here) are all-reduced, and the
result equals the single-rank answer; replicate sharding covers every replicate
exactly once."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    from oracle import pt_oracle as O
    from scsommerclock_b200 import sharding, synth
    r, w, _ = sharding.init_distributed(backend="gloo")
    rng = np.random.default_rng(7)                  # same data on every rank
    tree = synth.coalescent_tree(24, rng)
    S = tree.clade_matrix().astype(float)
    codes = synth.observe(tree, synth.throw_mutations(tree, 501, rng), 0.1, 0.01, 0.1, rng)
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    lo, hi = sharding.shard_range(muts.shape[0], r, w)
    part = torch.from_numpy(O.map_mutations_gt(muts[lo:hi], S, 0.01, 0.1))
    sharding.allreduce_soft_weights(part)
    full = O.map_mutations_gt(muts, S, 0.01, 0.1)
    t = sharding.max_over_ranks(float(r + 1))
    sharding.barrier()
    q.put((r, float(np.abs(part.numpy() - full).max()), lo, hi, t, sharding.replicate_shard(11, r, w)))


def test_row_sharding_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][2] == 0 and out[0][3] == out[1][2] and out[1][3] == 501      # the shards tile the rows
    assert all(o[1] < 1e-9 for o in out)                                        # all-reduced partials == full answer
    assert all(o[4] == 2.0 for o in out)                                        # max over ranks
    assert sorted(out[0][5] + out[1][5]) == list(range(11))


def test_shard_range_edges():
    from scsommerclock_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000001):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
