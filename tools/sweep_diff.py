#!/usr/bin/env python3
"""Diagnostic: rows of ReplicateSweepStream with n_sub sub-batches against one ReplicateSweep (1250 x 100 cells x 10k SNVs)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from scsommerclock_b200 import engine, synth
n_cells, n_snv, R = 100, 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 1250
n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda", 0)
ctx = engine.Context(0)
names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
columns = names + ["healthycell"]
trees = [bench.make_tree(n_cells, seed=777 + r) for r in range(R)]
newicks = [ts.newick(names, "healthycell", scale=0.05) for ts in trees]
packs = []
for r, ts in enumerate(trees):
    lo = torch.as_tensor(ts.lo[:-1], device=dev); hi = torch.as_tensor(ts.hi[:-1], device=dev)
    rate = torch.as_tensor(ts.branch_len[:-1], device=dev, dtype=torch.float32)
    pk, pitch, _ = synth.observe_packed_torch(lo, hi, torch.as_tensor(ts.leaf_rank), n_snv, 0.1, 0.01, 0.15, dev, seed=5000 + r, rate=rate, extra_cols=1, chunk=n_snv)
    packs.append(pk)
packed = torch.cat(packs, dim=0).contiguous()
pinned = torch.empty(packed.shape, dtype=torch.uint8, pin_memory=True); pinned.copy_(packed); torch.cuda.synchronize()
one = engine.ReplicateSweep(ctx, [n_snv] * R, columns, bench.W_MAXS)
one.parse_trees(newicks); one.enqueue(packed, 0.01, 0.05); want = one.rows(); y_one = one.d_y.cpu().numpy().reshape(R, -1)
pipe = engine.ReplicateSweepStream(ctx, [n_snv] * R, columns, bench.W_MAXS, n_sub=n_sub)
print("one :", one.place.info()); print("sub0:", pipe.subs[0].place.info())
for it in range(1):
    rows, status = pipe.run(pinned, newicks, 0.01, 0.05)
    bad = [i for i in range(R) if rows[i] != want[i]]
    y = np.concatenate([s.d_y.cpu().numpy().reshape(s.G, -1) for s in pipe.subs])
    dy = np.abs(y - y_one).max(axis=1)
    print("run", it, "differing rows:", len(bad), bad[:20], "max |dY|", dy.max(), "groups with dY>0:", int((dy > 0).sum()), "cuts", pipe.cuts[:6])
    for i in bad[:3]:
        print("  want", want[i][:150]); print("  got ", rows[i][:150])
