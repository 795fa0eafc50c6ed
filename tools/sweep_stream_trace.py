#!/usr/bin/env python3
"""Host-side trace of engine.ReplicateSweepStream.run (1250 pairs of 100 cells x 10k SNVs).  usage: ... [n_sub]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from scsommerclock_b200 import engine

n_cells, n_snv, R = 100, 10000, 1250
n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = engine.Context(0)
names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
columns = names + ["healthycell"]
trees = [bench.make_tree(n_cells, seed=777 + r) for r in range(R)]
newicks = [ts.newick(names, "healthycell", scale=0.05) for ts in trees]
from scsommerclock_b200 import synth
dev = torch.device("cuda", 0)
packs = []
for r, ts in enumerate(trees):
    lo = torch.as_tensor(ts.lo[:-1], device=dev); hi = torch.as_tensor(ts.hi[:-1], device=dev)
    rate = torch.as_tensor(ts.branch_len[:-1], device=dev, dtype=torch.float32)
    pk, pitch, _ = synth.observe_packed_torch(lo, hi, torch.as_tensor(ts.leaf_rank), n_snv, 0.1, 0.01, 0.15, dev, seed=5000 + r, rate=rate, extra_cols=1, chunk=n_snv)
    packs.append(pk)
packed = torch.cat(packs, dim=0).contiguous()
pinned = torch.empty(packed.shape, dtype=torch.uint8, pin_memory=True); pinned.copy_(packed); torch.cuda.synchronize()
del packs, packed
pipe = engine.ReplicateSweepStream(ctx, [n_snv] * R, columns, bench.W_MAXS, n_sub=n_sub)
for it in range(4):
    torch.cuda.synchronize()
    pipe.trace = [] if it == 3 else None
    for sub in pipe.subs:
        sub.marks = [] if it == 3 else None
    t0 = time.perf_counter()
    rows, status = pipe.run(pinned, newicks, 0.01, 0.05)
    torch.cuda.synchronize()
    dt = 1e3 * (time.perf_counter() - t0)
print("n_sub %d: %.2f ms" % (n_sub, dt))
for k, a, b in pipe.trace_events:
    print("  device: sub %d kernels %.2f -> %.2f ms" % (k, pipe.trace_t0.elapsed_time(a), pipe.trace_t0.elapsed_time(b)))
for k, sub in enumerate(pipe.subs):
    print("  device stages sub %d:" % k, ", ".join("%s %.2f" % (w, pipe.trace_t0.elapsed_time(e)) for w, e in sub.marks))
print("  per-sub placement kernel times (pre, -, kernel, reduce):", [[round(v, 3) for v in sub.place.last_times()] for sub in pipe.subs])
for what, k, t in pipe.trace:
    print("  %6.2f ms  %s %d" % (t, what, k))
t0 = time.perf_counter(); pipe.subs[0].parse_trees(newicks[:pipe.cuts[1]]); print("parse of one sub-batch alone: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter(); r = pipe.subs[0].rows(); print("rows of one sub-batch alone: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
