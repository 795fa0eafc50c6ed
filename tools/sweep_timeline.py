#!/usr/bin/env python3
"""Where does the end-to-end replicate sweep spend its time?  Host timestamps per call and CUDA events per sub-batch."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from scsommerclock_b200 import engine, synth

R, n_cells, n_snv, n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 1250, 100, 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
ctx = engine.Context(0)
names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
columns = names + ["healthycell"]
trees = [bench.make_tree(n_cells, seed=777 + r) for r in range(R)]
newicks = [ts.newick(names, "healthycell", scale=0.05) for ts in trees]
pitch = ((n_cells + 1 + 3) // 4 + 15) // 16 * 16
pinned = torch.randint(0, 256, (R * n_snv, pitch), dtype=torch.uint8).pin_memory()
pinned[:, 26:] = 0
cuts = [R * k // n_sub for k in range(n_sub + 1)]
subs = [engine.ReplicateSweep(ctx, [n_snv] * (cuts[k + 1] - cuts[k]), columns, bench.W_MAXS) for k in range(n_sub)]
streams = [torch.cuda.Stream(device=dev) for _ in range(n_sub)]
for it in range(4):
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_sub)]
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    T = [time.perf_counter()]
    for k, sub in enumerate(subs):
        sub.parse_trees(newicks[cuts[k]:cuts[k + 1]]); T.append(time.perf_counter())
        ev[k][0].record(streams[k])
        sub.enqueue(pinned[cuts[k] * n_snv:cuts[k + 1] * n_snv], 0.01, 0.05, stream=streams[k]); T.append(time.perf_counter())
        ev[k][1].record(streams[k])
    rows = []
    for sub in subs:
        sub.results(); T.append(time.perf_counter())
        rows.extend(sub.rows()); T.append(time.perf_counter())
    torch.cuda.synchronize()
    if it >= 2:
        t = (np.array(T) - T[0]) * 1e3
        print("host ms:", " ".join("%.2f" % v for v in t))
        print("gpu: per sub-batch start,end (ms after e0):", [(round(e0.elapsed_time(ev[k][0]), 2), round(e0.elapsed_time(ev[k][1]), 2)) for k in range(n_sub)])
# single-call H2D bandwidth for reference
d = torch.empty_like(pinned, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(pinned, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D %.1f MB in %.2f ms = %.1f GB/s" % (pinned.numel() / 1e6, dt * 1e3, pinned.numel() / dt / 1e9))
