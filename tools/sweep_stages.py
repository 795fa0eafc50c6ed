#!/usr/bin/env python3
"""Where the end-to-end replicate sweep spends its time: device time stamps (CUDA events) between the calls of
ReplicateSweep.enqueue, for sub-batches pipelined on their own streams exactly like bench.py's e2e leg.
usage: python tools/sweep_stages.py [n_sub]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from scsommerclock_b200 import engine, synth
from scsommerclock_b200.engine import lib, check, _dev_ptr, _stream_handle, ptr

n_cells, n_snv, R = 100, 10000, 1250
n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 4
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
NM = ["H2D matrix", "H2D masks", "gt_reduce", "set_clades", "expand", "placement", "sweep_model", "lrt", "D2H"]


def enqueue_marked(sw, src, st):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
    k = [0]
    def mark():
        ev[k[0]].record(st); k[0] += 1
    FP = np.full(sw.G, 0.01); FN = np.full(sw.G, 0.05)
    sw.FP, sw.FN = FP, FN
    with torch.cuda.stream(st):
        mark()
        if sw.d_gt is None:
            sw.d_gt = torch.empty((sw.total, sw.pitch), dtype=torch.uint8, device=sw.dev)
        sw.d_gt.copy_(src, non_blocking=True); mark()
        sw.d_bits.copy_(sw.h_bits, non_blocking=True); sw.d_children.copy_(sw.h_children, non_blocking=True); sw.d_info.copy_(sw.h_info, non_blocking=True); mark()
        sh = _stream_handle(st)
        check(lib().scs_gt_reduce_groups(ctx._h, _dev_ptr(sw.d_gt), int(sw.pitch), sw.G, _dev_ptr(sw.d_row0), int(sw.group_snvs.max()), sw.n_cols, _dev_ptr(sw.d_amask),
                                         int(sw.has_outg), _dev_ptr(sw.d_keep), _dev_ptr(sw.d_nkept), _dev_ptr(sw.d_missing), sh)); mark()
        sw.place.set_clades(sw.d_bits, sw.words, stream=st); mark()
        sw.place.set_genotypes(sw.d_gt, sw.pitch, sw.d_keep, stream=st); mark()
        sw.place.run(FP, FN, sw.d_y, stream=st); mark()
        check(lib().scs_sweep_model(ctx._h, sw.G, sw.n_rows, sw.n_cols, int(sw.words), _dev_ptr(sw.d_bits), _dev_ptr(sw.d_children), _dev_ptr(sw.d_info), _dev_ptr(sw.d_nkept),
                                    _dev_ptr(sw.d_missing), ptr(FP), ptr(FN), ptr(sw.w_maxs), sw.n_w, _dev_ptr(sw.d_y), sw.n_br, sw._lrt, _dev_ptr(sw.d_status), _dev_ptr(sw.d_work), sh)); mark()
        check(lib().scs_lrt_plan_run(sw._lrt, _dev_ptr(sw.d_y), _dev_ptr(sw.d_out), None, sh)); mark()
        sw.h_out.copy_(sw.d_out, non_blocking=True); sw.h_nkept.copy_(sw.d_nkept, non_blocking=True); sw.h_status.copy_(sw.d_status, non_blocking=True); mark()
        sw.done.record(st)
    return ev


cuts = [R * k // n_sub for k in range(n_sub + 1)]
subs = [engine.ReplicateSweep(ctx, [n_snv] * (cuts[k + 1] - cuts[k]), columns, bench.W_MAXS) for k in range(n_sub)]
streams = [torch.cuda.Stream(device=dev) for _ in range(n_sub)]
for it in range(4):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    for s_ in streams:
        s_.wait_stream(torch.cuda.current_stream())
    T = [time.perf_counter()]
    evs = []
    for k, sub in enumerate(subs):
        sub.parse_trees(newicks[cuts[k]:cuts[k + 1]]); T.append(time.perf_counter())
        evs.append(enqueue_marked(sub, pinned[cuts[k] * n_snv:cuts[k + 1] * n_snv], streams[k])); T.append(time.perf_counter())
    rows = []
    for sub in subs:
        rows.extend(sub.rows()); T.append(time.perf_counter())
    torch.cuda.synchronize()
    if it == 3:
        print("host ms:", " ".join("%.2f" % ((t - T[0]) * 1e3) for t in T))
        for k in range(n_sub):
            ts = [e0.elapsed_time(e) for e in evs[k]]
            print("sub %d: start %.2f | " % (k, ts[0]) + ", ".join("%s ->%.2f" % (NM[i], ts[i + 1]) for i in range(9)))
