#!/usr/bin/env python3
"""Time the LRT kernels on big trees: one CTA per problem (round 1) against the cluster-per-problem kernel.
usage: python tools/lrt_timing.py [n_leaves ...]"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from scsommerclock_b200 import engine
from test_gpu_parity import _synthetic_lrt_problem

ctx = engine.Context(0)
sizes = [int(a) for a in sys.argv[1:]] or [1000, 10000]
for n_leaves in sizes:
    for n_prob in (10,):
        probs = [_synthetic_lrt_problem(n_leaves, 7 * n_leaves + i) for i in range(n_prob)]
        Ys, ws, chs = [p[0] for p in probs], [p[1] for p in probs], [p[2] for p in probs]
        src = torch.as_tensor(np.concatenate(Ys), device="cuda")
        for kern, cluster, start in (("cta", "1", "0"), ("tree", "1", "0"), ("tree", "4", "0"), ("tree", "8", "0"), ("tree", "8", "1")):
            os.environ["SCS_LRT_KERNEL"] = kern
            os.environ["SCS_LRT_CLUSTER"] = cluster
            os.environ["SCS_LRT_START"] = start
            plan = engine.LrtPlan(ctx, chs)
            plan.set_weights(ws)
            out = torch.zeros((n_prob, 8), dtype=torch.float64, device="cuda")
            for _ in range(3):
                plan.run(src, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                plan.run(src, out)
            e1.record()
            torch.cuda.synchronize()
            o = out.cpu().numpy()
            print("leaves %6d problems %3d kernel %-4s cluster %s start %s: %8.3f ms per launch, iterations %s, status %s, LR[0] %.9f"
                  % (n_leaves, n_prob, kern, cluster, start, e0.elapsed_time(e1) / 10, [int(v) for v in o[:, 6]], sorted(set(int(v) for v in o[:, 7])), o[0, 0]))
            if kern == 'tree' and cluster in ('1', '8'):
                os.environ['SCS_LRT_PROF'] = '1'
                plan.run(src, out)
                torch.cuda.synchronize()
                del os.environ['SCS_LRT_PROF']
            plan.close()
