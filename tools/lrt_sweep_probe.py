#!/usr/bin/env python3
"""Runs the reference-recorded LRT sweep (tests/golden/lrt_sweep.npz) through the GPU solver at a ladder of barrier
parameters 0.1 * 0.2^k and stores every result: which point of SciPy's central path did the reference stop at?"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from scsommerclock_b200 import engine  # noqa: E402

g = np.load(os.path.join(REPO, "tests", "golden", "lrt_sweep.npz"))
off = g["off"]
n = len(off) - 1
Ys = [g["Y"][off[i]:off[i + 1]] for i in range(n)]
ws = [g["w"][off[i]:off[i + 1]] for i in range(n)]
chs = [g["child_int"][off[i]:off[i + 1]] for i in range(n)]
ctx = engine.Context(0)
res = {}
for k in range(3, 15):
    os.environ["SCS_LRT_MU"] = repr(0.1 * 0.2 ** k)
    out, xs = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    res["out_k%d" % k] = out
    res["x_k%d" % k] = np.concatenate(xs)
np.savez_compressed(os.path.join(REPO, "gpurun_out", "r2_lrt_probe.npz"), **res)
print("saved", len(res))
