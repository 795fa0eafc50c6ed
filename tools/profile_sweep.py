"""Replicate-sweep leg alone (2048 x 100 cells x 10k SNVs), for ncu captures of the
small-tree placement passes and the batched LRT kernel.  Run under gpurun."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from scsommerclock_b200 import engine  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = engine.Context(0)
dev = torch.device("cuda:0")
rep = bench.replicate_sweep(ctx, dev, 0, 1, R, 100, 10000, 2, 1)
print(json.dumps(rep))
