"""Timing harness for the placement passes on random data (no correctness claim):
python tools/exp_place.py [cells] [snvs] -> pass1 / merge / pass2 / reduce ms, pass-2 mode."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from scsommerclock_b200 import engine  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_snv = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
n_rows = 2 * n_cells
ctx = engine.Context(0)
dev = torch.device("cuda:0")
pitch = (n_cells + 3) // 4
pitch = (pitch + 15) // 16 * 16
g = torch.Generator(device=dev)
g.manual_seed(1)
packed = torch.randint(0, 256, (n_snv, pitch), dtype=torch.uint8, device=dev, generator=g)
packed &= torch.randint(0, 256, (n_snv, pitch), dtype=torch.uint8, device=dev, generator=g)   # mostly wild type
rng = np.random.default_rng(0)
words = (n_cells + 31) // 32
bits = rng.integers(0, 2**32, size=(n_rows, words), dtype=np.uint64).astype(np.uint32)
bits &= rng.integers(0, 2**32, size=(n_rows, words), dtype=np.uint64).astype(np.uint32)
pl = engine.Placement(ctx, n_snv, n_cells, n_rows)
pl.set_clades_host(bits, words)
pl.set_genotypes(packed, pitch, None, stream=torch.cuda.current_stream())
y = torch.zeros(n_rows, dtype=torch.float64, device=dev)
for it in range(4):
    pl.run(0.01, 0.1, y, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    t = pl.last_times()
print("cells %d snvs %d pass1 %.2f merge %.2f pass2 %.2f reduce %.3f ms  mode %s" % (n_cells, n_snv, t[0], t[1], t[2], t[3], pl.info()["pass2_mode"]))
