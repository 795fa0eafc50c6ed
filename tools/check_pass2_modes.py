"""The three placement variants (GEMM with a second pass, GEMM with the count cache, interval path) on the bench's own data path
(coalescent tree, device-generated genotypes, row filter), at bench sizes:
python tools/check_pass2_modes.py [cells] [snvs ...]
Prints per size the kernel times of both modes, the number of non-finite / mismatching clade rows
and where they sit (column tile), so a size-dependent fault can be located from one run."""
import os
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from scsommerclock_b200 import engine, run_PT_test as P, synth  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
sizes = [int(a) for a in sys.argv[2:]] or [250000]
ctx = engine.Context(0)
dev = torch.device("cuda:0")
tree_s = bench.make_tree(n_cells)
names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
columns = names + ["healthycell"]
tmp = tempfile.mkdtemp(prefix="scs_chk_")
tree_file = os.path.join(tmp, "t.newick")
with open(tree_file, "w") as f:
    f.write(tree_s.newick(names, "healthycell", scale=0.05))
tree, outg, _, _ = P.read_tree(tree_file, np.array(columns))
active = np.array([c != outg for c in columns], dtype=np.uint8)
bits, words, nodes = P._clade_bits(tree, columns, active)
n_rows, n_cols = bits.shape[0], len(columns)
lo = torch.as_tensor(tree_s.lo[:-1], device=dev)
hi = torch.as_tensor(tree_s.hi[:-1], device=dev)
rate = torch.as_tensor(tree_s.branch_len[:-1], device=dev, dtype=torch.float32)
stream = torch.cuda.current_stream()
bad = 0
for n_snv in sizes:
    packed, pitch, _ = synth.observe_packed_torch(lo, hi, torch.as_tensor(tree_s.leaf_rank), n_snv, bench.FN_TRUE, bench.FP_TRUE,
                                                  bench.MS_TRUE, dev, seed=1000, rate=rate, extra_cols=1)
    row_keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
    engine.gt_reduce(ctx, packed, pitch, n_snv, n_cols, active, True, row_keep)
    res = {}
    for mode in ("recompute", "cache", "interval"):
        os.environ["SCS_PLACE_PASS2"] = "cache" if mode == "interval" else mode
        os.environ["SCS_PLACE_ALGO"] = "auto" if mode == "interval" else "gemm"
        pl = engine.Placement(ctx, n_snv, n_cols, n_rows)
        pl.set_clades_host(bits, words)
        pl.bind_genotypes(packed, pitch, row_keep, stream=stream)
        y = torch.zeros(n_rows, dtype=torch.float64, device=dev)
        for _ in range(3):
            pl.run(bench.FP_TRUE, bench.FN_TRUE / 2, y, stream=stream)
            torch.cuda.synchronize()
        t = pl.last_times()
        info = pl.info()
        res[mode] = y.cpu().numpy()
        print("cells %d snvs %d %-9s pass1|permute %.2f merge %.2f pass2|interval %.2f reduce %.3f ms  algo %d pass2_mode %d runs %d run_len %d free %.1f GB  interval grid %d x %d thr, %d B smem, %d runs" % (
            n_cells, n_snv, mode, t[0], t[1], t[2], t[3], info["algo"], info["pass2_mode"], info["runs"], info["run_len"],
            torch.cuda.mem_get_info()[0] / 1e9, info["interval_grid"], info["interval_threads"], info["interval_smem_bytes"],
            info["interval_runs"]), flush=True)
        pl.close()
    for other in ("cache", "interval"):
        a, b = res["recompute"], res[other]
        nf = ~np.isfinite(b)
        rel = np.abs(a - b) / np.maximum(1.0, np.abs(a))
        rel[nf] = np.inf
        wrong = np.nonzero(rel > 2e-5)[0]
        print("  %-8s vs recompute: kept %d  sum recompute %.6f %s %.6f  non-finite %d / %d  rows off %d  max rel %.3g" % (
            other, int(row_keep.sum().item()), a[np.isfinite(a)].sum(), other, b[~nf].sum(), int((~np.isfinite(a)).sum()), int(nf.sum()),
            wrong.size, float(rel[~nf].max()) if (~nf).any() else -1), flush=True)
        if wrong.size:
            bad += 1
            tiles, cnt = np.unique(wrong // 128, return_counts=True)
            print("  off rows by column tile:", dict(zip(tiles.tolist()[:20], cnt.tolist()[:20])), " first rows", wrong[:10].tolist(),
                  " values", a[wrong[:5]].tolist(), b[wrong[:5]].tolist(), flush=True)
    del packed, row_keep
    torch.cuda.empty_cache()
sys.exit(1 if bad else 0)
