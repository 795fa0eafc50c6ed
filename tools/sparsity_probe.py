"""How sparse could pass 2 be?  After a full run, look at the per-(SNV, half tile) probability mass
and count the (row block, column tile) pairs that hold mass >= 2^-T for some SNV of the block, with
SNVs in file order and sorted by their best column."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from scsommerclock_b200 import engine, run_PT_test as P, synth
from scsommerclock_b200._cabi import lib, check

n_cells, n_snv = int(sys.argv[1]), int(sys.argv[2])
FPt = float(sys.argv[3]) if len(sys.argv) > 3 else bench.FP_TRUE
dev = torch.device("cuda", 0)
ctx = engine.Context(0)
ts = bench.make_tree(n_cells)
S = ts.clade_matrix()          # node-id order (leaves first); preorder variant below
bits, words = engine.pack_clades(S)
lo = torch.as_tensor(ts.lo[:-1], device=dev); hi = torch.as_tensor(ts.hi[:-1], device=dev)
rate = torch.as_tensor(ts.branch_len[:-1], device=dev, dtype=torch.float32)
packed, pitch, keep = synth.observe_packed_torch(lo, hi, torch.as_tensor(ts.leaf_rank), n_snv, bench.FN_TRUE, FPt, bench.MS_TRUE, dev, seed=1000, rate=rate)
FP, FN = FPt, bench.FN_TRUE / 2
# column orders to try: node-id order and tree pre-order (subtrees contiguous)
n = n_cells
pre = []
stack = [2 * n - 2]
while stack:
    v = stack.pop(); pre.append(v)
    if v >= n: stack.append(ts.right[v - n]); stack.append(ts.left[v - n])
orders = {"node_id": np.arange(2 * n - 1), "preorder": np.array(pre)}
for oname, order in orders.items():
    Sx = ts.clade_matrix(order)
    bits, words = engine.pack_clades(Sx)
    pl = engine.Placement(ctx, n_snv, n_cells, Sx.shape[0])
    pl.set_clades_host(bits, words); pl.set_genotypes(packed, pitch, keep)
    y = pl.run_host(FP, FN)
    info = pl.info(); M_pad = info["M_pad"]; parts = 2 * (info["N_pad"] // 128)
    stats = np.empty((parts, M_pad, 2), dtype=np.uint32); rowref = np.empty((M_pad, 2), dtype=np.uint32)
    check(lib().scs_placement_debug_stats(pl._h, stats.ctypes.data_as(C.c_void_p), rowref.ctypes.data_as(C.c_void_p)))
    a1 = (np.log1p(-FN) - np.log(FP)) / np.log(2); a0 = (np.log(FN) - np.log1p(-FP)) / np.log(2)
    kept = keep.cpu().numpy().astype(bool)
    rows = np.flatnonzero(kept)
    ref = rowref[rows, 0]; L = rowref[rows, 1].view(np.float32).astype(np.float64)
    best = a1 * (ref & 0xFFFF) + a0 * (ref >> 16) + L
    sp = stats[:, rows, :]
    sm = sp[:, :, 1].view(np.float32).astype(np.float64)
    with np.errstate(divide="ignore"):
        mass = np.log2(sm) + a1 * (sp[:, :, 0] & 0xFFFF) + a0 * (sp[:, :, 0] >> 16) - best[None, :]      # [parts, rows] log2 of the part's probability
    tile_mass = np.maximum(mass[0::2], mass[1::2])          # per column tile
    ntile = tile_mass.shape[0]
    best_tile = tile_mass.argmax(axis=0)
    for T in (30, 50):
        flag = tile_mass >= -T                              # [tiles, rows]
        per_row = flag.sum(axis=0)
        for sname, perm in (("file order", np.arange(rows.size)), ("sorted by best tile", np.argsort(best_tile, kind="stable"))):
            f = flag[:, perm]
            nb = (rows.size + 127) // 128
            pad = nb * 128 - rows.size
            fb = np.concatenate([f, np.zeros((ntile, pad), bool)], axis=1).reshape(ntile, nb, 128).any(axis=2)
            print("%s cols, T=%d, %s: tiles per row mean %.2f (max %d); (row block, tile) pairs needed %.1f%% of %d" % (oname, T, sname, per_row.mean(), per_row.max(), 100.0 * fb.mean(), fb.size), flush=True)
    pl.close()
