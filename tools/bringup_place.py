"""GPU bring-up of the placement kernels: raw counts and soft weights vs numpy.
Usage: python tools/bringup_place.py <n_snv> <n_cells> [--perf]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from scsommerclock_b200 import engine, synth

def oracle(codes, S, FP, FN):
    x1 = ((codes == 1) | (codes == 2)).astype(np.float32)
    x0 = (codes == 0).astype(np.float32)
    C1 = (x1 @ S.T.astype(np.float32)).astype(np.int64)      # exact: counts < 2^24
    C0 = (x0 @ S.T.astype(np.float32)).astype(np.int64)
    a1 = np.log1p(-FN) - np.log(FP)
    a0 = np.log(FN) - np.log1p(-FP)
    L = a1 * C1 + a0 * C0
    L -= L.max(axis=1, keepdims=True)
    P = np.exp(L)
    P /= P.sum(axis=1, keepdims=True)
    return C1, C0, P.sum(axis=0)

def main():
    n_snv, n_cells = int(sys.argv[1]), int(sys.argv[2])
    perf = "--perf" in sys.argv
    rng = np.random.default_rng(1)
    tree = synth.coalescent_tree(n_cells, rng)
    S = tree.clade_matrix()
    FN, FP, MS = 0.1, 0.01, 0.15
    stress = "--stress" in sys.argv
    if stress:       # dense random clade rows and genotypes: large counts in both planes
        S = (rng.random(S.shape) < 0.9).astype(np.uint8); S[-1] = 0
    ctx = engine.Context(0)
    print("ctx sm", ctx.sm_count, flush=True)
    bits, words = engine.pack_clades(S)
    if not perf:
        br = synth.throw_mutations(tree, n_snv, rng)
        codes = synth.observe(tree, br, FN, FP, MS, rng)
        if stress:
            codes = rng.choice([0, 1, 2, 3], size=codes.shape, p=[0.5, 0.3, 0.15, 0.05]).astype(np.uint8)
            codes[0] = 0; codes[1] = 1
        packed, pitch = engine.pack_genotypes(codes)
        pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
        print(pl.info(), flush=True)
        pl.set_genotypes_host(packed, pitch)
        pl.set_clades_host(bits, words)
        cnt = pl.debug_counts()
        C1, C0, Y = oracle(codes, S, FP, FN)
        ok1 = np.array_equal(cnt[:, :, 0], C1); ok0 = np.array_equal(cnt[:, :, 1], C0)
        print("counts exact:", ok1, ok0, flush=True)
        if not (ok1 and ok0):
            bad = np.argwhere(cnt[:, :, 0] != C1)
            print("first mismatches", bad[:10], cnt[:2, :8, 0], C1[:2, :8])
        y = pl.run_host(FP, FN)
        err = np.abs(y - Y) / np.maximum(np.abs(Y), 1e-300)
        print("Y max rel err %.3e  max abs err %.3e  sum %.6f vs %.6f" % (err[Y > 1e-6].max(), np.abs(y - Y).max(), y.sum(), Y.sum()), flush=True)
        keep = (codes == 1).any(axis=1).astype(np.uint8)
        pl.set_genotypes_host(packed, pitch, keep)
        y2 = pl.run_host(FP, FN)
        _, _, Y2 = oracle(codes[keep.astype(bool)], S, FP, FN)
        print("row_keep: max abs err %.3e" % np.abs(y2 - Y2).max(), flush=True)
        print("RESULT", "PASS" if (ok1 and ok0 and np.abs(y - Y).max() < 1e-4 * max(1, Y.max()) and np.abs(y2 - Y2).max() < 1e-4 * max(1, Y.max())) else "FAIL")
    else:
        import torch
        dev = torch.device("cuda:0")
        lo = torch.as_tensor(tree.lo[:-1], device=dev); hi = torch.as_tensor(tree.hi[:-1], device=dev)
        rate = torch.as_tensor(tree.branch_len[:-1], device=dev, dtype=torch.float32)
        t0 = time.time()
        packed, pitch, keep = synth.observe_packed_torch(lo, hi, torch.as_tensor(tree.leaf_rank), n_snv, FN, FP, MS, dev, rate=rate)
        torch.cuda.synchronize(); print("synth %.1fs" % (time.time() - t0), flush=True)
        pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
        print(pl.info(), flush=True)
        pl.set_genotypes(packed, pitch, keep)
        pl.set_clades_host(bits, words)
        torch.cuda.synchronize()
        for it in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); pl.run(FP, FN); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            flop = 2.0 * n_snv * S.shape[0] * n_cells
            print("run %d: %.3f ms  placements/s %.3e  alg TFLOP/s (both passes counted once each) %.1f" % (it, ms, n_snv * S.shape[0] / ms * 1e3, 2 * flop / ms / 1e9), flush=True)
        y = torch.empty(S.shape[0], dtype=torch.float64, device=dev)
        pl.run(FP, FN, y); torch.cuda.synchronize()
        print("sum Y", float(y.sum()), "kept", int(keep.sum()))

main()
