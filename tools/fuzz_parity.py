#!/usr/bin/env python3
"""One-off randomized parity sweep on a GPU (not part of the test suite): shapes around the boundaries of the round-2 code paths.
  (1) single datasets, interval path against the GEMM: cell counts around the pre-pass sub-block / thread-count boundaries;
  (2) batches of small trees (iv5) against the GEMM: cell counts around the lane-width boundaries, ragged groups;
  (3) the cluster LRT kernel against the CTA kernel on random tree sizes;
  (4) per-cell FN rates (scs_place_percell) against the float64 prefix-sum oracle (oracle/pt_oracle.map_mutations_gt_interval).
`python tools/fuzz_parity.py SEED SECONDS [percell]`: with a third argument only (4) runs."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from scsommerclock_b200 import engine, synth
from test_gpu_parity import _synthetic_lrt_problem

ctx = engine.Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
bad = []

def single(n_cells, n_snv):
    tree = synth.coalescent_tree(n_cells, rng)
    S = tree.clade_matrix()
    perm = rng.permutation(n_cells)
    S = np.ascontiguousarray(S[:, perm])
    codes = synth.observe(tree, synth.throw_mutations(tree, n_snv, rng), 0.1, 0.01, float(rng.uniform(0.0, 0.4)), rng)[:, perm]
    keep = (rng.random(n_snv) < 0.9).astype(np.uint8)
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    out = {}
    for algo in ("interval", "gemm"):
        os.environ["SCS_PLACE_ALGO"] = algo
        pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
        pl.set_clades_host(bits, words)
        d = torch.as_tensor(packed).cuda()
        active = np.ones(n_cells, dtype=np.uint8)
        pl.prepare(d, pitch, active, True)
        y = torch.zeros(S.shape[0], dtype=torch.float64, device="cuda")
        pl.run(0.01, 0.07, y)
        nk, miss = pl.prepare_result()
        out[algo] = (y.cpu().numpy(), nk, miss, pl.info()["algo"])
        pl.close()
    os.environ.pop("SCS_PLACE_ALGO")
    a, b = out["interval"], out["gemm"]
    ok = a[3] == 2 and b[3] == 1 and a[1] == b[1] and np.allclose(a[2], b[2], rtol=0, atol=1e-12) and np.abs(a[0] - b[0]).max() <= 5e-6 * max(1.0, np.abs(b[0]).max())
    if not ok:
        bad.append(("single", n_cells, n_snv, a[1], b[1], float(np.abs(a[0] - b[0]).max())))
    return ok

def batch(n_cells, R):
    names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
    columns = names + ["healthycell"]
    snvs = [int(v) for v in rng.integers(1, 300, size=R)]
    trees = [synth.coalescent_tree(n_cells, rng) for _ in range(R)]
    newicks = [t.newick(names, "healthycell", scale=0.05) for t in trees]
    mats = []
    for t, m in zip(trees, snvs):
        codes = synth.observe(t, synth.throw_mutations(t, m, rng), 0.1, 0.02, 0.2, rng)
        codes = np.concatenate([codes, np.zeros((m, 1), dtype=np.uint8)], axis=1)
        mats.append(engine.pack_genotypes(codes)[0])
    d_gt = torch.as_tensor(np.concatenate(mats, axis=0)).cuda()
    FP = rng.uniform(0.005, 0.03, size=R); FN = rng.uniform(0.03, 0.12, size=R)
    got = {}
    for algo in ("auto", "gemm"):
        os.environ["SCS_PLACE_ALGO"] = algo
        sw = engine.ReplicateSweep(ctx, snvs, columns, [100.0, 1000.0])
        sw.parse_trees(newicks); sw.enqueue(d_gt, FP, FN); out, nk, st = sw.results()
        got[algo] = (sw.d_y.cpu().numpy().copy(), out.copy(), nk.copy(), sw.place.info()["algo"])
        sw.close()
    os.environ.pop("SCS_PLACE_ALGO")
    a, b = got["auto"], got["gemm"]
    ok = a[3] == 3 and b[3] == 1 and np.array_equal(a[2], b[2]) and np.abs(a[0] - b[0]).max() <= 5e-6 * max(1.0, np.abs(b[0]).max())
    okl = (a[1][:, :, 7] == 0) & (b[1][:, :, 7] == 0)
    ok = ok and np.allclose(a[1][:, :, 0][okl], b[1][:, :, 0][okl], rtol=1e-4, atol=1e-4)
    if not ok:
        bad.append(("batch", n_cells, R, float(np.abs(a[0] - b[0]).max())))
    return ok

def lrt(n_leaves, n_prob):
    probs = [_synthetic_lrt_problem(n_leaves, int(rng.integers(1 << 30)), sparse=bool(rng.integers(2))) for _ in range(n_prob)]
    Ys, ws, chs = [p[0] for p in probs], [p[1] for p in probs], [p[2] for p in probs]
    os.environ["SCS_LRT_KERNEL"] = "tree"
    os.environ["SCS_LRT_CLUSTER"] = str(int(rng.choice([1, 2, 4, 8])))
    a, xa = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    os.environ["SCS_LRT_KERNEL"] = "cta"
    b, xb = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    os.environ.pop("SCS_LRT_KERNEL"); os.environ.pop("SCS_LRT_CLUSTER")
    ok = np.array_equal(a[:, 7], b[:, 7])
    m = a[:, 7] == 0
    ok = ok and np.array_equal(a[m, 1], b[m, 1]) and np.array_equal(a[m, 2], b[m, 2]) and np.allclose(a[m, 0], b[m, 0], rtol=1e-8, atol=1e-8)
    if not ok:
        bad.append(("lrt", n_leaves, n_prob, a[:, [0, 2, 7]].tolist(), b[:, [0, 2, 7]].tolist()))
    return ok

def percell(n_cells, n_snv):
    from oracle import pt_oracle as O
    tree = synth.coalescent_tree(n_cells, rng)
    perm = rng.permutation(n_cells)
    S = np.ascontiguousarray(tree.clade_matrix()[:, perm])
    codes = np.ascontiguousarray(synth.observe(tree, synth.throw_mutations(tree, n_snv, rng), 0.1, 0.01, float(rng.uniform(0.0, 0.5)), rng)[:, perm])
    keep = (rng.random(n_snv) < 0.9).astype(np.uint8)
    lo = 10.0 ** rng.uniform(-5, -1)
    rates = np.exp(rng.uniform(np.log(lo), np.log(0.95), n_cells))
    FP = float(10.0 ** rng.uniform(-4, -0.5))
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    y = torch.zeros(S.shape[0], dtype=torch.float64, device="cuda")
    engine.place_percell(ctx, torch.as_tensor(packed).cuda(), pitch, n_snv, n_cells, torch.as_tensor(keep).cuda(), bits, words, FP, rates, y)
    muts = np.where(codes == 3, np.nan, codes.astype(float))[keep.astype(bool)]
    want = O.map_mutations_gt_interval(muts, S.astype(float), FP, rates) if muts.shape[0] else np.zeros(S.shape[0])
    err = float((np.abs(y.cpu().numpy() - want) / np.maximum(1.0, np.abs(want))).max())
    if err > 2e-5:
        bad.append(("percell", n_cells, n_snv, FP, lo, err))
    return err <= 2e-5

t0 = time.time()
if len(sys.argv) > 3:
    n4 = 0
    edge = [2, 3, 31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 512, 513, 1000, 2047, 2048, 2049, 4095, 4097]
    while time.time() - t0 < budget:
        percell(edge[n4] if n4 < len(edge) else int(rng.integers(2, 3000)), int(rng.integers(1, 500))); n4 += 1
    print("percell cases", n4, "failures", len(bad))
    for b in bad[:20]:
        print(str(b)[:400])
    sys.exit(0)
n = [0, 0, 0]
edge_cells = [100, 127, 128, 129, 255, 256, 257, 496, 511, 512, 513, 1008, 1024, 1025, 1535, 2047, 2048, 2049, 3071, 4095, 4096, 4097]
k = 0
while time.time() - t0 < budget:
    c = edge_cells[k % len(edge_cells)] if k < 2 * len(edge_cells) else int(rng.integers(100, 6000))
    single(c, int(rng.integers(1, 700))); n[0] += 1
    cb = [2, 3, 15, 16, 17, 112, 127, 128, 129, 240, 255, 256, 257, 400, 511][k % 15] if k < 30 else int(rng.integers(2, 512))
    batch(cb, int(rng.integers(1, 12))); n[1] += 1
    lrt(int(rng.integers(257, 4000)), int(rng.integers(1, 6))); n[2] += 1
    k += 1
print("cases", n, "failures", len(bad))
for b in bad[:20]:
    print(str(b)[:400])
