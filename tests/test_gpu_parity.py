"""GPU parity tests (run with ``-m gpu`` on a B200): every CUDA entry point,
called through the C ABI, against the CPU oracle and the golden fixtures
recorded from the reference.

Tolerances (stated once, used below):
  * genotype matrix, clade masks, integer counts, dof, on_bound, decisions: exact;
  * soft branch weights Y: |gpu - ref| <= 2e-5 * max(1, |ref|)  (fp32 epilogue on
    exact integer differences, fp64 accumulation across SNV runs);
  * branch weights: relative 1e-10 (fp64, different summation order);
  * -2logLR: relative 1e-4 (the reference's own SciPy solve is only converged to
    about that), p-value: relative 2e-3.
"""
import os

import numpy as np
import pytest

from conftest import case_inputs, golden_path
from oracle import pt_oracle as O
from scsommerclock_b200 import engine, run_PT_test as P, synth

pytestmark = pytest.mark.gpu

CASES = ["example_clock", "example_noclock", "syn_a", "syn_b", "syn_c", "syn_c_incl", "syn_d_clock", "syn_d_noclock", "syn_e_mismatch",
         "syn_f_scite_log", "syn_g_scite_csv", "syn_h_path_rates"]
SYN = [c for c in CASES if c.startswith("syn")]
Y_TOL = 2e-5


def load(name):
    return np.load(golden_path(name), allow_pickle=False)


def oracle_counts(codes, S):
    x1 = ((codes == 1) | (codes == 2)).astype(np.int64)
    x0 = (codes == 0).astype(np.int64)
    return x1 @ S.T.astype(np.int64), x0 @ S.T.astype(np.int64)


def make_problem(n_snv, n_cells, seed, FN=0.1, FP=0.01, MS=0.15):
    rng = np.random.default_rng(seed)
    tree = synth.coalescent_tree(n_cells, rng) if n_cells > 1 else None
    if tree is None:
        S = np.array([[1], [0]], dtype=np.uint8)
        codes = rng.integers(0, 4, size=(n_snv, 1)).astype(np.uint8)
        codes[codes == 2] = 1
    else:
        S = tree.clade_matrix()
        codes = synth.observe(tree, synth.throw_mutations(tree, n_snv, rng), FN, FP, MS, rng)
    return codes, S


def assert_soft_close(got, want):
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    assert err.max() <= Y_TOL, (err.max(), int(err.argmax()))


# ------------------------------------------------------------ placement GEMM
@pytest.mark.parametrize("mode", ["radix", "planes"])
@pytest.mark.parametrize("n_snv,n_cells", [(1, 2), (5, 1), (127, 3), (128, 64), (129, 65), (300, 50), (1000, 200),
                                            (2000, 300), (700, 513)])
def test_counts_bit_exact_and_soft_weights(ctx, monkeypatch, mode, n_snv, n_cells):
    monkeypatch.setenv("SCS_PLACE_MODE", mode)      # one e5m2 radix plane (kind::f8f6f4) / two u8 planes (kind::i8)
    monkeypatch.setenv("SCS_PLACE_ALGO", "gemm")    # tree-shaped masks would otherwise take the interval path (tested below)
    codes, S = make_problem(n_snv, n_cells, seed=n_snv * 7 + n_cells)
    FP, FN = 0.01, 0.1
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
    assert pl.info()["mode"] == (1 if mode == "radix" else 0)
    pl.set_genotypes_host(packed, pitch)
    pl.set_clades_host(bits, words)
    cnt = pl.debug_counts()
    C1, C0 = oracle_counts(codes, S)
    assert np.array_equal(cnt[:, :, 0], C1) and np.array_equal(cnt[:, :, 1], C0)      # integer work: bit exact
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    want = O.map_mutations_gt(muts, S.astype(float), FP, FN)
    got = pl.run_host(FP, FN)
    assert_soft_close(got, want)
    # one-shot host entry point gives the same numbers
    got2 = engine.map_mutations(ctx, packed, pitch, bits, words, n_cells, FP, FN)
    assert np.array_equal(got, got2)
    pl.close()


def test_every_partial_column_tile_width(ctx, monkeypatch):
    """The single-K-chunk radix epilogue reads a tile in pipelined 16/32-column loads and skips
    what lies beyond the last valid clade row: cover every boundary of the last column tile."""
    monkeypatch.setenv("SCS_PLACE_ALGO", "gemm")     # (one or two random rows can happen to be tree shaped)
    rng = np.random.default_rng(99)
    n_snv, n_cells = 300, 40
    codes = rng.choice([0, 1, 2, 3], size=(n_snv, n_cells), p=[0.6, 0.2, 0.1, 0.1]).astype(np.uint8)
    packed, pitch = engine.pack_genotypes(codes)
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    for n_rows in [1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65, 79, 80, 81, 95, 96, 97, 111, 112, 113, 127, 128, 129, 161, 193, 255]:
        S = (rng.random((n_rows, n_cells)) < 0.3).astype(np.uint8)
        S[-1] = 0
        bits, words = engine.pack_clades(S)
        pl = engine.Placement(ctx, n_snv, n_cells, n_rows)
        assert pl.info()["mode"] == 1 and pl.info()["k_chunks"] == 1
        pl.set_genotypes_host(packed, pitch)
        pl.set_clades_host(bits, words)
        C1, C0 = oracle_counts(codes, S)
        cnt = pl.debug_counts()
        assert np.array_equal(cnt[:, :, 0], C1) and np.array_equal(cnt[:, :, 1], C0), n_rows
        assert_soft_close(pl.run_host(0.02, 0.2), O.map_mutations_gt(muts, S.astype(float), 0.02, 0.2))
        pl.close()


@pytest.mark.parametrize("pass2", ["recompute", "cache"])
@pytest.mark.parametrize("mode", ["radix", "planes"])
@pytest.mark.parametrize("n_snv,n_cells", [(300, 3968), (260, 4100), (140, 8100), (130, 10000), (130, 12500), (129, 15872)])
def test_counts_bit_exact_dense_many_cells(ctx, monkeypatch, mode, pass2, n_snv, n_cells):
    """Worst case for the radix packing: dense random clade rows and genotypes put both
    counts of every K chunk (3968 cells) near their maximum; 1, 2, 3 and 4 chunks (with 4 chunks
    a tile fills the whole TMEM slot pool and the epilogue no longer overlaps the next tile)."""
    monkeypatch.setenv("SCS_PLACE_MODE", mode)
    monkeypatch.setenv("SCS_PLACE_PASS2", pass2)     # second GEMM pass / sweep over the cached integer counts
    rng = np.random.default_rng(n_cells)
    n_rows = 256 if n_cells > 5000 else 384
    S = (rng.random((n_rows, n_cells)) < 0.9).astype(np.uint8)
    S[-1] = 0
    S[0] = 1
    codes = rng.choice([0, 1, 2, 3], size=(n_snv, n_cells), p=[0.5, 0.3, 0.15, 0.05]).astype(np.uint8)
    codes[0] = 0
    codes[1] = 1
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    pl = engine.Placement(ctx, n_snv, n_cells, n_rows)
    pl.set_genotypes_host(packed, pitch)
    pl.set_clades_host(bits, words)
    cnt = pl.debug_counts()
    x1 = ((codes == 1) | (codes == 2)).astype(np.float32)
    x0 = (codes == 0).astype(np.float32)
    C1 = (x1 @ S.T.astype(np.float32)).astype(np.int64)            # fp32 BLAS is exact below 2^24
    C0 = (x0 @ S.T.astype(np.float32)).astype(np.int64)
    assert C0.max() >= 0.85 * min(n_cells, 3968) * 0.5
    assert np.array_equal(cnt[:, :, 0], C1) and np.array_equal(cnt[:, :, 1], C0)
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    assert_soft_close(pl.run_host(0.01, 0.1), O.map_mutations_gt(muts, S.astype(float), 0.01, 0.1))
    assert pl.info()["pass2_mode"] == (2 if pass2 == "cache" else 1)
    pl.close()


@pytest.mark.parametrize("n_snv,n_cells,n_rows", [(1000, 200, 400), (5000, 2600, 700), (300, 70, 1)])
def test_count_cache_pass2_matches_recompute(ctx, monkeypatch, n_snv, n_cells, n_rows):
    """Pass 2 from the cached integer counts (default for one dataset with >= 2560 cells) against
    the second GEMM pass: same soft weights, with row filter, ragged tiles and several runs."""
    monkeypatch.setenv("SCS_PLACE_ALGO", "gemm")     # (a single all-zero row is tree shaped)
    rng = np.random.default_rng(n_cells)
    S = (rng.random((n_rows, n_cells)) < 0.2).astype(np.uint8)
    S[-1] = 0
    codes = rng.choice([0, 1, 2, 3], size=(n_snv, n_cells), p=[0.7, 0.12, 0.08, 0.1]).astype(np.uint8)
    keep = (rng.random(n_snv) < 0.8).astype(np.uint8)
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    got = {}
    for pass2 in ("recompute", "cache", None):
        if pass2 is None:
            monkeypatch.delenv("SCS_PLACE_PASS2")
        else:
            monkeypatch.setenv("SCS_PLACE_PASS2", pass2)
        monkeypatch.setenv("SCS_RUN_LEN", "4")          # several runs of row blocks even at this size
        pl = engine.Placement(ctx, n_snv, n_cells, n_rows)
        pl.set_genotypes_host(packed, pitch, keep)
        pl.set_clades_host(bits, words)
        got[pass2] = pl.run_host(0.01, 0.1)
        want_mode = {"recompute": 1, "cache": 2, None: 2 if n_cells >= 2560 else 1}[pass2]
        assert pl.info()["pass2_mode"] == want_mode
        pl.close()
    muts = np.where(codes == 3, np.nan, codes.astype(float))[keep.astype(bool)]
    want = O.map_mutations_gt(muts, S.astype(float), 0.01, 0.1)
    for v in got.values():
        assert_soft_close(v, want)
    np.testing.assert_allclose(got["cache"], got["recompute"], rtol=2e-6, atol=1e-9)


# ------------------------------------------------------------ interval path
def shuffled_problem(n_snv, n_cells, seed):
    """make_problem with the cell columns shuffled, so that the depth-first order is not the column order."""
    codes, S = make_problem(n_snv, n_cells, seed)
    perm = np.random.default_rng(seed + 1).permutation(n_cells)
    return np.ascontiguousarray(codes[:, perm]), np.ascontiguousarray(S[:, perm])


@pytest.mark.parametrize("n_snv,n_cells", [(1, 2), (127, 3), (300, 50), (1000, 200), (700, 513), (400, 1100), (150, 4200), (60, 10000),
                                            (40, 12500)])
def test_interval_path_matches_oracle_and_gemm(ctx, monkeypatch, n_snv, n_cells):
    """Tree-shaped clade masks: counts as differences of per-SNV prefix counts (no GEMM).  Same soft
    weights as the oracle (where it finishes in seconds) and as the tcgen05 GEMM on the same inputs;
    256 / 512 / 1024 threads per CTA, with and without the count cache in shared memory (12,500 cells)."""
    codes, S = shuffled_problem(n_snv, n_cells, seed=n_snv + 3 * n_cells)
    keep = np.ones(n_snv, dtype=np.uint8)
    if n_snv > 4:
        keep[::5] = 0
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    got = {}
    for algo in ("auto", "gemm"):
        # "auto" takes the interval path from 100 cells on (where it is faster), "interval" at any size
        monkeypatch.setenv("SCS_PLACE_ALGO", algo if (algo == "gemm" or n_cells >= 100) else "interval")
        pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
        pl.set_clades_host(bits, words)
        pl.set_genotypes_host(packed, pitch, keep)
        got[algo] = pl.run_host(0.01, 0.1)
        assert pl.info()["algo"] == (2 if algo == "auto" else 1)
        if algo == "auto":
            assert np.array_equal(pl.run_host(0.01, 0.1), got[algo])          # fixed reduction order
            if n_cells <= 1100:
                cnt = pl.debug_counts()                                      # builds the GEMM operands on demand
                C1, C0 = oracle_counts(codes, S)
                assert np.array_equal(cnt[:, :, 0], C1) and np.array_equal(cnt[:, :, 1], C0)
        pl.close()
    assert_soft_close(got["auto"], got["gemm"])
    np.testing.assert_allclose(got["auto"], got["gemm"], rtol=5e-6, atol=1e-7)
    if n_cells <= 513:
        muts = np.where(codes == 3, np.nan, codes.astype(float))[keep.astype(bool)]
        assert_soft_close(got["auto"], O.map_mutations_gt(muts, S.astype(float), 0.01, 0.1))


@pytest.mark.parametrize("variant", [4, 3, 1])
@pytest.mark.parametrize("threads,rows_per_run", [(32, 1), (64, 3), (96, 7), (1024, 256)])
def test_interval_path_block_shapes_and_run_lengths(ctx, monkeypatch, threads, rows_per_run, variant):
    """One warp per CTA up to 32 warps (most of them without a clade row), runs of 1..256 SNV rows, rows
    dropped by the filter, extreme error rates; all kernel variants (4: every per-row quantity in registers,
    software-pipelined rows -- the default; 3: leaves outside the sweeps and counts cached in shared memory;
    1: every row through the sweeps -- what very large trees get).
    The masks also hold an empty row, a duplicate of a leaf row and a duplicate of an internal row."""
    n_snv, n_cells = 333, 300
    if variant == 4 and threads > 640:
        threads = 640                  # variant 4 runs up to 640 threads per CTA (register budget)
    codes, S = shuffled_problem(n_snv, n_cells, seed=threads)
    S = np.concatenate([S, S[5:6], S[n_cells + 3:n_cells + 4], S[0:1]], axis=0)
    monkeypatch.setenv("SCS_PLACE_ALGO", {1: "interval1", 3: "interval3", 4: "interval"}[variant])
    codes[7] = 3                       # nothing observed: uniform over the rows of S
    codes[8] = 1                       # everything mutated
    codes[9] = 0
    keep = (np.arange(n_snv) % 4 != 1).astype(np.uint8)
    monkeypatch.setenv("SCS_IV_THREADS", str(threads))
    monkeypatch.setenv("SCS_IV_ROWS_PER_RUN", str(rows_per_run))
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
    pl.set_clades_host(bits, words)
    pl.set_genotypes_host(packed, pitch, keep)
    info = pl.info()
    assert info["algo"] == 2 and info["interval_variant"] == variant
    assert info["interval_threads"] == threads
    muts = np.where(codes == 3, np.nan, codes.astype(float))[keep.astype(bool)]
    for FP, FN in [(0.01, 0.1), (1e-6, 1e-6), (0.2, 0.4)]:
        assert_soft_close(pl.run_host(FP, FN), O.map_mutations_gt(muts, S.astype(float), FP, FN))
    pl.close()


@pytest.mark.parametrize("n_snv,n_cells", [(700, 300), (300, 2500), (520, 5000)])
def test_interval_leaf_split_matches_the_fused_kernel(ctx, monkeypatch, n_snv, n_cells):
    """Variant 4 with the leaf accumulators in the placement kernel (SCS_IV4_LEAF_SPLIT=0) and in their own kernel (= 1, the
    default from 2,048 cells on): the same soft weights to summation order, runs that end inside the matrix, filtered rows."""
    codes, S = shuffled_problem(n_snv, n_cells, seed=7 * n_snv + n_cells)
    keep = np.ones(n_snv, dtype=np.uint8)
    keep[::7] = 0
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    got = {}
    for split in ("0", "1"):
        monkeypatch.setenv("SCS_IV4_LEAF_SPLIT", split)
        monkeypatch.setenv("SCS_PLACE_ALGO", "interval")
        pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
        pl.set_clades_host(bits, words)
        pl.set_genotypes_host(packed, pitch, keep)
        got[split] = pl.run_host(0.01, 0.1)
        assert pl.info()["algo"] == 2 and pl.info()["interval_variant"] == 4
        pl.close()
    np.testing.assert_allclose(got["0"], got["1"], rtol=1e-6, atol=1e-7)
    monkeypatch.setenv("SCS_PLACE_ALGO", "gemm")
    pg = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
    pg.set_clades_host(bits, words)
    pg.set_genotypes_host(packed, pitch, keep)
    assert_soft_close(got["1"], pg.run_host(0.01, 0.1))
    pg.close()


def test_interval_path_needs_tree_shaped_masks_and_shared_memory(ctx, monkeypatch):
    """Masks that are not a laminar family, replicate batches and trees whose working set exceeds
    shared memory go through the GEMM; so do, by default, trees of fewer than 100 cells (all fixed
    cost either way) unless SCS_PLACE_ALGO=interval asks for the interval path."""
    rng = np.random.default_rng(3)
    codes, S = make_problem(50, 40, seed=1)
    bits, words = engine.pack_clades(S)
    pl = engine.Placement(ctx, 50, 40, S.shape[0])
    pl.set_clades_host(bits, words)
    assert pl.info()["algo"] == 1                       # default: small tree -> GEMM
    pl.close()
    monkeypatch.setenv("SCS_PLACE_ALGO", "interval")
    codes, S = make_problem(50, 40, seed=1)
    S2 = S.copy()
    S2[3] = 0
    S2[3, [0, 39]] = 1
    S2[4] = 0
    S2[4, [39, 20]] = 1                 # overlaps row 3 without nesting
    packed, pitch = engine.pack_genotypes(codes)
    for mat, want_algo in ((S, 2), (S2, 1)):
        bits, words = engine.pack_clades(mat)
        pl = engine.Placement(ctx, 50, 40, mat.shape[0])
        pl.set_genotypes_host(packed, pitch)             # genotypes first, clades second: either order works
        pl.set_clades_host(bits, words)
        assert pl.info()["algo"] == want_algo
        muts = np.where(codes == 3, np.nan, codes.astype(float))
        assert_soft_close(pl.run_host(0.01, 0.1), O.map_mutations_gt(muts, mat.astype(float), 0.01, 0.1))
        # the same plan, other masks: tree-shaped -> general and back
        other = S2 if mat is S else S
        bits2, _ = engine.pack_clades(other)
        pl.set_clades_host(bits2, words)
        assert pl.info()["algo"] == (1 if want_algo == 2 else 2)
        assert_soft_close(pl.run_host(0.01, 0.1), O.map_mutations_gt(muts, other.astype(float), 0.01, 0.1))
        pl.close()
    bits, words = engine.pack_clades(S)
    pb = engine.Placement(ctx, [50, 50], 40, S.shape[0])
    pb.set_clades_host(np.concatenate([bits, bits], axis=0), words)
    assert pb.info()["algo"] == 1
    pb.close()
    n_cells = 20000                                       # prefix counts + accumulators > 227 KB of shared memory
    tree = synth.coalescent_tree(n_cells, rng)
    Sb = tree.clade_matrix()
    bits, words = engine.pack_clades(Sb)
    pl = engine.Placement(ctx, 16, n_cells, Sb.shape[0])
    pl.set_clades_host(bits, words)
    assert pl.info()["algo"] == 1
    pl.close()


def test_radix_mode_falls_back_beyond_four_chunks(ctx, monkeypatch):
    monkeypatch.setenv("SCS_PLACE_MODE", "radix")
    n_cells, n_snv, n_rows = 16000, 140, 64            # 125 K blocks = 5 chunks: exact u8 planes instead
    pl = engine.Placement(ctx, n_snv, n_cells, n_rows)
    assert pl.info()["mode"] == 0
    rng = np.random.default_rng(3)
    S = (rng.random((n_rows, n_cells)) < 0.7).astype(np.uint8)
    S[-1] = 0
    codes = rng.choice([0, 1, 3], size=(n_snv, n_cells), p=[0.6, 0.3, 0.1]).astype(np.uint8)
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    pl.set_genotypes_host(packed, pitch)
    pl.set_clades_host(bits, words)
    cnt = pl.debug_counts()
    C1 = ((codes == 1).astype(np.float32) @ S.T.astype(np.float32)).astype(np.int64)
    C0 = ((codes == 0).astype(np.float32) @ S.T.astype(np.float32)).astype(np.int64)
    assert np.array_equal(cnt[:, :, 0], C1) and np.array_equal(cnt[:, :, 1], C0)
    pl.close()


def test_batched_replicates_match_individual_plans(ctx, monkeypatch):
    """A batch of replicate datasets (own SNVs, own tree, own error rates) placed in one launch
    per pass gives, group by group, the numbers of the single-dataset plans and of the oracle."""
    n_cells = 37
    rng = np.random.default_rng(21)
    snvs = [300, 1, 129, 1000, 255, 256, 257]
    rates = [(0.01, 0.1), (0.02, 0.2), (1e-6, 1e-6), (0.05, 0.3), (0.01, 0.1), (0.001, 0.05), (0.03, 0.15)]
    codes_l, S_l, keep_l = [], [], []
    for m in snvs:
        tree = synth.coalescent_tree(n_cells, rng)
        S_l.append(tree.clade_matrix())
        c = synth.observe(tree, synth.throw_mutations(tree, m, rng), 0.1, 0.01, 0.2, rng)
        codes_l.append(c)
        k = (c == 1).any(axis=1).astype(np.uint8)
        k[:1] = 1
        keep_l.append(k)
    n_rows = S_l[0].shape[0]
    packed, pitch = engine.pack_genotypes(np.vstack(codes_l))
    bits, words = engine.pack_clades(np.vstack(S_l))
    pl = engine.Placement(ctx, snvs, n_cells, n_rows)
    assert pl.info()["groups"] == len(snvs)
    pl.set_genotypes_host(packed, pitch, np.concatenate(keep_l))
    pl.set_clades_host(bits, words)
    FP = np.array([r[0] for r in rates])
    FN = np.array([r[1] for r in rates])
    got = pl.run_host(FP, FN)
    assert got.shape == (len(snvs), n_rows)
    for g, m in enumerate(snvs):
        muts = np.where(codes_l[g] == 3, np.nan, codes_l[g].astype(float))[keep_l[g].astype(bool)]
        want = O.map_mutations_gt(muts, S_l[g].astype(float), FP[g], FN[g])
        assert_soft_close(got[g], want)
        pk, pt = engine.pack_genotypes(codes_l[g])
        bg, wg = engine.pack_clades(S_l[g])
        monkeypatch.setenv("SCS_PLACE_ALGO", "gemm")          # the same kernels on one dataset: same numbers
        single = engine.map_mutations(ctx, pk, pt, bg, wg, n_cells, FP[g], FN[g], keep_l[g])
        np.testing.assert_allclose(got[g], single, rtol=0, atol=1e-9 * max(1.0, np.abs(single).max()))
        monkeypatch.delenv("SCS_PLACE_ALGO")                  # the single dataset's default (interval path): same to tolerance
        assert_soft_close(engine.map_mutations(ctx, pk, pt, bg, wg, n_cells, FP[g], FN[g], keep_l[g]), want)
    pl.close()


def test_streaming_placement_equals_resident_plan(ctx):
    """Host-resident matrix walked in row chunks (H2D of chunk c+1 under the GEMMs of chunk c):
    same soft weights, kept-SNV count and missing rates as one device-resident plan."""
    import torch
    n_cells, n_snv = 150, 5000
    codes, S = make_problem(n_snv, n_cells, seed=77)
    codes = np.concatenate([codes, np.zeros((n_snv, 1), dtype=np.uint8)], axis=1)      # an outgroup column
    S = np.concatenate([S, np.zeros((S.shape[0], 1), dtype=np.uint8)], axis=1)
    active = np.ones(n_cells + 1, dtype=np.uint8)
    active[-1] = 0
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    dev = torch.device("cuda", ctx.device)
    d = torch.as_tensor(packed, device=dev)
    keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
    n_kept, miss = engine.gt_reduce(ctx, d, pitch, n_snv, n_cells + 1, active, True, keep)
    pl = engine.Placement(ctx, n_snv, n_cells + 1, S.shape[0])
    pl.set_clades_host(bits, words)
    pl.set_genotypes(d, pitch, keep)
    want = pl.run_host(0.01, 0.1)
    st = engine.StreamingPlacement(ctx, n_snv, n_cells + 1, S.shape[0], n_chunks=7)
    assert st.n_chunks >= 6
    st.set_clades_host(bits, words)
    y = torch.zeros(S.shape[0], dtype=torch.float64, device=dev)
    pinned = torch.as_tensor(packed).pin_memory()
    for _ in range(2):                                           # buffers are reused across calls
        got_kept, got_miss = st.run(pinned, pitch, active, True, 0.01, 0.1, y)
        torch.cuda.synchronize()
        assert got_kept == n_kept
        np.testing.assert_allclose(got_miss, miss, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-6, atol=1e-6)


def test_row_keep_and_extreme_rates(ctx):
    codes, S = make_problem(900, 120, seed=5)
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    keep = (codes == 1).any(axis=1).astype(np.uint8)
    keep[::7] = 0
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    pl = engine.Placement(ctx, 900, 120, S.shape[0])
    pl.set_clades_host(bits, words)
    pl.set_genotypes_host(packed, pitch, keep)
    for FP, FN in [(1e-6, 1e-6), (0.2, 0.4), (1e-3, 0.3)]:      # LAMBDA_MIN rates are the reference's default
        want = O.map_mutations_gt(muts[keep.astype(bool)], S.astype(float), FP, FN)
        assert_soft_close(pl.run_host(FP, FN), want)
    pl.close()


def test_all_missing_and_all_wildtype_rows(ctx):
    n_cells = 40
    rng = np.random.default_rng(9)
    tree = synth.coalescent_tree(n_cells, rng)
    S = tree.clade_matrix()
    codes = np.zeros((6, n_cells), dtype=np.uint8)
    codes[1] = 3                      # nothing observed: uniform over all rows of S
    codes[2] = 1                      # everything mutated
    codes[3, ::2] = 3
    codes[4, 5] = 1
    codes[5, :] = 2
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    want = O.map_mutations_gt(muts, S.astype(float), 0.01, 0.1)
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    got = engine.map_mutations(ctx, packed, pitch, bits, words, n_cells, 0.01, 0.1)
    assert_soft_close(got, want)


def test_placement_properties_at_scale(ctx):
    """Size-independent properties on a problem the oracle would not finish:
    every kept SNV distributes exactly one unit of mass, SNV order does not
    matter, and two halves add up to the whole."""
    import torch
    n_cells, n_snv = 1000, 40000
    rng = np.random.default_rng(2)
    tree = synth.coalescent_tree(n_cells, rng)
    S = tree.clade_matrix()
    bits, words = engine.pack_clades(S)
    dev = torch.device("cuda", ctx.device)
    lo = torch.as_tensor(tree.lo[:-1], device=dev)
    hi = torch.as_tensor(tree.hi[:-1], device=dev)
    rate = torch.as_tensor(tree.branch_len[:-1], device=dev, dtype=torch.float32)
    packed, pitch, keep = synth.observe_packed_torch(lo, hi, torch.as_tensor(tree.leaf_rank), n_snv, 0.1, 0.01, 0.2, dev,
                                                      seed=4, rate=rate)
    pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
    pl.set_clades_host(bits, words)
    pl.set_genotypes(packed, pitch, keep)
    y = pl.run_host(0.01, 0.1)
    n_kept = int(keep.sum())
    assert abs(y.sum() - n_kept) <= 1e-6 * n_kept
    assert (y >= 0).all()
    perm = torch.randperm(n_snv, device=dev)
    pl.set_genotypes(packed[perm].contiguous(), pitch, keep[perm].contiguous())
    y_perm = pl.run_host(0.01, 0.1)
    np.testing.assert_allclose(y_perm, y, rtol=1e-5, atol=1e-4)
    half = n_snv // 2
    pa = engine.Placement(ctx, half, n_cells, S.shape[0])
    pa.set_clades_host(bits, words)
    pa.set_genotypes(packed[:half].contiguous(), pitch, keep[:half].contiguous())
    ya = pa.run_host(0.01, 0.1)
    pa.set_genotypes(packed[half:].contiguous(), pitch, keep[half:].contiguous())
    yb = pa.run_host(0.01, 0.1)
    np.testing.assert_allclose(ya + yb, y, rtol=1e-5, atol=1e-4)
    # same inputs, same bits: the reduction order is fixed
    pl.set_genotypes(packed, pitch, keep)
    assert np.array_equal(pl.run_host(0.01, 0.1), y)


def test_placement_properties_at_full_baseline_size(ctx):
    """BASELINE configs[3] at full size (10,000 cells x 1,000,000 SNVs, ~16 GB of device memory): the
    oracle cannot run here, so check what must hold at any size -- one unit of mass per kept SNV,
    non-negative weights, bitwise run-to-run reproducibility, and additivity over SNV rows (the two
    halves, placed by a separate plan, add up to the whole)."""
    import torch
    if ctx.total_mem_mib < 60000:
        pytest.skip("needs a GPU with >= 60 GB")
    n_cells, n_snv = 10000, 1000000
    rng = np.random.default_rng(1234)
    tree = synth.coalescent_tree(n_cells, rng)
    S = tree.clade_matrix()
    bits, words = engine.pack_clades(S)
    dev = torch.device("cuda", ctx.device)
    lo = torch.as_tensor(tree.lo[:-1], device=dev)
    hi = torch.as_tensor(tree.hi[:-1], device=dev)
    rate = torch.as_tensor(tree.branch_len[:-1], device=dev, dtype=torch.float32)
    packed, pitch, keep = synth.observe_packed_torch(lo, hi, torch.as_tensor(tree.leaf_rank), n_snv, 0.1, 0.01, 0.15, dev,
                                                      seed=11, rate=rate)
    n_kept = int(keep.sum())
    pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
    assert pl.info()["mode"] == 1 and pl.info()["k_chunks"] == 3
    pl.set_clades_host(bits, words)
    pl.set_genotypes(packed, pitch, keep)
    y = pl.run_host(0.01, 0.05)
    assert (y >= 0).all()
    assert abs(y.sum() - n_kept) <= 2e-6 * n_kept
    assert np.array_equal(pl.run_host(0.01, 0.05), y)
    pl.close()
    half = n_snv // 2
    ph = engine.Placement(ctx, half, n_cells, S.shape[0])
    ph.set_clades_host(bits, words)
    ph.set_genotypes(packed[:half], pitch, keep[:half])
    ya = ph.run_host(0.01, 0.05)
    ph.set_genotypes(packed[half:], pitch, keep[half:])
    yb = ph.run_host(0.01, 0.05)
    ph.close()
    np.testing.assert_allclose(ya + yb, y, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("name", CASES)
def test_placement_on_golden_cases(ctx, name):
    g = load(name)
    vcf, tree_file, meta = case_inputs(g)
    tree, outg, FP0, FN0 = P.read_tree(tree_file, g["samples"])
    columns = list(g["samples"])
    active = np.array([c != outg for c in columns], dtype=np.uint8)
    bits, words, nodes = P._clade_bits(tree, columns, active)
    codes = g["muts_codes"]
    packed, pitch = engine.pack_genotypes(codes)
    keep = (((codes == 1) | (codes == 2)) & active.astype(bool)[None, :]).any(axis=1).astype(np.uint8)
    if outg not in columns:
        keep[:] = 1
    got = engine.map_mutations(ctx, packed, pitch, bits, words, len(columns), float(g["FP"]), float(g["FN"]), keep)
    assert got.shape == g["soft"].shape
    assert_soft_close(got, g["soft"])


# ------------------------------------------------------------------ decode
@pytest.mark.parametrize("name", CASES)
def test_vcf_decode_bit_exact(ctx, name):
    g = load(name)
    vcf, tree_file, meta = case_inputs(g)
    muts, _, _ = P.get_mut_df(vcf, meta["exclude"], meta["include"], ctx=ctx)
    assert list(muts.columns) == list(g["samples"])
    assert np.array_equal(muts.codes, g["muts_codes"])
    assert np.array_equal(np.array(muts.index).reshape(-1, 2), g["pos"])
    want = O.get_mut_df(vcf, meta["exclude"], meta["include"])
    assert muts._dec.n_noalt == want["skipped"][0] and muts._dec.n_filtered == want["skipped"][1]


def _synthetic_vcf_text(n_cells, n_lines, seed):
    """CellCoal-style records (GT:DP:RC:GQ:TG, multi-allelic ALT, mixed FILTERs) built in memory."""
    rng = np.random.default_rng(seed)
    tree = synth.coalescent_tree(n_cells, rng)
    codes = synth.observe(tree, synth.throw_mutations(tree, n_lines, rng), 0.1, 0.01, 0.15, rng)
    names = ["cell%04d" % (i + 1) for i in range(n_cells)] + ["outgcell"]
    out = ["##fileformat=VCFv4.3", "##source=synthetic",
           "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names)]
    bases = "ACGT"
    for i in range(n_lines):
        ref = bases[rng.integers(4)]
        alts = [b for b in bases if b != ref]
        a = 1 + int(rng.integers(3))
        a2 = 1 + (a % 3)
        recs = []
        for c in codes[i]:
            if c == 3:
                recs.append(".|.:4:4,0,0,0:8:0|0")
            elif c == 0:
                recs.append("0|0:12:12,0,0,0:10:0|0")
            else:
                alt = a2 if rng.random() < 0.03 else a                 # a few calls of a second alt allele
                gt = "%d|%d" % (alt, alt) if rng.random() < 0.1 else "0|%d" % alt
                recs.append(gt + ":9:4,0,5,0:33:0|%d" % a)
        recs.append("0|0:20:20,0,0,0:10:0|0")
        u = rng.random()
        flt = "PASS" if u > 0.15 else ("wildtype" if u < 0.05 else ("s50" if u < 0.1 else "singleton"))
        out.append("1\t%d\tsnv%06d\t%s\t%s\t.\t%s\tAA=%s;NS=%d\tGT:DP:RC:GQ:TG\t%s"
                   % (1 + 3 * i, i, ref, ",".join(alts), flt, ref, n_cells, "\t".join(recs)))
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("n_cells,n_lines", [(60, 1500), (333, 400), (2, 50)])
def test_vcf_decode_bit_exact_on_generated_files(ctx, tmp_path, n_cells, n_lines):
    p = tmp_path / "gen.vcf"
    p.write_text(_synthetic_vcf_text(n_cells, n_lines, seed=n_cells + n_lines))
    for excl, incl in (("", ""), ("tumcell000[1-3]", ""), ("", "(tumcell00[0-2].)|healthycell")):
        want = O.get_mut_df(str(p), excl, incl)
        if len(want["samples"]) == 0:
            continue
        muts, _, _ = P.get_mut_df(str(p), excl, incl, ctx=ctx)
        wc = np.where(np.isnan(want["muts"]), 3, np.nan_to_num(want["muts"])).astype(np.uint8)
        assert list(muts.columns) == list(want["samples"])
        assert np.array_equal(muts.codes, wc)
        assert muts.index == want["pos"]
        assert [muts._dec.n_noalt, muts._dec.n_filtered] == want["skipped"]


def test_empty_inputs_fail_cleanly(ctx, tmp_path):
    head = "##fileformat=VCFv4.3\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tcell1\toutgcell\n"
    p = tmp_path / "none_kept.vcf"
    p.write_text(head + "1\t5\t.\tA\tC\t.\tLowQual\t.\tGT\t0|1\t0|0\n")
    m, _, _ = P.get_mut_df(str(p), "", "", ctx=ctx)
    assert m.shape == (0, 2) and m.codes.shape == (0, 2)
    with pytest.raises(engine._cabi.ScsError):
        engine.Placement(ctx, 0, 2, 4)                       # nothing to place
    with pytest.raises(engine._cabi.ScsError):
        engine.Placement(ctx, 10, 70000, 4)                  # beyond the 16-bit count packing
    pl = engine.Placement(ctx, 3, 2, 4)
    with pytest.raises(engine._cabi.ScsError):
        pl.run_host(0.0, 0.1)                                # error rates must lie in (0, 1)
    pl.close()


def test_vcf_decode_rejects_malformed_input(ctx, tmp_path):
    good = ("##fileformat=VCFv4.3\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tcell1\tcell2\toutgcell\n"
            "1\t10\t.\tA\tC\t.\tPASS\t.\tGT:DP\t0|1:3\t0|0:4\t0|0:2\n")
    p = tmp_path / "ok.vcf"
    p.write_text(good)
    m, _, _ = P.get_mut_df(str(p), "", "", ctx=ctx)
    assert m.shape == (1, 3) and m.codes.tolist() == [[1, 0, 0]]
    for bad in (good.replace("0|0:4\t0|0:2", "0|0:4"),                 # a sample column is missing
                good.replace("0|1:3", "0|x:3"),                         # GT does not end in a digit
                good.replace("1\t10", "chrX\t10"),                      # int('X') fails in the reference
                good.replace("GT:DP", "DP:AD")):                        # no GT in FORMAT
        q = tmp_path / "bad.vcf"
        q.write_text(bad)
        with pytest.raises(engine._cabi.ScsError):
            P.get_mut_df(str(q), "", "", ctx=ctx)


@pytest.mark.parametrize("n_snv,n_cells", [(5000, 77), (1, 1), (333, 16), (4097, 17), (10000, 100), (2500, 250), (1500, 512),
                                            (1200, 513), (900, 1000)])
def test_gt_reduce_matches_numpy(ctx, n_snv, n_cells):
    """Row filter + per-cell missing rates (run_PT_test.py:350-358): the fused narrow kernel
    (<= 512 cells) and the wide pair of kernels."""
    import torch
    rng = np.random.default_rng(8 + n_cells)
    codes = rng.choice([0, 0, 0, 1, 2, 3], size=(n_snv, n_cells)).astype(np.uint8)
    codes[::9] = np.where(codes[::9] == 3, 3, 0)
    packed, pitch = engine.pack_genotypes(codes)
    active = np.ones(n_cells, dtype=np.uint8)
    if n_cells > 4:
        active[[3, n_cells - 1]] = 0
    dev = torch.device("cuda", ctx.device)
    d = torch.as_tensor(packed, device=dev)
    keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
    n_kept, miss = engine.gt_reduce(ctx, d, pitch, n_snv, n_cells, active, True, keep)
    want_keep = (((codes == 1) | (codes == 2)) & active.astype(bool)[None, :]).any(axis=1)
    assert np.array_equal(keep.cpu().numpy().astype(bool), want_keep) and n_kept == want_keep.sum()
    if want_keep.any():
        np.testing.assert_array_equal(miss, (codes[want_keep] == 3).mean(axis=0))
    n_all, miss_all = engine.gt_reduce(ctx, d, pitch, n_snv, n_cells, active, False, keep)
    assert n_all == n_snv
    np.testing.assert_array_equal(miss_all, (codes == 3).mean(axis=0))


@pytest.mark.parametrize("algo,n_snv,n_cells", [("interval", 3000, 77), ("interval", 1, 5), ("interval", 2500, 1000), ("interval3", 700, 300),
                                                 ("interval1", 600, 130), ("gemm", 2000, 250), ("gemm", 900, 600), ("interval", 300, 10000)])
def test_placement_prepare_matches_gt_reduce(ctx, monkeypatch, algo, n_snv, n_cells):
    """scs_placement_prepare (one fused pass on the interval path: row flags, missing counts, rows in depth-first
    cell order) against scs_gt_reduce + bind_genotypes: same kept count, same missing fractions, same soft
    weights bit for bit; and accumulated over row chunks (what the host-buffer path does)."""
    import torch
    monkeypatch.setenv("SCS_PLACE_ALGO", algo)
    codes, S = shuffled_problem(n_snv, n_cells, seed=n_snv + n_cells)
    codes = np.concatenate([codes, np.zeros((n_snv, 1), dtype=np.uint8)], axis=1)      # an outgroup column
    codes[::11, -1] = 1                                                                # calls in the outgroup alone do not keep a row
    codes[::7] = np.where(codes[::7] == 1, 0, codes[::7])                              # rows without any call: dropped
    S = np.concatenate([S, np.zeros((S.shape[0], 1), dtype=np.uint8)], axis=1)
    n_cols = n_cells + 1
    active = np.ones(n_cols, dtype=np.uint8)
    active[-1] = 0
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    dev = torch.device("cuda", ctx.device)
    d = torch.as_tensor(packed, device=dev)
    keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
    want_kept, want_miss = engine.gt_reduce(ctx, d, pitch, n_snv, n_cols, active, True, keep)
    pl = engine.Placement(ctx, n_snv, n_cols, S.shape[0])
    pl.set_clades_host(bits, words)
    pl.bind_genotypes(d, pitch, keep)
    want = pl.run_host(0.01, 0.1)
    pl.close()
    for filt in (True, False):
        pl = engine.Placement(ctx, n_snv, n_cols, S.shape[0])
        pl.set_clades_host(bits, words)
        pl.prepare(d, pitch, active, filt)
        got = pl.run_host(0.01, 0.1)
        n_kept, miss = pl.prepare_result()
        if filt:
            assert n_kept == want_kept
            np.testing.assert_array_equal(miss, want_miss)
            assert np.array_equal(got, want)
        else:
            assert n_kept == n_snv
            np.testing.assert_array_equal(miss, (codes == 3).mean(axis=0))
        pl.close()
    if n_snv >= 600:
        # three row chunks through one chunk-sized plan, counters accumulated on the device
        rows = (n_snv + 2) // 3
        pl = engine.Placement(ctx, rows, n_cols, S.shape[0])
        pl.set_clades_host(bits, words)
        y = np.zeros(S.shape[0])
        for c in range(3):
            chunk = d[c * rows:(c + 1) * rows]
            if chunk.shape[0] < rows:                    # short last chunk: pad with all-missing rows (no call: dropped)
                pad = torch.full((rows - chunk.shape[0], pitch), 0xFF, dtype=torch.uint8, device=dev)
                chunk = torch.cat([chunk, pad], dim=0)
            chunk = chunk.contiguous()
            pl.prepare(chunk, pitch, active, True, accumulate=c > 0)
            y += pl.run_host(0.01, 0.1)
        n_kept, miss = pl.prepare_result()
        assert n_kept == want_kept
        np.testing.assert_allclose(miss, want_miss, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(y, want, rtol=1e-6, atol=1e-6)
        pl.close()


def test_gt_reduce_many_rows_per_lane(ctx):
    """Enough rows that every lane of the fused narrow kernel flushes its byte-wide counters
    (> 255 rows per lane): random packed bytes on the device, checked with torch integer ops."""
    import torch
    dev = torch.device("cuda", ctx.device)
    n_snv, n_cells, pitch = 45_000_000, 16, 16
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    d = torch.zeros((n_snv, pitch), dtype=torch.uint8, device=dev)
    d[:, :4] = torch.randint(0, 256, (n_snv, 4), dtype=torch.uint8, device=dev, generator=g)
    d[::3, :4] &= 0xF0           # rows with fewer mutated cells, some with none at all
    d[::7, :4] = 0xCC            # only missing / wild type: dropped by the filter
    active = np.ones(n_cells, dtype=np.uint8)
    active[5] = 0
    keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
    n_kept, miss = engine.gt_reduce(ctx, d, pitch, n_snv, n_cells, active, True, keep)
    want_keep = torch.zeros(n_snv, dtype=torch.bool, device=dev)
    is3 = []
    for c in range(n_cells):
        code = (d[:, c // 4] >> (2 * (c % 4))) & 3
        if active[c]:
            want_keep |= (code == 1) | (code == 2)
        is3.append(code == 3)
    assert torch.equal(keep.bool(), want_keep) and n_kept == int(want_keep.sum())
    want_miss = np.array([int((m & want_keep).sum()) for m in is3]) / n_kept
    np.testing.assert_array_equal(miss, want_miss)


# ------------------------------------------------------ weights and the LRT
@pytest.mark.parametrize("name", CASES)
def test_branch_weights_on_golden_cases(ctx, name):
    g = load(name)
    vcf, tree_file, meta = case_inputs(g)
    tree, outg, FP0, FN0 = P.read_tree(tree_file, g["samples"])
    MS = dict(zip(g["MS_leaf_names"], g["MS_leaf"]))
    for k, w in enumerate(g["w_maxs"]):
        P.add_br_weights(tree, float(g["FP"]), float(g["FN"]), MS, float(w))
        raw = np.array([n.weights[0] for n in tree.iter_descendants()])
        np.testing.assert_allclose(raw, g["raw_weights"][k], rtol=1e-10)


@pytest.mark.parametrize("name", CASES)
def test_lrt_on_golden_model_data(ctx, name):
    g = load(name)
    problems = [(g["Y"][k], g["constr"][k], g["weights_norm"][k]) for k in range(len(g["w_maxs"]))]
    res = P.get_LRT_poisson_batch(problems, ctx=ctx)
    for k, (LR, dof, on_bound, p_val, opt) in enumerate(res):
        assert dof == int(g["dof"][k])
        assert on_bound == int(g["on_bound"][k])
        assert abs(LR - g["LR"][k]) <= 1e-4 * max(1.0, abs(g["LR"][k]))
        assert abs(p_val - g["p_val"][k]) <= 2e-3 * g["p_val"][k] + 1e-300      # far tails underflow to 0 in both
        assert int(p_val < 0.05) == int(g["hyp"][k])                     # identical clock / no-clock decision
        x = np.array([v[0] for v in opt])
        assert np.abs(g["constr"][k] @ x).max() < 1e-8                    # the fit is exactly ultrametric
        assert (x > 1e-6).all()


def test_lrt_is_at_least_as_good_as_the_reference_optimum(ctx):
    """The fitted lengths must not have a worse (barrier-free) null likelihood than
    the reference's SciPy solution by more than the interior-point slack."""
    g = load("example_clock")
    for k in (0, len(g["w_maxs"]) - 1):
        Y, w = g["Y"][k], g["weights_norm"][k]
        (LR, dof, ob, p, opt), = P.get_LRT_poisson_batch([(Y, g["constr"][k], w)], ctx=ctx)
        assert LR <= g["LR"][k] + 1e-3


def test_lrt_batch_of_replicates(ctx):
    """1,000 independent replicate problems in one launch equal 1,000 single launches."""
    g = load("syn_a")
    base = (g["Y"][0], g["constr"][0], g["weights_norm"][0])
    rng = np.random.default_rng(0)
    probs = []
    for i in range(1000):
        Y = base[0] * rng.uniform(0.5, 1.5, size=base[0].size)
        probs.append((Y, base[1], base[2]))
    res = P.get_LRT_poisson_batch(probs, ctx=ctx)
    for i in (0, 17, 999):
        single, = P.get_LRT_poisson_batch([probs[i]], ctx=ctx)
        assert single[:4] == res[i][:4]
    # spot-check against SciPy on a few replicates
    for i in (3, 500):
        LR, dof, ob, p, x = O.get_LRT_poisson(probs[i][0], g["constr"][0], g["init"][0], probs[i][2])
        assert abs(res[i][0] - LR) <= 1e-4 * max(1.0, LR) and res[i][1] == dof and res[i][2] == ob


def test_lrt_degenerate_input_reports_failure(ctx):
    g = load("syn_c")
    (res,) = P.get_LRT_poisson_batch([(np.zeros_like(g["Y"][0]), g["constr"][0], g["weights_norm"][0])], ctx=ctx)
    assert len(res) == 6 and all(np.isnan(v) for v in res)            # the reference's failure tuple (:678)


# ------------------------------------------------------------- whole path
@pytest.mark.parametrize("name", CASES)
def test_run_poisson_tree_test_tsv(ctx, name, tmp_path):
    g = load(name)
    vcf, tree_file, meta = case_inputs(g)
    out = tmp_path / "out.tsv"
    P.run_poisson_tree_test(vcf, tree_file, str(out), [float(w) for w in g["w_maxs"]], meta["exclude"], meta["include"],
                            meta["FN_in"], meta["FP_in"])
    got = out.read_text().split("\n")
    ref = str(g["tsv"]).split("\n")
    assert got[0] == ref[0]
    rc, gc = ref[1].split("\t"), got[1].split("\t")
    assert rc[:3] == gc[:3]                                              # muts, FN, FP
    for i in range(3, len(rc), 4):
        assert abs(float(rc[i]) - float(gc[i])) <= 1e-4 * max(1.0, abs(float(rc[i]))) + 6e-4   # 3 printed decimals
        assert rc[i + 1] == gc[i + 1]                                    # dof
        assert abs(float(rc[i + 2]) - float(gc[i + 2])) <= 2e-3 * float(rc[i + 2]) + 1e-300
        assert rc[i + 3] == gc[i + 3]                                    # H0 / H1


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_whole_path_on_random_datasets_against_the_oracle(ctx, tmp_path, seed):
    """VCF text + Newick file -> TSV through the product path (GPU decode, row filter, placement, branch weights, LRT) against
    the oracle's whole path (NumPy + SciPy trust-constr) on random small datasets: a random coalescent tree of 6 .. 40 cells,
    100 .. 400 SNVs with missing calls, a few records the decoder has to skip, error rates taken from the file path
    (:287-294).  Same tolerances as the golden TSVs; a w_max where SciPy itself did not reach the optimum (LR above ours by
    more than the tolerance: the cases test_lrt_sweep_against_the_reference classifies) is allowed to differ in LR / p, never
    in muts, FN, FP or dof."""
    import gzip
    rng = np.random.default_rng(seed)
    n_cells = int(rng.integers(6, 41))
    n_snv = int(rng.integers(100, 401))
    names = ["cell%04d" % (i + 1) for i in range(n_cells)] + ["outgcell"]
    tree = synth.coalescent_tree(n_cells, rng)
    codes = synth.observe(tree, synth.throw_mutations(tree, n_snv, rng), 0.15, 0.01, float(rng.uniform(0.05, 0.3)), rng)
    d = tmp_path / ("sim_WGA0.%d-0-0.0%d" % (1 + seed % 3, 1 + seed % 2))
    d.mkdir()
    vcf = d / "rep.vcf.gz"
    with gzip.open(vcf, "wt") as f:
        f.write("##fileformat=VCFv4.3\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        for i in range(n_snv):
            gts = [".|." if c == 3 else ("0|0" if c == 0 else ("0|1" if c == 1 else "1|1")) for c in codes[i]] + ["0|0"]
            filt = "PASS" if rng.random() > 0.05 else "LowQual"                       # (skipped by both)
            f.write("1\t%d\t.\tA\tC\t.\t%s\t.\tGT\t%s\n" % (10 + i, filt, "\t".join(gts)))
    tf = d / "rep.newick"
    tf.write_text(tree.newick(names[:-1], names[-1], scale=0.05) + "\n")
    w = [100.0, 500.0, 1000.0]
    out = tmp_path / "out.tsv"
    P.run_poisson_tree_test(str(vcf), str(tf), str(out), w, "", "", None, None)
    want_tsv, want = O.run_poisson_tree_test(str(vcf), str(tf), w)
    got = out.read_text().split("\n")
    ref = want_tsv.split("\n")
    assert got[0] == ref[0]
    rc, gc = ref[1].split("\t"), got[1].split("\t")
    assert rc[:3] == gc[:3]                                              # muts, FN, FP
    for k, i in enumerate(range(3, len(rc), 4)):
        assert rc[i + 1] == gc[i + 1]                                    # dof
        lr_ref, lr_got = float(rc[i]), float(gc[i])
        if lr_ref - lr_got > 1e-4 * max(1.0, abs(lr_ref)) + 6e-4:
            continue                                                     # SciPy stopped above the optimum (our null likelihood is the higher one)
        assert abs(lr_ref - lr_got) <= 1e-4 * max(1.0, abs(lr_ref)) + 6e-4
        assert abs(float(rc[i + 2]) - float(gc[i + 2])) <= 2e-3 * float(rc[i + 2]) + 1e-300
        assert rc[i + 3] == gc[i + 3]                                    # H0 / H1


def test_run_poisson_tree_test_batch(ctx, tmp_path):
    """A sweep of (VCF, tree) pairs through the batched path: the two example-sized pairs share a
    shape and go through ONE placement launch per pass; every TSV matches the reference's."""
    names = ["syn_d_clock", "syn_a", "syn_d_noclock", "syn_e_mismatch",
         "syn_f_scite_log", "syn_g_scite_csv", "syn_h_path_rates"]
    gs = [load(n) for n in names]
    vcfs, trees = zip(*[case_inputs(g)[:2] for g in gs])
    outs = [str(tmp_path / (n + ".tsv")) for n in names]
    w = [100.0, 500.0, 1000.0]
    res = P.run_poisson_tree_test_batch(list(vcfs), list(trees), outs, w, ctx=ctx)
    assert [r[0] for r in res] == list(range(len(names)))
    for g, out in zip(gs, outs):
        got = open(out).read().split("\n")[1].split("\t")
        ref = str(g["tsv"]).split("\n")[1].split("\t")
        assert got[:3] == ref[:3]
        gw = list(g["w_maxs"])
        for j, wm in enumerate(w):
            if wm not in gw:
                continue
            k = gw.index(wm)
            LR, dof, p, hyp = float(got[3 + 4 * j]), got[4 + 4 * j], float(got[5 + 4 * j]), got[6 + 4 * j]
            assert abs(LR - g["LR"][k]) <= 1e-4 * max(1.0, abs(g["LR"][k])) + 6e-4
            assert dof == str(int(g["dof"][k])) and hyp == "H%d" % int(g["hyp"][k])
            assert abs(p - g["p_val"][k]) <= 2e-3 * g["p_val"][k] + 1e-300


def test_replicate_sweep_equals_the_per_pair_path(ctx, tmp_path, monkeypatch):
    """engine.ReplicateSweep (native tree parsing on host threads, per-group row filters, batched placement, branch
    weights + LRT branch order on the device, one LRT launch, native TSV formatting) against the per-pair path
    (read_tree / _clade_bits / get_model_data-style bookkeeping in Python per tree) on 24 distinct random trees with
    their own genotypes and error rates: identical muts / FN / FP / dof / decisions, LR and p to 1e-5."""
    import gzip
    rng = np.random.default_rng(77)
    n_cells, n_rep = 14, 24
    names = ["cell%04d" % (i + 1) for i in range(n_cells)] + ["outgcell"]
    vcfs, trees = [], []
    for r in range(n_rep):
        tree = synth.coalescent_tree(n_cells, rng)
        n_snv = int(rng.integers(60, 140))
        codes = synth.observe(tree, synth.throw_mutations(tree, n_snv, rng), 0.15, 0.01, 0.1, rng)
        d = tmp_path / ("sim_WGA0.%d-0-0.0%d" % (1 + r % 3, 1 + r % 2))
        d.mkdir(exist_ok=True)
        vcf = d / ("rep%02d.vcf.gz" % r)
        with gzip.open(vcf, "wt") as f:
            f.write("##fileformat=VCFv4.3\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
            for i in range(n_snv):
                gts = [".|." if c == 3 else ("0|0" if c == 0 else "0|1") for c in codes[i]] + ["0|0"]
                f.write("1\t%d\t.\tA\tC\t.\tPASS\t.\tGT\t%s\n" % (10 + i, "\t".join(gts)))
        tf = d / ("rep%02d.newick" % r)
        tf.write_text(tree.newick(names[:-1], names[-1], scale=0.05) + "\n")
        vcfs.append(str(vcf))
        trees.append(str(tf))
    w = [100.0, 400.0, 1000.0]
    outs_a = [str(tmp_path / ("a%02d.tsv" % r)) for r in range(n_rep)]
    outs_b = [str(tmp_path / ("b%02d.tsv" % r)) for r in range(n_rep)]
    res_a = P.run_poisson_tree_test_batch(vcfs, trees, outs_a, w, ctx=ctx)
    monkeypatch.setenv("SCS_SWEEP_PATH", "pairs")
    res_b = P.run_poisson_tree_test_batch(vcfs, trees, outs_b, w, ctx=ctx)
    assert [r[0] for r in res_a] == list(range(n_rep)) == [r[0] for r in res_b]
    n_fail = 0
    for r in range(n_rep):
        (ha, fa), (hb, fb) = _tsv_fields(outs_a[r]), _tsv_fields(outs_b[r])
        assert ha == hb and fa[:3] == fb[:3], (r, fa[:3], fb[:3])
        for k, (ra, rb) in enumerate(zip(res_a[r][1], res_b[r][1])):
            if np.isnan(rb[0]):
                assert np.isnan(ra[0])
                n_fail += 1
                continue
            assert ra[1] == rb[1] and ra[2] == rb[2], (r, k, ra, rb)
            # (the batch takes the small-tree interval path, the per-pair path the GEMM: the same exact counts, fp32 terms summed in another order)
            assert abs(ra[0] - rb[0]) <= 1e-5 * max(1.0, abs(rb[0])) and abs(ra[3] - rb[3]) <= 1e-5 * rb[3] + 1e-300, (r, k, ra, rb)
            assert fa[6 + 4 * k] == fb[6 + 4 * k]
    assert n_fail <= n_rep


def test_cli_entry_point(ctx, tmp_path):
    g = load("syn_a")
    vcf, tree_file, meta = case_inputs(g)
    out = tmp_path / "cli.tsv"
    P.main([vcf, tree_file, "-o", str(out), "-w", "100", "1000"])
    lines = out.read_text().split("\n")
    assert lines[0].split("\t")[:4] == ["muts", "FN", "FP", "-2logLR_100"] and len(lines) == 2


def test_ten_thousand_cells_against_the_interval_oracle(ctx, monkeypatch):
    """BASELINE's cell count against a float64 oracle: the dense oracle cannot do 10,000 cells x 20,000
    clade rows in seconds, its prefix-count form (oracle.pt_oracle.map_mutations_gt_interval, checked
    against the dense form and the reference's fixtures on the CPU) can.  Both device algorithms."""
    n_snv, n_cells = 1500, 10000
    codes, S = shuffled_problem(n_snv, n_cells, seed=4242)
    keep = (((codes == 1) | (codes == 2)).any(axis=1)).astype(np.uint8)
    muts = np.where(codes == 3, np.nan, codes.astype(float))[keep.astype(bool)]
    want = O.map_mutations_gt_interval(muts, S.astype(float), 0.01, 0.05)
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    for algo, want_algo in (("auto", 2), ("gemm", 1)):
        monkeypatch.setenv("SCS_PLACE_ALGO", algo)
        pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
        pl.set_clades_host(bits, words)
        pl.set_genotypes_host(packed, pitch, keep)
        got = pl.run_host(0.01, 0.05)
        assert pl.info()["algo"] == want_algo
        pl.close()
        assert_soft_close(got, want)


# ------------------------------------------------- LRT: differential sweep against the reference's SciPy solve
def _ultrametric_gap(x, child_int):
    """max |height via first child - height via second child| over the internal nodes (0 = the clock constraints hold)."""
    m1 = len(x) // 2
    h = np.zeros(m1)
    gap = 0.0
    for i in range(m1):                 # children have smaller indices (post-order)
        a = x[2 * i] + (h[child_int[2 * i]] if child_int[2 * i] >= 0 else 0.0)
        b = x[2 * i + 1] + (h[child_int[2 * i + 1]] if child_int[2 * i + 1] >= 0 else 0.0)
        gap = max(gap, abs(a - b))
        h[i] = 0.5 * (a + b)
    return gap


def test_lrt_sweep_against_the_reference(ctx, monkeypatch):
    """240 random clock tests (3..60 leaves; clock-like, rate-shifted, sparse and very sparse counts with zero-count
    branches and up to 40 fitted lengths on the bound; non-integer soft counts; weights clipped at different w_max)
    solved by the reference's own get_LRT_poisson (tests/golden/make_golden.py::make_lrt_sweep).

    Every problem must give the reference's dof and the reference's decision at alpha = 0.05.  -2logLR (1e-4), on_bound
    (exact) and p (2e-3) must match too -- except where the REFERENCE'S number is provably not the optimum it is after,
    which SciPy's trust-constr produces in three ways, each checked here instead of waved through:
      * early stop: SciPy terminates as soon as barrier parameter < barrier_tol (1e-5: from its stage 6 on, mu = 0.1 * 0.2^k)
        and the trust radius < xtol -- then the reference sits on the central path at mu_k, k < 10: the kernel, run at that
        mu_k (SCS_LRT_MU), must reproduce its LR to 1e-4 and its on_bound exactly, and the default answer must be the better
        optimum (smaller LR);
      * no convergence: its LR is far above the constrained optimum -- the kernel's fitted lengths must be feasible
        (ultrametric to 1e-9, inside the bounds) and have the HIGHER null likelihood when re-evaluated here in float64;
      * a fitted length within 1.5x of the on_bound threshold 1e-3 (:687) falling on the other side of it.
    Problems with ll_H1 < 0 -- where the reference's negative scale_fac makes it minimise the wrong sign of its objective
    -- must come back with status 4 and the failure tuple, not with a number."""
    g = load("lrt_sweep")
    off = g["off"]
    n = len(off) - 1
    Ys = [g["Y"][off[i]:off[i + 1]] for i in range(n)]
    ws = [g["w"][off[i]:off[i + 1]] for i in range(n)]
    chs = [g["child_int"][off[i]:off[i + 1]] for i in range(n)]
    out, xs = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    ladder = {}
    for k in (6, 7, 8, 9):
        monkeypatch.setenv("SCS_LRT_MU", repr(0.1 * 0.2 ** k))
        ladder[k] = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=False)[0]
    monkeypatch.delenv("SCS_LRT_MU")
    assert n >= 200

    def close(o, i, with_p=True):
        return (o[7] == 0 and int(o[2]) == int(g["on_bound"][i]) and abs(o[0] - g["LR"][i]) <= 1e-4 * max(1.0, abs(g["LR"][i]))
                and (not with_p or abs(o[3] - g["p_val"][i]) <= 2e-3 * g["p_val"][i] + 1e-300))

    n_neg = n_exact = n_early = n_noconv = n_flip = n_bound5 = n_zero = 0
    bad = []
    for i in range(n):
        LR, dof, ob, p, h0, h1, it, status = out[i]
        if g["ll_H1"][i] < 0:
            n_neg += 1
            assert status == 4 and np.isnan(LR), (i, status, LR)
            continue
        assert not np.isnan(g["LR"][i])
        n_bound5 += int(g["on_bound"][i] >= 5)
        n_zero += int((Ys[i] == 0).any())
        assert status == 0 and int(dof) == int(g["dof"][i]), (i, status, dof)
        assert int(p < 0.05) == int(g["p_val"][i] < 0.05), (i, p, float(g["p_val"][i]))       # identical clock / no-clock decision
        x = xs[i]
        U = Ys[i].sum()
        assert (x > 1e-6).all() and (x < U).all() and _ultrametric_gap(x, chs[i]) <= 1e-9 * max(1.0, U)
        if close(out[i], i):
            n_exact += 1
            continue
        ref_llH0 = g["ll_H1"][i] - 0.5 * g["LR"][i]
        my_llH0 = float(np.sum((Ys[i] * np.log(x) - x) * ws[i]))
        assert abs(my_llH0 - h0) <= 1e-9 * max(1.0, abs(h0))
        if LR <= g["LR"][i] + 1e-9 and any(close(ladder[k][i], i, with_p=False) for k in ladder):
            n_early += 1                    # the reference stopped at an earlier point of the same central path
        elif g["LR"][i] - LR > 1e-4 * max(1.0, abs(g["LR"][i])) and my_llH0 > ref_llH0:
            n_noconv += 1                   # the reference's point is not the constrained optimum: ours is feasible and better
        elif (abs(LR - g["LR"][i]) <= 1e-4 * max(1.0, abs(g["LR"][i]))
              and abs(int(ob) - int(g["on_bound"][i])) <= int(((x > 1e-3 / 1.5) & (x < 1e-3 * 1.5)).sum())):
            n_flip += 1
        else:
            bad.append((i, str(g["kind"][i]), int(g["n_leaves"][i]), float(LR), float(g["LR"][i]), int(ob), int(g["on_bound"][i]), float(p), float(g["p_val"][i])))
    assert not bad, bad[:10]
    counts = dict(exact=n_exact, early_stop=n_early, reference_not_converged=n_noconv, threshold_flip=n_flip, negative_ll_H1=n_neg)
    print("lrt sweep:", counts)
    assert n_exact >= 190 and n_early <= 12 and n_noconv <= 4 and n_flip <= 2, counts
    assert n_bound5 >= 20 and n_zero >= 40 and n_neg >= 5, (n_bound5, n_zero, n_neg)


@pytest.mark.parametrize("n_cells,R", [(3, 5), (14, 40), (100, 12), (130, 6), (300, 4), (511, 3)])
def test_small_tree_interval_path_matches_the_gemm_and_the_oracle(ctx, monkeypatch, n_cells, R):
    """scs_placement_set_trees: batches of small trees through the interval path (scs_place_iv5.cuh, 8 / 16 / 32 lanes per
    SNV) against the same batch through the tcgen05 GEMM (SCS_PLACE_ALGO=gemm) and, per pair, against the float64 oracle:
    ragged group sizes (a task boundary inside a group, a group smaller than one trip), filtered rows, per-pair error rates,
    an outgroup column that is in no tree."""
    import torch
    from oracle import c_oracle
    rng = np.random.default_rng(1000 + n_cells)
    names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
    columns = names + ["healthycell"]
    snvs = [int(v) for v in rng.integers(1, 60, size=R)]
    snvs[0] = 2500                               # more than one task (2048 rows)
    trees = [synth.coalescent_tree(n_cells, rng) for _ in range(R)]
    newicks = [t.newick(names, "healthycell", scale=0.05) for t in trees]
    mats, codes_all = [], []
    for t, m in zip(trees, snvs):
        codes = synth.observe(t, synth.throw_mutations(t, m, rng), 0.1, 0.02, 0.2, rng)
        codes[rng.random(m) < 0.1] = 0           # rows without any mutated cell: dropped by the row filter
        codes = np.concatenate([codes, np.zeros((m, 1), dtype=np.uint8)], axis=1)
        codes_all.append(codes)
        packed, pitch = engine.pack_genotypes(codes)
        mats.append(packed)
    d_gt = torch.as_tensor(np.concatenate(mats, axis=0)).cuda()
    FP = rng.uniform(0.005, 0.03, size=R)
    FN = rng.uniform(0.03, 0.12, size=R)
    w = [100.0, 1000.0]
    got = {}
    for algo in ("auto", "gemm"):
        monkeypatch.setenv("SCS_PLACE_ALGO", algo)
        sw = engine.ReplicateSweep(ctx, snvs, columns, w)
        sw.parse_trees(newicks)
        sw.enqueue(d_gt, FP, FN)
        out, n_kept, status = sw.results()
        got[algo] = (sw.d_y.cpu().numpy().reshape(R, -1).copy(), out.copy(), n_kept.copy(), status.copy(), sw.place.info())
        sw.close()
    assert got["auto"][4]["algo"] == 3 and got["auto"][4]["interval_variant"] == 5 and got["gemm"][4]["algo"] == 1
    y_iv, y_g = got["auto"][0], got["gemm"][0]
    assert np.array_equal(got["auto"][2], got["gemm"][2]) and np.array_equal(got["auto"][3], got["gemm"][3])
    assert np.abs(y_iv - y_g).max() <= 5e-6 * max(1.0, np.abs(y_g).max())
    ok = got["gemm"][1][:, :, 7] == 0
    assert np.array_equal(got["auto"][1][:, :, 7] == 0, ok)
    np.testing.assert_allclose(got["auto"][1][:, :, 0][ok], got["gemm"][1][:, :, 0][ok], rtol=1e-4, atol=1e-4)
    # the oracle, pair by pair: S from the product's own tree code (row order of iter_descendants), float64 placement
    for r in (0, 1, R - 1):
        bits, words, children, info = engine.tree_parse_batch([newicks[r]], columns, 2 * n_cells)
        S = engine.unpack_clades(bits[0], len(columns))
        keep = (((codes_all[r] == 1) | (codes_all[r] == 2))[:, :n_cells]).any(axis=1)
        want = c_oracle.place(codes_all[r][keep], S.astype(np.uint8), float(FP[r]), float(FN[r]))
        assert_soft_close(y_iv[r], want)


def test_replicate_sweep_stream_equals_one_batch(ctx):
    """engine.ReplicateSweepStream (host matrices + Newick text, sub-batches pipelined over a copy stream, a parser thread
    and the compute stream; TSV lines formatted on host threads) gives the lines of ONE ReplicateSweep over the same pairs,
    character for character, on repeated runs (buffers reused) and with per-pair error rates."""
    import torch
    rng = np.random.default_rng(5)
    n_cells, R = 30, 150
    names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
    columns = names + ["healthycell"]
    snvs = [int(v) for v in rng.integers(150, 400, size=R)]
    trees = [synth.coalescent_tree(n_cells, rng) for _ in range(R)]
    newicks = [t.newick(names, "healthycell", scale=0.05) for t in trees]
    mats = []
    for t, m in zip(trees, snvs):
        codes = synth.observe(t, synth.throw_mutations(t, m, rng), 0.1, 0.01, 0.15, rng)
        codes = np.concatenate([codes, np.zeros((m, 1), dtype=np.uint8)], axis=1)
        packed, pitch = engine.pack_genotypes(codes)
        mats.append(packed)
    host = torch.as_tensor(np.concatenate(mats, axis=0)).pin_memory()
    FP = rng.uniform(0.005, 0.02, size=R)
    FN = rng.uniform(0.03, 0.08, size=R)
    w = [100.0, 500.0, 1000.0]
    one = engine.ReplicateSweep(ctx, snvs, columns, w)
    one.parse_trees(newicks)
    one.enqueue(host.cuda(), FP, FN)
    want = one.rows()
    want_status = one.results()[2].copy()
    one.close()
    assert len(want) == R and sum(r.endswith("H0") or r.endswith("H1") for r in want) == R
    for n_sub in (1, 3, 7):
        stream = engine.ReplicateSweepStream(ctx, snvs, columns, w, n_sub=n_sub)
        for _ in range(2):
            rows, status = stream.run(host, newicks, FP, FN)
            assert rows == want
            assert np.array_equal(status, want_status)
        stream.close()


# ------------------------------------------------- the cluster-per-problem LRT kernel (scs_lrt_tree.cuh)
def _synthetic_lrt_problem(n_leaves, seed, sparse=False):
    """Counts, weights and get_model_data's child_int of a random coalescent tree (branch 2 i + c hangs below internal node i)."""
    rng = np.random.default_rng(seed)
    t = synth.coalescent_tree(n_leaves, rng)
    left, right = np.asarray(t.left), np.asarray(t.right)
    m1 = len(left)
    child = np.empty(2 * m1, dtype=np.int32)
    child[0::2] = np.where(left >= n_leaves, left - n_leaves, -1)
    child[1::2] = np.where(right >= n_leaves, right - n_leaves, -1)
    ids = np.empty(2 * m1, dtype=np.int64)
    ids[0::2], ids[1::2] = left, right
    length = np.asarray(t.branch_len, dtype=np.float64)[ids]
    Y = rng.poisson(length / length.mean() * (0.8 if sparse else 25.0)).astype(np.float64)
    Y += rng.uniform(0.0, 0.3, size=Y.size) * (Y > 0)           # soft counts are not integers
    w = rng.uniform(0.4, 1.6, size=Y.size)
    w *= Y.size / w.sum()
    return Y, w, child


def _assert_same_lrt(a, b, xa=None, xb=None):
    assert np.array_equal(a[:, 7], b[:, 7]), (a[:, 7], b[:, 7])                 # status
    ok = a[:, 7] == 0
    assert np.array_equal(a[ok, 1], b[ok, 1]) and np.array_equal(a[ok, 2], b[ok, 2])          # dof, on_bound
    np.testing.assert_allclose(a[ok, 0], b[ok, 0], rtol=1e-8, atol=1e-8)        # -2 log LR
    np.testing.assert_allclose(a[ok, 3], b[ok, 3], rtol=1e-6, atol=1e-300)      # p
    np.testing.assert_allclose(a[ok, 4], b[ok, 4], rtol=1e-10)                  # ll_H0
    np.testing.assert_allclose(a[ok, 5], b[ok, 5], rtol=1e-12)                  # ll_H1
    if xa is not None:
        for i in np.flatnonzero(ok):
            np.testing.assert_allclose(xa[i], xb[i], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("cluster", ["1", "2", "8"])
@pytest.mark.parametrize("top_cap", [None, "9", "0"])
def test_lrt_tree_kernel_on_the_reference_sweep(ctx, monkeypatch, cluster, top_cap):
    """The cluster kernel forced onto the 240 small problems of the reference sweep (where the default is one warp per
    problem): same statuses, dof, on_bound; LR, p and fitted lengths equal to rounding.  top_cap = 9 / 0 pushes all but
    the last few levels (every level) below the cut, i.e. through the cluster-wide sweeps in global memory."""
    g = load("lrt_sweep")
    off = g["off"]
    n = len(off) - 1
    Ys = [g["Y"][off[i]:off[i + 1]] for i in range(n)]
    ws = [g["w"][off[i]:off[i + 1]] for i in range(n)]
    chs = [g["child_int"][off[i]:off[i + 1]] for i in range(n)]
    base, xb = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    monkeypatch.setenv("SCS_LRT_KERNEL", "tree")
    monkeypatch.setenv("SCS_LRT_CLUSTER", cluster)
    if top_cap is not None:
        monkeypatch.setenv("SCS_LRT_TOP_CAP", top_cap)
    out, xt = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    _assert_same_lrt(base, out, xb, xt)


@pytest.mark.parametrize("n_leaves,n_prob", [(400, 5), (1000, 10), (3000, 3), (10000, 10), (12000, 2)])
def test_lrt_tree_kernel_on_big_trees(ctx, monkeypatch, n_leaves, n_prob):
    """Big random trees: the default (cluster kernel, automatic cluster size) against one CTA per problem with the
    working set in global memory (the round-1 kernel).  12,000 leaves has levels below the cut by itself."""
    probs = [_synthetic_lrt_problem(n_leaves, 100 * n_leaves + i, sparse=(i % 3 == 2)) for i in range(n_prob)]
    Ys, ws, chs = [p[0] for p in probs], [p[1] for p in probs], [p[2] for p in probs]
    out, xt = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    monkeypatch.setenv("SCS_LRT_KERNEL", "cta")
    base, xb = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    assert (base[:, 7] == 0).sum() >= n_prob - 1
    _assert_same_lrt(base, out, xb, xt)
    # gather + reuse through a plan (what bench.py and the product path do): counts picked from a longer source vector
    monkeypatch.delenv("SCS_LRT_KERNEL")
    import torch
    rng = np.random.default_rng(5)
    total = sum(len(y) for y in Ys)
    perm = rng.permutation(total + 7).astype(np.int32)[:total]
    src = np.zeros(total + 7)
    src[perm] = np.concatenate(Ys)
    plan = engine.LrtPlan(ctx, chs)
    plan.set_weights(ws)
    plan.set_gather(perm, total + 7)
    d_out = torch.zeros((n_prob, 8), dtype=torch.float64, device="cuda")
    for _ in range(2):
        plan.run(torch.as_tensor(src, device="cuda"), d_out)
    torch.cuda.synchronize()
    plan.close()
    _assert_same_lrt(out, d_out.cpu().numpy())


def test_lrt_second_attempt_after_a_failed_solve(ctx, monkeypatch):
    """run_PT_test.py:671-678 retries a failed trust-constr solve with SLSQP and returns the failure tuple only when both
    fail.  Here: scs_lrt_poisson_batch re-solves the problems whose first attempt ended with status 1 / 3 by path following
    over eight barrier parameters.  The first attempt is made to fail (SCS_LRT_MAX_ITER=2 Newton steps): without the retry
    every problem reports status 1, with it every problem comes back converged at the answer of an undisturbed run --
    small trees (warp kernel first) and big ones (cluster kernel first)."""
    g = load("lrt_sweep")
    off = g["off"]
    n = len(off) - 1
    keep = [i for i in range(n) if g["ll_H1"][i] >= 0][:120]
    Ys = [g["Y"][off[i]:off[i + 1]] for i in keep]
    ws = [g["w"][off[i]:off[i + 1]] for i in keep]
    chs = [g["child_int"][off[i]:off[i + 1]] for i in keep]
    big = [_synthetic_lrt_problem(1500, 4242 + i, sparse=(i == 1)) for i in range(2)]
    Ys += [p[0] for p in big]
    ws += [p[1] for p in big]
    chs += [p[2] for p in big]
    base, xb = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    assert (base[:, 7] == 0).all()
    monkeypatch.setenv("SCS_LRT_MAX_ITER", "2")
    monkeypatch.setenv("SCS_LRT_NO_RETRY", "1")
    first, _ = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=False)
    assert (first[:, 7] == 1).all()                                  # the first attempt alone: iteration limit everywhere
    monkeypatch.delenv("SCS_LRT_NO_RETRY")
    again, xa = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    assert (again[:, 7] == 0).all()
    assert (again[:, 6] > base[:, 6]).all()                          # it took the long way round
    assert np.array_equal(again[:, 1], base[:, 1]) and np.array_equal(again[:, 2], base[:, 2])      # dof, on_bound
    np.testing.assert_allclose(again[:, 0], base[:, 0], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(again[:, 3], base[:, 3], rtol=1e-5, atol=1e-300)
    for a, b in zip(xa, xb):
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-9)


# ------------------------------------------------- the product path under an initialised process group
def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tsv_fields(path):
    lines = open(path).read().split("\n")
    return lines[0].split("\t"), lines[1].split("\t")


def _assert_same_tsv(a, b, lr_rtol):
    (ha, fa), (hb, fb) = _tsv_fields(a), _tsv_fields(b)
    assert ha == hb and fa[:3] == fb[:3]
    for i in range(3, len(fa), 4):
        assert abs(float(fa[i]) - float(fb[i])) <= lr_rtol * max(1.0, abs(float(fb[i]))) + 2e-3       # 3 printed decimals
        assert fa[i + 1] == fb[i + 1] and fa[i + 3] == fb[i + 3]
        assert abs(float(fa[i + 2]) - float(fb[i + 2])) <= 2e-3 * float(fb[i + 2]) + 1e-300


def test_two_ranks_row_sharded_dataset_and_replicate_sweep(ctx, tmp_path):
    """torchrun with two ranks (sharing this GPU: gloo carries the all-reduces, NCCL refuses duplicate devices).
    One dataset: each rank decodes and places its own block of VCF record lines, kept-SNV counts, per-cell
    missing counts and the soft branch weights are all-reduced, rank 0 writes the TSV -- equal to the
    single-rank TSV.  Replicate sweep: whole pairs are dealt to the ranks and every pair is decoded from ALL
    of its lines (row sharding is opt-in and the sweep driver does not opt in): each TSV equals the single-rank one."""
    import json
    import subprocess
    import sys
    names = ["syn_d_clock", "syn_a", "syn_d_noclock", "syn_e_mismatch", "example_clock"]
    gs = [load(n) for n in names]
    vcfs, trees = zip(*[case_inputs(g)[:2] for g in gs])
    w = [100.0, 500.0, 1000.0]
    # single rank, this process
    one_ref = str(tmp_path / "one_ref.tsv")
    P.run_poisson_tree_test(vcfs[0], trees[0], one_ref, w, "", "", None, None)
    ref_outs = [str(tmp_path / ("ref_%s.tsv" % n)) for n in names]
    P.run_poisson_tree_test_batch(list(vcfs), list(trees), ref_outs, w, ctx=ctx)
    # two ranks
    outs = [str(tmp_path / ("r2_%s.tsv" % n)) for n in names]
    spec = {"world": 2, "w_maxs": w, "one": {"vcf": vcfs[0], "tree": trees[0], "out": str(tmp_path / "one_r2.tsv")},
            "sweep": {"vcfs": list(vcfs), "trees": list(trees), "outs": outs}, "log": str(tmp_path / "ranks.json")}
    spec_file = str(tmp_path / "spec.json")
    json.dump(spec, open(spec_file, "w"))
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "two_rank_product.py")
    env = dict(os.environ, SCS_DIST_BACKEND="gloo", SCS_DEVICE=str(ctx.device))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), script, spec_file]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    _assert_same_tsv(spec["one"]["out"], one_ref, 1e-5)
    pairs = sorted(sum((json.load(open(spec["log"] + ".%d" % k))["pairs"] for k in range(2)), []))
    assert pairs == list(range(len(names)))                      # every pair handled by exactly one rank
    for a, b in zip(outs, ref_outs):
        _assert_same_tsv(a, b, 1e-6)


# ------------------------------------------------- per-cell false-negative rates
def test_per_cell_fn_product_path_on_golden_cases(ctx):
    """map_mutations_gt with a pd.Series of per-cell FN rates (run_PT_test.py:413-418) through the product mirror -- VCF
    decode, read_tree, row filter, scs_place_percell -- against what the unmodified reference returned (percell_fn.npz).
    Tolerance: Y_TOL like every soft-weight comparison (fp32 softmax terms on float64 logit differences)."""
    import pandas as pd
    from conftest import REPO
    pc = np.load(golden_path("percell_fn"), allow_pickle=False)
    for name in pc["cases"]:
        name = str(name)
        muts, _, _ = P.get_mut_df(os.path.join(REPO, str(pc[name + "_vcf"])), "", "", ctx=ctx)
        assert list(muts.columns) == list(pc[name + "_all_cols"])
        tree, outg, FP0, FN0 = P.read_tree(os.path.join(REPO, str(pc[name + "_tree"])), muts.columns)
        FP, rates, cols = float(pc[name + "_FP"]), pc[name + "_FN_cells"], [str(c) for c in pc[name + "_cols"]]
        want = pc[name + "_soft"]
        M = P.map_mutations_gt(tree, muts, FP, pd.Series(rates, index=cols))
        assert M.shape == (int(pc[name + "_n_muts"]), want.size)
        assert M.columns == [str(c) for c in pc[name + "_node_names"]]
        assert_soft_close(M.soft, want)
        for node, y in zip(tree.iter_descendants(), want):                   # :443-445
            assert abs(node.dist - y) <= Y_TOL * max(1.0, abs(y)) and node.mut_no_soft == node.dist
        # the same rates as a dict and as a plain array over the placed cells
        assert np.array_equal(P.map_mutations_gt(tree, muts, FP, dict(zip(cols, rates))).soft, M.soft)
        assert np.array_equal(P.map_mutations_gt(tree, muts, FP, rates).soft, M.soft)
        with pytest.raises(KeyError):
            P.map_mutations_gt(tree, muts, FP, dict(zip(cols[1:], rates[1:])))
        with pytest.raises(TypeError):
            P.map_mutations_gt(tree, muts, rates, 0.1)


@pytest.mark.parametrize("n_snv,n_cells", [(1, 2), (127, 3), (300, 50), (1000, 200), (700, 513), (400, 1100), (600, 2500), (150, 4200),
                                            (300, 7000), (60, 10000), (50, 12287), (40, 12500), (30, 13500), (24, 16000)])
def test_per_cell_fn_kernel_matches_oracle(ctx, monkeypatch, n_snv, n_cells):
    """scs_place_percell against the oracle's float64 prefix-sum form (pinned to the reference by
    test_per_cell_fn_against_the_reference, and to the dense form below 600 cells) on shuffled random trees: 64..1024
    threads per CTA, accumulators / cell order in shared memory (up to 12,700 / 14,300 cells) and in global memory beyond,
    rows dropped by the filter, rates spread over three decades.  With equal rates it must agree with the count kernels."""
    import torch
    codes, S = shuffled_problem(n_snv, n_cells, seed=5 * n_snv + n_cells)
    rng = np.random.default_rng(n_cells)
    keep = np.ones(n_snv, dtype=np.uint8)
    if n_snv > 4:
        keep[::5] = 0
    rates = np.exp(rng.uniform(np.log(5e-4), np.log(0.6), n_cells))
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    dev = torch.device("cuda", ctx.device)
    d_gt = torch.from_numpy(packed).to(dev)
    d_keep = torch.from_numpy(keep).to(dev)
    y = torch.zeros(S.shape[0], dtype=torch.float64, device=dev)
    muts = np.where(codes == 3, np.nan, codes.astype(float))[keep.astype(bool)]
    ms = engine.place_percell(ctx, d_gt, pitch, n_snv, n_cells, d_keep, bits, words, 0.01, rates, y)
    got = y.cpu().numpy()
    assert ms >= 0.0
    assert_soft_close(got, O.map_mutations_gt_interval(muts, S.astype(float), 0.01, rates))
    if n_cells <= 600:
        assert_soft_close(got, O.map_mutations_gt(muts, S.astype(float), 0.01, rates))
    assert abs(got.sum() - keep.sum()) <= 5e-6 * max(1, keep.sum())          # every kept SNV's row of M sums to 1
    engine.place_percell(ctx, d_gt, pitch, n_snv, n_cells, d_keep, bits, words, 0.01, rates, y)
    assert np.array_equal(y.cpu().numpy(), got)                               # fixed reduction order
    # trees of 1,024 .. 12,287 cells took the register kernel (scs_place_percell2_kernel<R>): the general kernel on the same input
    if 2048 <= S.shape[0] <= 48 * 512:
        monkeypatch.setenv("SCS_PERCELL_KERNEL", "1")
        engine.place_percell(ctx, d_gt, pitch, n_snv, n_cells, d_keep, bits, words, 0.01, rates, y)
        monkeypatch.delenv("SCS_PERCELL_KERNEL")
        assert_soft_close(y.cpu().numpy(), O.map_mutations_gt_interval(muts, S.astype(float), 0.01, rates))
        np.testing.assert_allclose(y.cpu().numpy(), got, rtol=1e-5, atol=1e-6)
    # no row flags: every row placed
    engine.place_percell(ctx, d_gt, pitch, n_snv, n_cells, None, bits, words, 0.01, rates, y)
    assert abs(y.sum().item() - n_snv) <= 5e-6 * n_snv
    # equal rates: the scalar branch, i.e. the count kernels
    engine.place_percell(ctx, d_gt, pitch, n_snv, n_cells, d_keep, bits, words, 0.02, np.full(n_cells, 0.15), y)
    pl = engine.Placement(ctx, n_snv, n_cells, S.shape[0])
    pl.set_clades_host(bits, words)
    pl.set_genotypes_host(packed, pitch, keep)
    assert_soft_close(y.cpu().numpy(), pl.run_host(0.02, 0.15))
    pl.close()


def test_per_cell_fn_argument_checks(ctx):
    import torch
    codes, S = shuffled_problem(50, 20, seed=3)
    packed, pitch = engine.pack_genotypes(codes)
    bits, words = engine.pack_clades(S)
    dev = torch.device("cuda", ctx.device)
    d_gt = torch.from_numpy(packed).to(dev)
    y = torch.zeros(S.shape[0], dtype=torch.float64, device=dev)
    rates = np.full(20, 0.1)
    for bad in (0.0, 1.0, -0.1, np.nan):
        r = rates.copy()
        r[7] = bad
        with pytest.raises(engine._cabi.ScsError):
            engine.place_percell(ctx, d_gt, pitch, 50, 20, None, bits, words, 0.01, r, y)
    with pytest.raises(engine._cabi.ScsError):
        engine.place_percell(ctx, d_gt, pitch, 50, 20, None, bits, words, 1.5, rates, y)
    with pytest.raises(ValueError):
        engine.place_percell(ctx, d_gt, pitch, 50, 20, None, bits, words, 0.01, rates[:-1], y)
    crossing = S.copy()
    crossing[0] = 0
    crossing[0, [0, 1]] = 1
    crossing[1] = 0
    crossing[1, [1, 2]] = 1                                                   # two rows that overlap without nesting
    cb, cw = engine.pack_clades(crossing)
    with pytest.raises(engine._cabi.ScsError, match="laminar"):
        engine.place_percell(ctx, d_gt, pitch, 50, 20, None, cb, cw, 0.01, rates, y)
    # a column outside every clade (the outgroup) needs no rate
    S2 = np.concatenate([S, np.zeros((S.shape[0], 1), dtype=S.dtype)], axis=1)
    codes2 = np.concatenate([codes, np.full((50, 1), 1, dtype=codes.dtype)], axis=1)
    p2, pitch2 = engine.pack_genotypes(codes2)
    b2, w2 = engine.pack_clades(S2)
    engine.place_percell(ctx, torch.from_numpy(p2).to(dev), pitch2, 50, 21, None, b2, w2, 0.01, np.append(rates, np.nan), y)
    want = y.cpu().numpy().copy()
    engine.place_percell(ctx, d_gt, pitch, 50, 20, None, bits, words, 0.01, rates, y)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-12, atol=1e-12)
