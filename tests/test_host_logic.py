"""CPU tests of the host-side logic: tree surgery, clade masks, LRT model data,
packing, and that the C-ABI library exports what the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REPO, case_inputs, golden_path
from oracle import pt_oracle as O
from scsommerclock_b200 import engine, run_PT_test as P, synth
from scsommerclock_b200.tree import NewickError, Tree

CASES = ["example_clock", "example_noclock", "syn_a", "syn_b", "syn_c", "syn_c_incl", "syn_d_clock", "syn_d_noclock", "syn_e_mismatch",
         "syn_f_scite_log", "syn_g_scite_csv", "syn_h_path_rates"]


def load(name):
    return np.load(golden_path(name), allow_pickle=False)


def _mut_from_golden(g):
    c = g["muts_codes"].astype(float)
    c[c == 3] = np.nan
    return {"muts": c, "samples": g["samples"]}


def test_header_declares_what_the_library_exports():
    hdr = open(os.path.join(REPO, "include", "scsommerclock_b200.h")).read()
    declared = set(re.findall(r"\b(scs_[a-z0-9_]+)\s*\(", hdr))
    lib_path = os.path.join(REPO, "scsommerclock_b200", "libscs_b200.so")
    if not os.path.exists(lib_path):
        from scsommerclock_b200 import build
        build.build()
    L = ctypes.CDLL(lib_path)      # loads without a GPU: no compute call is made here
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    from scsommerclock_b200 import _cabi
    assert declared == set(_cabi.EXPORTED_SYMBOLS)
    assert L.scs_version is not None


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine._cabi.ScsError):
        engine.Context(0)


def test_newick_formats_and_errors():
    t = Tree("((a:1,b:2):0.5,(c:1,(d:1,e:1):2):1,f:3);", format=1)
    assert t.get_leaf_names() == list("abcdef") and len(t) == 6
    assert [n.name for n in t.traverse()] == ["", "", "", "f", "a", "b", "c", "", "d", "e"]
    with pytest.raises(NewickError):
        Tree("((a:1,b:2):0.5,(c:1,(d:1,e:1):2):1,f:3);", format=2)
    t2 = Tree("((a:1,b:2)0.9:0.5,c:3)1:0;", format=2)
    assert t2.children[0].support == 0.9
    with pytest.raises(NewickError):
        Tree("((a:1,b:2):0.5,(c:1", format=1)
    with pytest.raises(NewickError):
        Tree("(a:1,,b:1);", format=1)


def test_set_outgroup_and_delete_child_order():
    t = Tree("((a:1,b:2):0.5,(c:1,(d:1,e:1):2):1,f:3);", format=1)
    t.set_outgroup(t & "d")
    assert t.write() == "(d:0.5,(e:1,(c:1,((a:1,b:2):0.5,f:3):1):2):0.5);"
    (t & "d").delete()
    assert len(t.children) == 1                     # the root keeps a single child, like ete3
    (t & "c").delete()                              # parent collapses, its other child is appended to the grandparent
    assert [n.name for n in t.iter_leaves()] == ["e", "a", "b", "f"]


@pytest.mark.parametrize("name", CASES)
def test_read_tree_and_clade_bits_match_golden_recorded_on_the_ete3_stand_in(name):
    """read_tree + the clade rows against the fixtures.  What this pins and what it does not: the fixtures were recorded by the
    unmodified reference's read_tree (:210-334) running on scsommerclock_b200/tree.py, the restatement of the ete3 3.1.x
    TreeNode methods it calls (ete3 is not installable here) -- so the reference's renames, regexes, outgroup handling, leaf
    pruning and S construction are pinned to the reference's code, the tree surgery underneath (set_outgroup, delete,
    traversal order) to that restatement, whose ete3 semantics are checked by hand-derived cases in
    test_newick_formats_and_errors / test_set_outgroup_and_delete_child_order, not by ete3 binaries."""
    g = load(name)
    vcf, tree_file, meta = case_inputs(g)
    tree, outg, FP, FN = P.read_tree(tree_file, g["samples"])
    assert [n.name for n in tree.iter_descendants()] == list(g["node_names"])
    columns = list(g["samples"])
    active = np.array([c != outg for c in columns], dtype=np.uint8)
    bits, words, nodes = P._clade_bits(tree, columns, active)
    S = engine.unpack_clades(bits, len(columns))[:, active.astype(bool)]
    assert np.array_equal(S.astype(np.uint8), g["S"])                   # bit exact


@pytest.mark.parametrize("name", CASES)
def test_get_model_data_matches_golden(name):
    g = load(name)
    vcf, tree_file, meta = case_inputs(g)
    for k, w in enumerate(g["w_maxs"]):
        r = O.get_gt_tree(tree_file, _mut_from_golden(g), float(w), meta["FN_in"], meta["FP_in"])   # oracle fills the node attributes
        Y, constr, init, wn, cols = P.get_model_data(r["tree"])
        assert list(cols) == list(g["constr_cols"][k])
        assert np.array_equal(constr.toarray(), g["constr"][k])
        np.testing.assert_allclose(Y, g["Y"][k], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(init, g["init"][k], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(wn[:, 0], g["weights_norm"][k], rtol=1e-12)
        back = P.ClockConstraints.from_dense(g["constr"][k])
        assert np.array_equal(back.child_int, constr.child_int)
        assert np.array_equal(back.toarray(), g["constr"][k])
        np.testing.assert_allclose(constr @ init, g["constr"][k] @ init, atol=1e-9)
        # the sweep driver's w_max-independent walk gives the same branch order and topology
        order, child_int = P._branch_order(r["tree"])
        assert [n.name for n in order] == list(cols)
        assert np.array_equal(child_int, constr.child_int)


def test_packing_round_trips():
    rng = np.random.default_rng(3)
    for n_cells in (1, 3, 4, 31, 32, 33, 257):
        codes = rng.integers(0, 4, size=(17, n_cells)).astype(np.uint8)
        packed, pitch = engine.pack_genotypes(codes)
        assert pitch % 16 == 0 and packed.shape == (17, pitch)
        assert np.array_equal(engine.unpack_genotypes(packed, n_cells), codes)
        S = rng.integers(0, 2, size=(9, n_cells)).astype(np.uint8)
        bits, words = engine.pack_clades(S)
        assert np.array_equal(engine.unpack_clades(bits, n_cells), S.astype(bool))
    f = np.array([[0.0, 1.0, 2.0, np.nan]])
    assert np.array_equal(engine.unpack_genotypes(engine.pack_genotypes(f)[0], 4), [[0, 1, 2, 3]])


def test_synthetic_tree_is_a_laminar_family():
    rng = np.random.default_rng(0)
    t = synth.coalescent_tree(37, rng)
    S = t.clade_matrix()
    assert S.shape == (74, 37) and S[-2].sum() == 37 and S[-1].sum() == 0
    for v in range(37, 73):
        a, b = t.left[v - 37], t.right[v - 37]
        assert (S[v] == (S[a] | S[b])).all() and not (S[a] & S[b]).any()
    nw = t.newick(["c%d" % i for i in range(37)], "outg")
    assert len(Tree(nw, format=1)) == 38


def test_vcf_record_lines_are_sharded_at_line_boundaries():
    """Row sharding of one large VCF over ranks: every rank gets the header block plus a contiguous
    block of whole record lines, and the blocks tile the body."""
    import unittest.mock as mock
    import torch.distributed as dist
    head = b"##fileformat=VCFv4.3\n#CHROM\tPOS\tID\n\n"
    text = head + b"".join(b"1\t%d\tsnv%06d\tpayload %s\n" % (i, i, b"x" * (i % 17)) for i in range(1, 700))
    for world in (1, 2, 3, 8):
        got = []
        for rank in range(world):
            with mock.patch.object(dist, "get_rank", return_value=rank), mock.patch.object(dist, "get_world_size", return_value=world):
                got.append(P._shard_record_lines(text))
        assert all(g.startswith(head) and g.endswith(b"\n") for g in got)
        assert b"".join(g[len(head):] for g in got) == text[len(head):]
    # more ranks than lines: some ranks get an empty body, nothing is lost
    tiny = head + b"1\t5\tsnv\n"
    got = []
    for rank in range(4):
        with mock.patch.object(dist, "get_rank", return_value=rank), mock.patch.object(dist, "get_world_size", return_value=4):
            got.append(P._shard_record_lines(tiny)[len(head):])
    assert b"".join(got) == b"1\t5\tsnv\n"


def test_clade_intervals_of_tree_masks_and_of_general_masks():
    """The interval path's host analysis: tree-shaped masks (any cell order, empty and duplicate
    rows included) are recognised and every row is exactly its interval in the returned order;
    a matrix that is not a laminar family is rejected."""
    from scsommerclock_b200 import synth
    rng = np.random.default_rng(5)
    for n in (2, 3, 33, 100, 257):
        ts = synth.coalescent_tree(n, rng)
        cols = rng.permutation(n + 1)                       # leaf rank -> column (one extra column: the outgroup, in no clade)
        S = np.zeros((2 * n + 1, n + 1), dtype=np.uint8)    # all tree nodes, one duplicate row, one empty row
        inv = np.zeros(n, dtype=np.int64)
        inv[ts.leaf_rank] = np.arange(n)                    # depth-first rank -> cell
        for b in range(2 * n - 1):
            S[b, cols[inv[ts.lo[b]:ts.hi[b]]]] = 1
        S[2 * n - 1] = S[0]
        bits, words = engine.pack_clades(S)
        ok, perm, lo, hi = engine.clade_intervals(bits, n + 1)
        assert ok
        assert sorted(perm.tolist()) == list(range(n + 1))
        for b in range(S.shape[0]):
            members = np.zeros(n + 1, dtype=np.uint8)
            members[perm[lo[b]:hi[b]]] = 1
            assert np.array_equal(members, S[b]), (n, b)
    S = np.array([[1, 1, 0, 0], [0, 1, 1, 0], [0, 0, 1, 1]], dtype=np.uint8)      # overlapping, not nested
    bits, words = engine.pack_clades(S)
    assert not engine.clade_intervals(bits, 4)[0]
    S = (rng.random((40, 70)) < 0.3).astype(np.uint8)
    bits, words = engine.pack_clades(S)
    assert not engine.clade_intervals(bits, 70)[0]


def test_vcf_texts_are_prefetched_in_order(tmp_path):
    """The sweep driver reads / gunzips VCFs ahead on host threads: same bytes, same order, any window."""
    import gzip
    files = []
    for k in range(7):
        body = ("##fileformat=VCFv4.2\n#CHROM\tPOS\n" + "".join("1\t%d\n" % (k * 100 + j) for j in range(50 * (k + 1)))).encode()
        f = tmp_path / ("v%d.vcf%s" % (k, ".gz" if k % 2 else ""))
        if k % 2:
            with gzip.open(f, "wb") as fh:
                fh.write(body)
        else:
            f.write_bytes(body)
        files.append(str(f))
    want = [P._read_text(f) for f in files]
    for workers, window in [(1, None), (2, 1), (3, 2), (4, None), (8, 50)]:
        assert list(P._prefetch_texts(files, workers=workers, window=window)) == want
    assert list(P._prefetch_texts([])) == []


def interval_counts(codes, S):
    """CPU restatement of the interval path's integer arithmetic (csrc/scs_place_interval.cuh): the cell
    order and the intervals come from the library's host analysis (scs_clade_intervals), the counts are
    differences of per-SNV prefix counts in that order."""
    bits, words = engine.pack_clades(S)
    ok, perm, lo, hi = engine.clade_intervals(bits, S.shape[1])
    assert ok
    c = codes[:, perm]                                            # depth-first cell order
    z = np.zeros((codes.shape[0], 1), dtype=np.int64)
    p1 = np.concatenate([z, np.cumsum((c == 1) | (c == 2), axis=1)], axis=1)
    p0 = np.concatenate([z, np.cumsum(c == 0, axis=1)], axis=1)
    return p1[:, hi] - p1[:, lo], p0[:, hi] - p0[:, lo]


@pytest.mark.parametrize("name", CASES)
def test_interval_counts_equal_the_dense_contraction_on_golden_cases(name):
    """run_PT_test.py:421-430 contracts the genotypes against S; for a tree's S the same integers are
    prefix-count differences.  Checked here on the reference's own clade matrices (fixtures), cell by
    cell, without a GPU; the soft weights built from those counts reproduce the recorded ones."""
    g = load(name)
    S = g["S"].astype(np.uint8)
    codes = g["muts_codes"]
    samples = list(g["samples"])
    if codes.shape[1] != S.shape[1]:                              # the fixture's S has no outgroup column
        keep_cols = [i for i, c in enumerate(samples) if c != "healthycell"]
        codes = codes[:, keep_cols]
    assert codes.shape[1] == S.shape[1]
    C1, C0 = interval_counts(codes, S)
    x1 = ((codes == 1) | (codes == 2)).astype(np.int64)
    x0 = (codes == 0).astype(np.int64)
    assert np.array_equal(C1, x1 @ S.T.astype(np.int64)) and np.array_equal(C0, x0 @ S.T.astype(np.int64))
    # the softmax over clade rows of a1 C1 + a0 C0 is the reference's normalised placement (:368-383, :421-430)
    FP, FN = float(g["FP"]), float(g["FN"])
    a1, a0 = np.log((1 - FN) / FP), np.log(FN / (1 - FP))
    keep = x1.any(axis=1) if "healthycell" in samples else np.ones(codes.shape[0], dtype=bool)
    logit = a1 * C1[keep] + a0 * C0[keep]
    M = np.exp(logit - logit.max(axis=1, keepdims=True))
    soft = (M / M.sum(axis=1, keepdims=True)).sum(axis=0)
    np.testing.assert_allclose(soft, g["soft"], rtol=1e-9, atol=1e-9)


def test_interval_counts_on_random_trees_with_shuffled_columns():
    from scsommerclock_b200 import synth
    rng = np.random.default_rng(11)
    for n in (2, 5, 64, 301):
        tree = synth.coalescent_tree(n, rng)
        S = tree.clade_matrix()
        codes = synth.observe(tree, synth.throw_mutations(tree, 200, rng), 0.1, 0.01, 0.2, rng)
        cols = rng.permutation(n)
        S, codes = np.ascontiguousarray(S[:, cols]), np.ascontiguousarray(codes[:, cols])
        C1, C0 = interval_counts(codes, S)
        x1 = ((codes == 1) | (codes == 2)).astype(np.int64)
        x0 = (codes == 0).astype(np.int64)
        assert np.array_equal(C1, x1 @ S.T.astype(np.int64)) and np.array_equal(C0, x0 @ S.T.astype(np.int64))


def test_clade_intervals_on_general_laminar_families():
    """Any laminar family is accepted (multifurcating trees, duplicate and empty rows, cells in no clade),
    a family with one crossing pair is rejected; accepted orders make every row exactly its interval --
    which is all the interval path's integer arithmetic needs."""
    rng = np.random.default_rng(2024)

    def random_family(cells, out):
        if len(cells) < 2:
            return
        k = int(rng.integers(2, min(5, len(cells)) + 1))              # arity 2..5
        cuts = np.sort(rng.choice(np.arange(1, len(cells)), size=k - 1, replace=False))
        for part in np.split(cells, cuts):
            if rng.random() < 0.85:
                out.append(part)
            random_family(part, out)

    def is_laminar(rows):
        sets = [frozenset(np.nonzero(r)[0].tolist()) for r in rows]
        return all(not (a & b) or a <= b or b <= a for i, a in enumerate(sets) for b in sets[i + 1:])

    for n in (3, 17, 90):
        for trial in range(6):
            cells = rng.permutation(n)
            fam = [cells[: n - 2]]                                   # the last two cells sit in no clade
            random_family(cells[: n - 2], fam)
            S = np.zeros((len(fam) + 2, n), dtype=np.uint8)
            for r, part in enumerate(fam):
                S[r, part] = 1
            S[-2] = S[int(rng.integers(0, len(fam)))]                # a duplicate; the last row stays empty
            S = S[rng.permutation(S.shape[0])]
            assert is_laminar(S)
            bits, words = engine.pack_clades(S)
            ok, perm, lo, hi = engine.clade_intervals(bits, n)
            assert ok, (n, trial)
            for b in range(S.shape[0]):
                members = np.zeros(n, dtype=np.uint8)
                members[perm[lo[b]:hi[b]]] = 1
                assert np.array_equal(members, S[b])
            # one crossing pair: a row that takes one cell of each of two disjoint rows
            big = [r for r in range(S.shape[0]) if S[r].sum() >= 2]
            pairs = [(a, b) for a in big for b in big if a < b and not (S[a] & S[b]).any()]
            if pairs:
                a, b = pairs[int(rng.integers(0, len(pairs)))]
                bad = np.zeros(n, dtype=np.uint8)
                bad[np.nonzero(S[a])[0][0]] = 1
                bad[np.nonzero(S[b])[0][0]] = 1
                S2 = np.concatenate([S, bad[None, :]], axis=0)
                assert not is_laminar(S2)
                bits2, _ = engine.pack_clades(S2)
                assert not engine.clade_intervals(bits2, n)[0]


# ---------------------------------------------------------------- replicate sweeps: native host work
def _python_tree_rows(tree_file, columns):
    tree, outg, FP, FN = P.read_tree(tree_file, np.array(columns))
    active = np.array([c != outg for c in columns], dtype=np.uint8)
    bits, words, nodes = P._clade_bits(tree, columns, active)
    idx = {id(n): i for i, n in enumerate(nodes)}
    ch = np.full((bits.shape[0], 2), -1, dtype=np.int32)
    for i, n in enumerate(nodes):
        if n.children:
            ch[i] = [idx[id(c)] for c in n.children]
    return bits, ch, len(nodes), len(tree)


@pytest.mark.parametrize("name", [c for c in CASES if "scite" not in c])
def test_native_tree_parser_matches_read_tree_on_golden_trees(name):
    """scs_tree_parse_batch (C++, host threads) against read_tree + _clade_bits (which the fixtures pin to the reference's
    S): same rows, same child order, for the reference's example trees and every synthetic case."""
    g = load(name)
    vcf, tree_file, meta = case_inputs(g)
    columns = list(g["samples"])
    bits, ch, n_nodes, n_leaves = _python_tree_rows(tree_file, columns)
    kind, raw, FP, FN = P._tree_source(tree_file)
    b2, words, c2, info = engine.tree_parse_batch([raw], columns, bits.shape[0], strip_tags=(kind == "cellphy"))
    assert np.array_equal(b2[0], bits) and np.array_equal(c2[0], ch) and tuple(info[0][:2]) == (n_nodes, n_leaves)
    tree, outg, FP0, FN0 = P.read_tree(tree_file, np.array(columns))
    assert (FP, FN) == (FP0, FN0)                       # the file-level rates of _tree_source are read_tree's


def test_native_tree_parser_on_random_trees_with_missing_samples(tmp_path):
    rng = np.random.default_rng(5)
    n = 40
    names = ["cell%04d" % (i + 1) for i in range(n)]
    cols = [c for i, c in enumerate(["tumcell%04d" % (i + 1) for i in range(n)] + ["healthycell"]) if i not in (3, 17, 22)]
    texts, want = [], []
    for k in range(60):
        st = synth.coalescent_tree(n, rng)
        nw = st.newick(names, "outgcell", scale=0.05) + "\n"
        tf = tmp_path / ("t%d.newick" % k)
        tf.write_text(nw)
        texts.append(nw)
        want.append(_python_tree_rows(str(tf), cols))
    n_rows = want[0][0].shape[0]
    for threads in (1, 0):
        bits, words, ch, info = engine.tree_parse_batch(texts, cols, n_rows, threads=threads)
        for k in range(len(texts)):
            assert np.array_equal(bits[k], want[k][0]) and np.array_equal(ch[k], want[k][1]) and tuple(info[k][:2]) == want[k][2:]
    for bad in ["((a:1,b:1):1,c:1)", "((tumcell0001:1,,tumcell0002:1):1,healthycell:1);", "((tumcell0001:1,tumcell0002:1):1,tumcell0005:1);",
                "((tumcell0001:1,tumcell0002:x):1,healthycell:1);", "((tumcell0001:1,tumcell0002:1,tumcell0005:1):1,healthycell:1);"]:
        with pytest.raises(engine._cabi.ScsError):
            engine.tree_parse_batch([texts[0], bad], cols, n_rows)


def test_native_tsv_rows_match_python_formatting():
    rng = np.random.default_rng(0)
    vals = [0.0, 1.0, 0.5, 0.05, 0.0001, 0.0003, 9.999e-5, 1.5e-7, 1e-100, 3.3e-16, 0.6725292599860104, 0.049999999, 1e-5, 2.5e-5, 0.1,
            0.30000000000000004, 5e-324, 1e-300] + list(10 ** rng.uniform(-20, 0, 300))
    out = np.zeros((len(vals), 2, 8))
    for i, v in enumerate(vals):
        out[i, 0] = [rng.uniform(0, 200), 29, 1, v, 0, 0, 5, 0]
        out[i, 1] = [rng.uniform(0, 2), 13, 0, min(1.0, v * 3), 0, 0, 5, 0]
    out[3, 1, 7] = 1
    out[3, 1, 0] = np.nan
    nk = rng.integers(1, 100000, len(vals))
    FN = rng.uniform(0, 0.5, len(vals))
    FP = rng.uniform(0, 0.1, len(vals))
    rows = engine.format_pt_rows(out, nk, FN, FP)
    for i in range(len(vals)):
        s = f"{nk[i]}\t{FN[i]:.4f}\t{FP[i]:.4f}"
        for w in range(2):
            LR, dof, ob, p, _, _, _, st = out[i, w]
            if st != 0 or not np.isfinite(LR):
                LR = dof = p = np.nan
                s += f"\t{LR:0>5.3f}\t{dof}\t{p}\tH{int(p < 0.05)}"
            else:
                s += f"\t{LR:0>5.3f}\t{int(dof)}\t{float(p)}\tH{int(p < 0.05)}"
        assert s == rows[i]


def test_sub_batch_cuts_cover_every_pair_once():
    """engine.sub_batch_cuts: strictly increasing boundaries from 0 to the number of pairs, at most n_sub pieces, a short
    first and last piece when there are at least three."""
    for G in (1, 2, 3, 5, 17, 150, 1250):
        for n_sub in (1, 2, 3, 4, 7, 12, 5000):
            cuts = engine.sub_batch_cuts(G, n_sub)
            assert cuts[0] == 0 and cuts[-1] == G and all(b > a for a, b in zip(cuts, cuts[1:]))
            assert len(cuts) - 1 <= max(1, min(n_sub, G))
    cuts = engine.sub_batch_cuts(1250, 4)
    sizes = np.diff(cuts)
    assert list(sizes) == [208, 417, 417, 208]


def test_per_cell_rates_alignment():
    """FN[muts_in.columns].values of run_PT_test.py:414 in the mirror: Series / dict keyed by sample name, or a plain
    sequence over all columns or over the placed cells; the outgroup column may be absent."""
    import pandas as pd
    from scsommerclock_b200 import run_PT_test as P
    columns = ["tumcell0002", "healthycell", "tumcell0001", "tumcell0003"]
    active = np.array([1, 0, 1, 1], dtype=np.uint8)
    ser = pd.Series({"tumcell0001": 0.1, "tumcell0002": 0.2, "tumcell0003": 0.3})
    want = np.array([0.2, np.nan, 0.1, 0.3])
    for FN in (ser, dict(ser), [0.2, 0.1, 0.3], np.array([0.2, 0.9, 0.1, 0.3])):
        got = P._per_cell_rates(FN, columns, active)
        np.testing.assert_array_equal(np.isnan(got[active == 1]), False)
        np.testing.assert_allclose(got[active == 1], want[active == 1])
    with pytest.raises(KeyError):
        P._per_cell_rates({"tumcell0001": 0.1, "tumcell0002": 0.2}, columns, active)
    with pytest.raises(ValueError):
        P._per_cell_rates([0.1, 0.2], columns, active)
