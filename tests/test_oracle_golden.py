"""The CPU oracle (oracle/pt_oracle.py) against fixtures recorded from the
unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import case_inputs, golden_path
from oracle import pt_oracle as O

CASES = ["example_clock", "example_noclock", "syn_a", "syn_b", "syn_c", "syn_c_incl", "syn_d_clock", "syn_d_noclock", "syn_e_mismatch",
         "syn_f_scite_log", "syn_g_scite_csv", "syn_h_path_rates"]
SYN = [c for c in CASES if c.startswith("syn") and not c.startswith("syn_d")]


def load(name):
    return np.load(golden_path(name), allow_pickle=False)


def codes(muts):
    return np.where(np.isnan(muts), 3, np.nan_to_num(muts)).astype(np.uint8)


@pytest.mark.parametrize("name", CASES)
def test_get_mut_df_bit_exact(name):
    g = load(name)
    vcf, tree, meta = case_inputs(g)
    m = O.get_mut_df(vcf, meta["exclude"], meta["include"])
    assert list(m["samples"]) == list(g["samples"])
    assert np.array_equal(codes(m["muts"]), g["muts_codes"])
    assert np.array_equal(np.array(m["pos"]).reshape(-1, 2), g["pos"])


def _mut_from_golden(g):
    c = g["muts_codes"].astype(float)
    c[c == 3] = np.nan
    return {"muts": c, "samples": g["samples"]}


@pytest.mark.parametrize("name", CASES)
def test_tree_clades_and_placement(name):
    g = load(name)
    vcf, tree, meta = case_inputs(g)
    r = O.get_gt_tree(tree, _mut_from_golden(g), float(g["w_maxs"][0]), meta["FN_in"], meta["FP_in"])
    assert r["FP"] == float(g["FP"]) and r["FN"] == float(g["FN"])
    assert list(r["columns"]) == list(g["red_cols"])
    assert np.array_equal(r["S"].astype(np.uint8), g["S"])              # clade masks: bit exact
    assert r["n_muts"] == g["M_index_pos"].shape[0]
    np.testing.assert_allclose(r["soft"], g["soft"], rtol=1e-9, atol=1e-12)
    names = [l.name for l in r["tree"].get_leaves()]
    assert names == list(g["MS_leaf_names"])
    ms = dict(zip(r["columns"], r["MS"]))
    np.testing.assert_allclose([ms[n] for n in names], g["MS_leaf"], rtol=0, atol=0)
    raw = np.array([n.weights[0] for n in r["tree"].iter_descendants()])
    np.testing.assert_allclose(raw, g["raw_weights"][0], rtol=1e-12)


@pytest.mark.parametrize("name", SYN)
def test_placement_literal_loop_matches_vectorised(name):
    g = load(name)
    vcf, tree, meta = case_inputs(g)
    r = O.get_gt_tree(tree, _mut_from_golden(g), float(g["w_maxs"][0]), meta["FN_in"], meta["FP_in"])
    soft_lit, M = O.map_mutations_gt_literal(r["muts_red"], r["S"], r["FP"], r["FN"])
    np.testing.assert_allclose(soft_lit, r["soft"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(M[:64], g["M_head"][:M.shape[0]], rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("name", CASES)
def test_model_data_and_lrt(name):
    g = load(name)
    vcf, tree, meta = case_inputs(g)
    mut = _mut_from_golden(g)
    ws = list(g["w_maxs"])
    if name.startswith("example") or name.startswith("syn_d"):
        ws = ws[:1] + ws[-1:]           # the SciPy solve takes seconds per w
    for w in ws:
        k = list(g["w_maxs"]).index(w)
        r = O.get_gt_tree(tree, mut, float(w), meta["FN_in"], meta["FP_in"])
        Y, constr, init, wn, cols = O.get_model_data(r["tree"])
        np.testing.assert_allclose(Y, g["Y"][k], rtol=1e-9, atol=1e-12)
        assert np.array_equal(constr, g["constr"][k])
        assert list(cols) == list(g["constr_cols"][k])
        np.testing.assert_allclose(init, g["init"][k], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(wn[:, 0], g["weights_norm"][k], rtol=1e-12)
        # same SciPy call on the fixture's own model data -> same optimiser path
        LR, dof, on_bound, p, x = O.get_LRT_poisson(g["Y"][k], g["constr"][k], g["init"][k], g["weights_norm"][k])
        assert dof == int(g["dof"][k]) and on_bound == int(g["on_bound"][k])
        np.testing.assert_allclose(LR, g["LR"][k], rtol=1e-6)
        np.testing.assert_allclose(p, g["p_val"][k], rtol=1e-5)


@pytest.mark.parametrize("name", SYN)
def test_tsv_text(name):
    g = load(name)
    vcf, tree, meta = case_inputs(g)
    tsv, res = O.run_poisson_tree_test(vcf, tree, [float(w) for w in g["w_maxs"]], meta["exclude"], meta["include"],
                                       meta["FN_in"], meta["FP_in"])
    ref = str(g["tsv"]).split("\n")
    got = tsv.split("\n")
    assert got[0] == ref[0]
    rc, gc = ref[1].split("\t"), got[1].split("\t")
    assert rc[:3] == gc[:3]
    for i in range(3, len(rc), 4):
        assert abs(float(rc[i]) - float(gc[i])) <= 2e-3 * max(1.0, abs(float(rc[i])))
        assert rc[i + 1] == gc[i + 1] and rc[i + 3] == gc[i + 3]
        assert abs(float(rc[i + 2]) - float(gc[i + 2])) <= 1e-3 * float(rc[i + 2])


def test_pvalue_mixture_known_values():
    from scipy.stats.distributions import chi2
    assert O.lrt_pvalue(30.0, 29, 0) == pytest.approx(chi2.sf(30.0, 29))
    want = 0.5 * chi2.sf(30.0, 29) + 0.5 * chi2.sf(30.0, 28)
    assert O.lrt_pvalue(30.0, 29, 1) == pytest.approx(want)
    # dof - k <= 0 terms are NaN in SciPy and get weight 0
    assert np.isfinite(O.lrt_pvalue(3.0, 2, 3))


def test_interval_form_of_the_placement_oracle():
    """The oracle's prefix-count form of map_mutations_gt (for tree-shaped S) against its dense form and
    against the soft weights the reference recorded; its cell order against the product's host analysis."""
    from scsommerclock_b200 import engine, synth
    for name in CASES:
        g = load(name)
        S = g["S"].astype(float)
        codes = g["muts_codes"]
        samples = list(g["samples"])
        if codes.shape[1] != S.shape[1]:
            codes = codes[:, [i for i, c in enumerate(samples) if c != "healthycell"]]
        muts = np.where(codes == 3, np.nan, codes.astype(float))
        if "healthycell" in samples:
            muts = muts[((codes == 1) | (codes == 2)).any(axis=1)]
        FP, FN = float(g["FP"]), float(g["FN"])
        got = O.map_mutations_gt_interval(muts, S, FP, FN)
        np.testing.assert_allclose(got, g["soft"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(got, O.map_mutations_gt(muts, S, FP, FN), rtol=1e-9, atol=1e-9)
    rng = np.random.default_rng(8)
    for n in (2, 33, 300):
        tree = synth.coalescent_tree(n, rng)
        cols = rng.permutation(n)
        S = np.ascontiguousarray(tree.clade_matrix()[:, cols])
        codes = np.ascontiguousarray(synth.observe(tree, synth.throw_mutations(tree, 150, rng), 0.1, 0.01, 0.2, rng)[:, cols])
        muts = np.where(codes == 3, np.nan, codes.astype(float))
        np.testing.assert_allclose(O.map_mutations_gt_interval(muts, S.astype(float), 0.01, 0.1),
                                   O.map_mutations_gt(muts, S.astype(float), 0.01, 0.1), rtol=1e-9, atol=1e-9)
        perm, lo, hi = O.clade_intervals(S)
        bits, words = engine.pack_clades(S)
        ok, perm2, lo2, hi2 = engine.clade_intervals(bits, n)
        assert ok
        for b in range(S.shape[0]):                                   # both orders describe the same clades
            assert set(perm[lo[b]:hi[b]].tolist()) == set(perm2[lo2[b]:hi2[b]].tolist()) == set(np.nonzero(S[b])[0].tolist())
    bad = np.array([[1, 1, 0, 0], [0, 1, 1, 0]], dtype=float)
    assert O.clade_intervals(bad) is None


def percell_case(pc, name):
    """What the reference's map_mutations_gt was given in one case of percell_fn.npz: the filtered {0,1,2,NaN} matrix, S,
    FP and the per-cell rates."""
    codes = pc[name + "_codes"]
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    return muts, pc[name + "_S"].astype(float), float(pc[name + "_FP"]), pc[name + "_FN_cells"]


def test_per_cell_fn_against_the_reference():
    """The pd.Series branch of map_mutations_gt (run_PT_test.py:413-418): the oracle's dense and prefix-sum forms against
    what the unmodified reference returned for random per-cell rates (percell_fn.npz, make_golden.make_percell_fn)."""
    pc = np.load(golden_path("percell_fn"), allow_pickle=False)
    for name in pc["cases"]:
        name = str(name)
        muts, S, FP, rates = percell_case(pc, name)
        assert muts.shape[0] == int(pc[name + "_n_muts"])
        soft, M = O.map_mutations_gt(muts, S, FP, rates, return_M=True)
        np.testing.assert_allclose(soft, pc[name + "_soft"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(M[:32], pc[name + "_M_head"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(O.map_mutations_gt_interval(muts, S, FP, rates), pc[name + "_soft"], rtol=1e-9, atol=1e-9)
        # equal rates: the float branch
        same = np.full(S.shape[1], 0.2)
        np.testing.assert_allclose(O.map_mutations_gt(muts, S, FP, same), O.map_mutations_gt(muts, S, FP, 0.2), rtol=1e-12, atol=1e-12)


def test_reference_copy_reproduces_a_fixture():
    """oracle/_ref (the byte-for-byte copy of the reference that bench.py's CPU arms run) gives the fixture's TSV and soft
    weights -- the arm times the same module that made the golden vectors."""
    from oracle import ref_module
    if ref_module.make() is None:
        pytest.skip("oracle/_ref is made where /root/reference exists")
    try:
        ref = ref_module.load()
        g = load("syn_a")
        vcf, tree, meta = case_inputs(g)
        call_data = ref.get_mut_df(vcf, meta["exclude"], meta["include"])
        t, FP, FN, M = ref.get_gt_tree(tree, call_data, float(g["w_maxs"][0]))
        np.testing.assert_allclose(M.values.sum(axis=0), g["soft"], rtol=1e-12)
        Y, constr, init, wn, cols = ref.get_model_data(t)
        LR, dof, on_bound, p_val, _ = ref.get_LRT_poisson(Y, constr, init, wn[:, 0])
        assert dof == int(g["dof"][0]) and on_bound == int(g["on_bound"][0])
        assert abs(LR - g["LR"][0]) <= 1e-6 * max(1.0, abs(g["LR"][0]))
    finally:
        np.seterr(all="warn")          # the reference sets np.seterr(all='raise') process-wide on import
