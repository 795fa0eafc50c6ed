#!/usr/bin/env python3
"""Generate the golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports /root/reference/run_PT_test.py as a module and calls its own
functions on (a) the reference's example_data and (b) small synthetic VCF/tree
pairs written to tests/golden/data/ that exercise the decoder's edge cases.
Inputs, intermediates and outputs are stored in tests/golden/*.npz.

Two environment notes, repeated in every fixture's ``meta``:
  * ete3 cannot be installed here (no network, not in the wheelhouse); the
    reference's ``from ete3 import Tree`` is served by
    ``scsommerclock_b200.tree.Tree`` (a restatement of the ete3 3.1.x TreeNode
    methods the reference calls).
  * the reference's global ``np.seterr(all='raise')`` makes SciPy 1.18's
    trust-constr raise FloatingPointError on harmless internal underflow, so
    ``scipy.optimize.minimize`` is wrapped in ``np.errstate(under='ignore')``.
    Nothing else of the reference is touched.
"""
import gzip
import importlib.util
import json
import os
import shutil
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
warnings.filterwarnings("ignore")

META = {
    "reference": "cbg-ethz/scSomMerClock run_PT_test.py (unmodified, imported from /root/reference)",
    "ete3": "absent; served by scsommerclock_b200.tree.Tree (ete3 3.1.x TreeNode semantics restated)",
    "numpy_errstate": "scipy.optimize.minimize wrapped in np.errstate(under='ignore')",
}


def load_reference():
    from scsommerclock_b200 import tree as T
    ete3 = types.ModuleType("ete3")
    ete3.Tree = T.Tree
    parser = types.ModuleType("ete3.parser")
    newick = types.ModuleType("ete3.parser.newick")
    newick.NewickError = T.NewickError
    parser.newick = newick
    ete3.parser = parser
    sys.modules.update({"ete3": ete3, "ete3.parser": parser, "ete3.parser.newick": newick})
    spec = importlib.util.spec_from_file_location("ref_run_PT_test", os.path.join(REF, "run_PT_test.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    _min = ref.minimize

    def minimize_no_underflow_trap(*a, **k):
        with np.errstate(under="ignore"):
            return _min(*a, **k)

    ref.minimize = minimize_no_underflow_trap
    import scipy, pandas
    META.update(numpy=np.__version__, scipy=scipy.__version__, pandas=pandas.__version__)
    return ref


def codes_of(df):
    v = df.values.astype(float)
    return np.where(np.isnan(v), 3, np.nan_to_num(v)).astype(np.uint8)


def run_case(ref, name, vcf, tree_file, w_maxs, exclude="", include="", FN_in=None, FP_in=None, rel_to=REPO):
    out = {}
    call_data = ref.get_mut_df(vcf, exclude, include)
    muts = call_data[0]
    out["muts_codes"] = codes_of(muts)
    out["samples"] = np.array(list(muts.columns))
    out["pos"] = np.array([list(p) for p in muts.index.values], dtype=np.int64).reshape(-1, 2)
    per_w = []
    tsv_header = "muts\tFN\tFP"
    tsv_model = ""
    for wi, w_max in enumerate(w_maxs):
        tree, FP, FN, M = ref.get_gt_tree(tree_file, call_data, w_max, FN_in, FP_in)
        nodes = list(tree.iter_descendants())
        if wi == 0:
            cols = [c for c in M.index.names] and list(muts.columns)
            red_cols = [c for c in muts.columns if c != "healthycell"]
            S = np.zeros((len(M.columns), len(red_cols)), dtype=np.uint8)
            for i, node in enumerate(nodes):
                for leaf in node.name.split("+"):
                    S[i, red_cols.index(leaf)] = 1
            out["S"] = S
            out["red_cols"] = np.array(red_cols)
            out["M_index_pos"] = np.array([list(p) for p in M.index.values], dtype=np.int64).reshape(-1, 2)
            out["soft"] = M.values.sum(axis=0)
            out["M_head"] = M.values[:64].copy()
            out["FP"], out["FN"] = FP, FN
            out["MS_leaf_names"] = np.array([l.name for l in tree.get_leaves()])
            out["MS_leaf"] = np.array([l.missing for l in tree.get_leaves()])
            out["node_names"] = np.array([n.name for n in nodes])
        Y, constr, init, weights_norm, constr_cols = ref.get_model_data(tree)
        LR, dof, on_bound, p_val, opt_vals = ref.get_LRT_poisson(Y, constr, init, weights_norm[:, 0])
        x_opt = np.array([v[0] for v in opt_vals])
        hyp = int(p_val < 0.05)
        tsv_header += "\t" + "\t".join([f"{i}_{w_max:.0f}" for i in ["-2logLR", "dof", "p-value", "hypothesis"]])
        tsv_model += f"\t{LR:0>5.3f}\t{dof}\t{p_val}\tH{hyp}"
        raw_w = np.array([n.weights[0] for n in nodes])
        per_w.append(dict(w_max=w_max, Y=Y, constr=constr, init=init, weights_norm=weights_norm[:, 0],
                          constr_cols=constr_cols, LR=LR, dof=dof, on_bound=on_bound, p_val=p_val, x_opt=x_opt,
                          raw_weights=raw_w, hyp=hyp))
        print(f"  {name} w={w_max}: LR={LR:.4f} dof={dof} on_bound={on_bound} p={p_val:.6g} H{hyp}", flush=True)
    tsv = f"{tsv_header}\n{M.shape[0]}\t{FN:.4f}\t{FP:.4f}{tsv_model}"
    # also let the reference write its own file through its public entry point, to pin the TSV text
    flat = {k: v for k, v in out.items()}
    flat["w_maxs"] = np.array(w_maxs, dtype=float)
    for k in ("Y", "init", "weights_norm", "x_opt", "raw_weights"):
        flat[k] = np.stack([d[k] for d in per_w])
    flat["constr"] = np.stack([d["constr"] for d in per_w])
    flat["constr_cols"] = np.stack([d["constr_cols"] for d in per_w])
    flat["LR"] = np.array([d["LR"] for d in per_w])
    flat["dof"] = np.array([d["dof"] for d in per_w])
    flat["on_bound"] = np.array([d["on_bound"] for d in per_w])
    flat["p_val"] = np.array([d["p_val"] for d in per_w])
    flat["hyp"] = np.array([d["hyp"] for d in per_w])
    flat["tsv"] = np.array(tsv)
    meta = dict(META, case=name, vcf=os.path.relpath(vcf, rel_to), tree=os.path.relpath(tree_file, rel_to),
                exclude=exclude, include=include, FN_in=FN_in, FP_in=FP_in)
    flat["meta"] = np.array(json.dumps(meta))
    # the Newick input travels with the fixture (the GPU box has no /root/reference)
    flat["tree_text"] = np.array(open(tree_file).read())
    flat["tree_basename"] = np.array(os.path.basename(tree_file))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **flat)


# ------------------------------------------------------- synthetic inputs
def write_vcf(path, names, rows, with_tg=True, reads="RC", chrom_prefix="", sep="|"):
    """rows: list of dicts(pos, ref, alts, filter, gts [list of str], tgs [list of str])."""
    fmt = "GT:DP:%s" % reads + (":TG" if with_tg else "")
    with gzip.open(path, "wt") if path.endswith(".gz") else open(path, "w") as f:
        f.write("##fileformat=VCFv4.3\n##source=scsommerclock_b200 synthetic golden input\n")
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        for r in rows:
            recs = []
            for k, gt in enumerate(r["gts"]):
                gt = gt.replace("|", sep)
                rd = "5,0,3,0" if reads == "RC" else "5,3"
                rec = f"{gt}:8:{rd}"
                if with_tg:
                    rec += ":" + r["tgs"][k]
                recs.append(rec)
            f.write(f"{chrom_prefix}1\t{r['pos']}\tsnv{r['pos']:06d}\t{r['ref']}\t{r['alts']}\t.\t{r['filter']}\t"
                    f"NS={len(names)}\t{fmt}\t" + "\t".join(recs) + "\n")
            if r.get("blank_after"):
                f.write("\n")


def make_synthetic(tag, n_cells, n_snv, seed, FN, FP, MS, with_tg=True, reads="RC", chrom_prefix="", sep="|",
                   name_style="cell", quirks=True, log_rates=None, tree_suffix=".raxml.bestTree", rate_mult_clade=None):
    from scsommerclock_b200 import synth
    rng = np.random.default_rng(seed)
    tree = synth.coalescent_tree(n_cells, rng)
    rate = None
    if rate_mult_clade is not None:
        # amplify the rate in the largest proper subtree (a clock violation)
        sizes = tree.hi - tree.lo
        cand = np.argsort(-sizes)
        top = [v for v in cand if sizes[v] <= n_cells // 2][0]
        rate = np.ones(2 * n_cells - 1)
        inside = (tree.lo >= tree.lo[top]) & (tree.hi <= tree.hi[top])
        rate[inside] = rate_mult_clade
    br = synth.throw_mutations(tree, n_snv, rng, rate)
    codes = synth.observe(tree, br, FN, FP, MS, rng)
    if name_style == "cell":
        names = ["cell%04d" % (i + 1) for i in range(n_cells)] + ["outgcell"]
    else:
        names = ["tumcell%04d" % (i + 1) for i in range(n_cells)] + ["healthycell"]
    bases = "ACGT"
    rows = []
    truth = (tree.leaf_rank[None, :] >= tree.lo[br][:, None]) & (tree.leaf_rank[None, :] < tree.hi[br][:, None])
    for i in range(n_snv):
        ref_b = bases[rng.integers(4)]
        alts = [b for b in bases if b != ref_b]
        alt_i = 1 + int(rng.integers(3))
        if reads == "AD":                      # Monovar style: the reference needs a single ALT with AD
            alts, alt_i = alts[:1], 1
        gts, tgs = [], []
        for c in range(n_cells):
            code = codes[i, c]
            if code == 3:
                gts.append(".|.")
            elif code == 0:
                gts.append("0|0")
            else:
                gts.append(f"{alt_i}|{alt_i}" if rng.random() < 0.15 else f"0|{alt_i}")
            tgs.append(f"0|{alt_i}" if truth[i, c] else "0|0")
        gts.append("0|0" if rng.random() > 0.02 else ".|.")   # outgroup
        tgs.append("0|0")
        n_called = sum(g not in ("0|0", ".|.") for g in gts)
        flt = "PASS"
        if n_called == 0:
            flt = "wildtype"
        elif n_called == 1 and rng.random() < 0.5:
            flt = "singleton"
        elif rng.random() < 0.03:
            flt = "s50"
        elif rng.random() < 0.02:
            flt = "LowQual"
        rows.append(dict(pos=3 * i + 1, ref=ref_b, alts=",".join(alts), filter=flt, gts=gts, tgs=tgs))
    if quirks:
        n = len(names)
        base = 3 * n_snv + 10

        def row(pos, gts, tgs=None, flt="PASS"):
            return dict(pos=pos, ref="A", alts="C,G,T", filter=flt, gts=gts, tgs=tgs or ["0|0"] * n)
        g = ["0|0"] * n
        # two different single alts -> ISA violation (dropped by the PASS filter)
        q = list(g); q[0] = "0|1"; q[1] = "0|2"; rows.append(row(base, q))
        # two alts, one of them twice -> majority alt kept, the other cell set to 0
        q = list(g); q[0] = "0|1"; q[1] = "0|1"; q[2] = "0|3"; q[3] = ".|."; rows.append(row(base + 1, q))
        # tie between two alts with counts (2,2): first (smaller) alt wins
        q = list(g); q[0] = "0|2"; q[1] = "2|2"; q[2] = "0|3"; q[3] = "3|3"; rows.append(row(base + 2, q))
        # '1|0' : the reference reads gt[-1] == '0' as the alt index
        q = list(g); q[4] = "1|0"; q[5] = "0|1"; q[6] = "0|1"; rows.append(row(base + 3, q))
        # only a true genotype, nothing called -> kept row of zeros (dropped later by sum > 0)
        t = ["0|0"] * n; t[2] = "0|1"; rows.append(row(base + 4, list(g), t))
        # nothing called, nothing true -> skipped entirely
        rows.append(row(base + 5, list(g)))
        # mutation only in the outgroup -> removed with the outgroup column
        q = list(g); q[-1] = "0|1"; rows.append(row(base + 6, q))
        # all missing except one hom call; filter text containing PASS as a substring
        q = [".|."] * n; q[7 % (n - 1)] = "2|2"; rows.append(row(base + 7, q, flt="PASS;extra"))
        # homozygous true genotype forms
        t = ["0|0"] * n; t[1] = "1|1"; q = list(g); q[1] = "1|1"; r_ = row(base + 8, q, t); r_["blank_after"] = True; rows.append(r_)
        q = list(g); q[3] = "0|2"; q[8 % (n - 1)] = "0|2"; rows.append(row(base + 9, q, flt="q20"))
    d = os.path.join(HERE, "data")
    os.makedirs(d, exist_ok=True)
    vcf = os.path.join(d, tag + ".vcf.gz")
    write_vcf(vcf, names, rows, with_tg=with_tg, reads=reads, chrom_prefix=chrom_prefix, sep=sep)
    tree_file = os.path.join(d, tag + tree_suffix)
    with open(tree_file, "w") as f:
        f.write(tree.newick(names[:-1], names[-1], scale=0.05) + "\n")
    if log_rates is not None:
        with open(os.path.join(d, tag + ".raxml.log"), "w") as f:
            f.write("CellPhy-style log (synthetic)\n   SEQ_ERROR: %s,  ADO_RATE: %s\n" % log_rates)
    return vcf, tree_file


def scite_newick(tree, n_cells):
    """infSCITE-style Newick (run_PT_test.py:233-252): every node is an integer, leaves 0..n_cells-1 are the
    tumour cells in VCF column order, n_cells is the outgroup, internal nodes carry the larger labels."""
    counter = [n_cells + 1]

    def rec(v):
        if v < n_cells:
            return str(v)
        a, b = tree.left[v - n_cells], tree.right[v - n_cells]
        body = "(%s,%s)" % (rec(a), rec(b))
        lab = counter[0]
        counter[0] += 1
        return body + str(lab)

    body = rec(2 * n_cells - 2)
    return "(%s,%d)%d;" % (body, n_cells, counter[0])


def make_scite_and_path_cases(ref):
    """The two tree-file branches of read_tree the CellPhy cases do not reach: infSCITE trees (:233-280; integer
    node labels, error rates from the .log or the .errors.csv next to the tree, FN doubled) and plain Newick files
    whose error rates are parsed out of the file PATH (:281-294)."""
    from scsommerclock_b200 import synth
    d = os.path.join(HERE, "data")
    n_cells, n_snv = 14, 300
    v, t = make_synthetic("syn_f", n_cells, n_snv, 31, FN=0.15, FP=0.01, MS=0.1, quirks=False, tree_suffix=".plain.newick")
    rng = np.random.default_rng(31)
    tree = synth.coalescent_tree(n_cells, rng)          # the same tree make_synthetic drew (same seed, first draw)
    os.remove(t)
    # (1) infSCITE tree + .log holding "best value for beta:\\t..." (a literal backslash-t: that is what the reference's pattern matches)
    t1 = os.path.join(d, "syn_f_ml0.newick")
    with open(t1, "w") as f:
        f.write(scite_newick(tree, n_cells) + "\n")
    with open(os.path.join(d, "syn_f.log"), "w") as f:
        f.write("infSCITE-style log (synthetic)\nbest value for beta:\\t0.0750\nbest value for alpha:\\t1.2e-2\n")
    run_case(ref, "syn_f_scite_log", v, t1, [100, 1000])
    # (2) infSCITE tree + .errors.csv, two trees in the file (everything after the first ';' is dropped, :237-238)
    t2 = os.path.join(d, "syn_g_ml0.newick")
    with open(t2, "w") as f:
        f.write(scite_newick(tree, n_cells) + "\n" + scite_newick(tree, n_cells) + "\n")
    with open(os.path.join(d, "syn_g.errors.csv"), "w") as f:
        f.write("FP,FN\n0.02,0.06\n")
    run_case(ref, "syn_g_scite_csv", v, t2, [400])
    # (3) plain Newick, rates from the path (FN = 0.2 -> 0.1 after :346, FP = 0.01)
    sub = os.path.join(d, "sim_WGA0.2-0-0.01")
    os.makedirs(sub, exist_ok=True)
    names = ["cell%04d" % (i + 1) for i in range(n_cells)]
    t3 = os.path.join(sub, "syn_h.newick")
    with open(t3, "w") as f:
        f.write(tree.newick(names, "outgcell", scale=0.05) + "\n")
    run_case(ref, "syn_h_path_rates", v, t3, [100, 1000])


def make_lrt_sweep(ref, n_problems=240, seed=2024):
    """Random clock-test problems solved by the reference's own get_LRT_poisson (:638-700, SciPy trust-constr),
    set up by its own get_model_data (:518-635) on random trees of 3..60 leaves: clock-like and rate-shifted
    counts, sparse counts (zero-count branches, many fitted lengths on the bound), non-integer soft counts,
    weight vectors clipped at different w_max.  One npz: per problem Y, weights, constr (as child_int, the
    tree form, checked here to reproduce the dense matrix), init and the reference's LR / dof / on_bound / p."""
    import tempfile
    from scsommerclock_b200 import synth
    from scsommerclock_b200.run_PT_test import ClockConstraints
    rng = np.random.default_rng(seed)
    tmp = tempfile.mkdtemp(prefix="scs_lrt_")
    recs = []
    n_raise = 0
    k = 0
    while len(recs) < n_problems:
        k += 1
        n = int(rng.choice([3, 4, 5, 6, 8, 10, 14, 20, 30, 40, 50, 60]))
        st = synth.coalescent_tree(n, rng)
        names = ["tumcell%04d" % (i + 1) for i in range(n)]
        tf = os.path.join(tmp, "t%d.newick" % k)
        with open(tf, "w") as f:
            f.write(st.newick(names, "healthycell", scale=0.05))
        tree, _, _, _ = ref.read_tree(tf, np.array(names + ["healthycell"]))
        nodes = list(tree.iter_descendants())
        kind = ["clock", "shift", "sparse", "verysparse", "frac"][int(rng.integers(5))]
        mu = {"clock": 400.0, "shift": 400.0, "sparse": 6.0, "verysparse": 1.5, "frac": 60.0}[kind] * n
        w_max = float(rng.choice([50, 100, 400, 1000]))
        # per node: expected count from the node's height difference in the synthetic tree (clade size identifies the node)
        by_name = {}
        inv = np.empty(n, dtype=np.int64)
        inv[st.leaf_rank] = np.arange(n)
        for v_ in range(2 * n - 2):
            cells = sorted(names[c] for c in inv[st.lo[v_]:st.hi[v_]])
            by_name["+".join(cells)] = v_
        mult = np.ones(2 * n - 1)
        if kind == "shift":
            sizes = st.hi - st.lo
            top = [v_ for v_ in np.argsort(-sizes) if sizes[v_] <= max(1, n // 2)][0]
            inside = (st.lo >= st.lo[top]) & (st.hi <= st.hi[top])
            mult[inside] = float(rng.choice([3.0, 5.0, 8.0]))
        total_len = (st.branch_len[:-1] * mult[:-1]).sum()
        p_ado = np.exp(rng.uniform(np.log(1e-6), np.log(0.5), size=len(nodes)))
        raw = np.clip(1.0 / ((1 - p_ado) * p_ado), None, w_max)
        for i, node in enumerate(nodes):
            v_ = by_name.get(node.name)
            lam = 0.0 if v_ is None else mu * st.branch_len[v_] * mult[v_] / total_len
            y = float(rng.poisson(lam))
            if kind == "frac" or rng.random() < 0.3:
                y = y * float(rng.uniform(0.7, 1.3)) + float(rng.uniform(0, 0.4))
            node.mut_no_soft = y
        root_row = [i for i, node in enumerate(nodes) if node.name.count("+") == n - 1][0]
        wsel = np.delete(raw, root_row)
        wn = wsel.size * wsel / wsel.sum()
        j = 0
        for i, node in enumerate(nodes):
            if i == root_row:
                node.weights_norm = np.full(2, -1.0)
            else:
                node.weights_norm = np.array([wn[j], wn[j]])
                j += 1
        try:
            Y, constr, init, weights_norm, cols = ref.get_model_data(tree)
            LR, dof, on_bound, p_val, opt = ref.get_LRT_poisson(Y, constr, init, weights_norm[:, 0])[:5]
        except (FloatingPointError, AssertionError, ZeroDivisionError) as exc:
            n_raise += 1
            continue
        failed = isinstance(LR, float) and np.isnan(LR)
        cc = ClockConstraints.from_dense(constr)
        assert np.array_equal(cc.toarray(), constr)
        w = weights_norm[:, 0]
        ll_H1 = float(np.sum((Y * np.log(np.where(Y > 0, Y, 1)) - Y) * w))
        x = np.full(Y.size, np.nan) if failed else np.array([v_[0] for v_ in opt])
        recs.append(dict(Y=Y, w=w, child_int=cc.child_int, init=init, x=x, LR=np.nan if failed else LR, dof=-1 if failed else dof,
                         on_bound=-1 if failed else on_bound, p_val=np.nan if failed else p_val, kind=kind, n=n, ll_H1=ll_H1))
        print("  lrt %3d: n=%2d %-10s w_max=%4.0f ll_H1=%9.2f LR=%s dof=%s on_bound=%s p=%s" %
              (len(recs), n, kind, w_max, ll_H1, recs[-1]["LR"], recs[-1]["dof"], recs[-1]["on_bound"], recs[-1]["p_val"]), flush=True)
    off = np.zeros(len(recs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([r["Y"].size for r in recs])
    meta = dict(META, case="lrt_sweep", seed=seed, raised_and_skipped=n_raise,
                note="problems where the reference itself raised under np.seterr(all='raise') are skipped; LR = NaN marks its failure tuple")
    np.savez_compressed(os.path.join(HERE, "lrt_sweep.npz"), off=off, Y=np.concatenate([r["Y"] for r in recs]),
                        w=np.concatenate([r["w"] for r in recs]), child_int=np.concatenate([r["child_int"] for r in recs]),
                        init=np.concatenate([r["init"] for r in recs]), x=np.concatenate([r["x"] for r in recs]),
                        LR=np.array([r["LR"] for r in recs]), dof=np.array([r["dof"] for r in recs]),
                        on_bound=np.array([r["on_bound"] for r in recs]), p_val=np.array([r["p_val"] for r in recs]),
                        kind=np.array([r["kind"] for r in recs]), n_leaves=np.array([r["n"] for r in recs]),
                        ll_H1=np.array([r["ll_H1"] for r in recs]), meta=np.array(json.dumps(meta)))
    shutil.rmtree(tmp, ignore_errors=True)


def make_percell_fn(ref, seed=77):
    """The pd.Series branch of the reference's map_mutations_gt (:413-418): one false-negative rate per cell.  Nothing on
    the reference's command-line path passes a Series (get_gt_tree hands over a float), so the branch is driven directly:
    the reference's own get_mut_df and read_tree, the outgroup / empty-row filter of get_gt_tree (:350-353), then ITS
    map_mutations_gt with random per-cell rates.  One file, percell_fn.npz, keys prefixed by the case name."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    d = os.path.join(HERE, "data")
    cases = [("syn_a", "syn_a.vcf.gz", "syn_a.raxml.bestTree"), ("syn_c", "syn_c.vcf.gz", "syn_c.newick"),
             ("syn_b", "syn_b.vcf.gz", "syn_b.raxml.bestTree"),
             ("example_clock", "data_simulated_clock.vcf.gz", "data_simulated_clock.raxml.bestTree")]
    out = {"cases": np.array([c[0] for c in cases])}
    for name, vcf, tree_file in cases:
        call_data = ref.get_mut_df(os.path.join(d, vcf), "", "")
        muts = call_data[0]
        tree, outg, FP, FN = ref.read_tree(os.path.join(d, tree_file), muts.columns.values)
        FP = max(FP, ref.LAMBDA_MIN)
        if outg in muts.columns:                                            # :350-353
            muts_red = muts.drop(outg, axis=1)
            muts_red = muts_red[muts_red.sum(axis=1) > 0]
        else:
            muts_red = muts
        lo, hi = (0.02, 0.45) if name != "syn_b" else (1e-4, 0.9)           # syn_b: rates spread over four decades
        rates = np.exp(rng.uniform(np.log(lo), np.log(hi), muts_red.shape[1]))
        FN_series = pd.Series(rates, index=muts_red.columns)
        M = ref.map_mutations_gt(tree, muts_red, FP, FN_series)
        out[name + "_vcf"] = np.array(os.path.join("tests", "golden", "data", vcf))
        out[name + "_tree"] = np.array(os.path.join("tests", "golden", "data", tree_file))
        out[name + "_cols"] = np.array(list(muts_red.columns))
        out[name + "_all_cols"] = np.array(list(muts.columns))
        out[name + "_codes"] = codes_of(muts_red)                           # what map_mutations_gt was given
        red_cols = list(muts_red.columns)
        S = np.zeros((len(M.columns), len(red_cols)), dtype=np.uint8)
        for i, node in enumerate(tree.iter_descendants()):
            for leaf in node.name.split("+"):
                S[i, red_cols.index(leaf)] = 1
        out[name + "_S"] = S
        out[name + "_FN_cells"] = rates
        out[name + "_FP"] = np.array(FP)
        out[name + "_soft"] = M.values.sum(axis=0)
        out[name + "_M_head"] = M.values[:32].copy()
        out[name + "_n_muts"] = np.array(M.shape[0])
        out[name + "_node_names"] = np.array(list(M.columns))
        print(f"  percell {name}: {M.shape[0]} SNVs x {M.shape[1]} rows, FN in [{rates.min():.4g}, {rates.max():.4g}]", flush=True)
    out["meta"] = np.array(json.dumps(dict(META, case="percell_fn", seed=seed)))
    np.savez_compressed(os.path.join(HERE, "percell_fn.npz"), **out)


def main():
    ref = load_reference()
    if "--only-percell" in sys.argv:
        make_percell_fn(ref)
        return
    if "--only-lrt-sweep" in sys.argv:
        make_lrt_sweep(ref)
        return
    if "--only-scite" in sys.argv:
        make_scite_and_path_cases(ref)
        return
    ex = os.path.join(REF, "example_data")
    default_w = [100, 200, 300, 400, 500, 600, 700, 800, 900, 1000]
    if "--skip-examples" not in sys.argv:
        # the reference's own example inputs travel with the fixtures (the GPU box has no /root/reference):
        # byte-for-byte copies under tests/golden/data/, and the reference is run on the copies
        d = os.path.join(HERE, "data")
        os.makedirs(d, exist_ok=True)
        for tag in ("clock", "noclock"):
            for ext in (".vcf.gz", ".raxml.bestTree"):
                shutil.copyfile(os.path.join(ex, "data_simulated_%s%s" % (tag, ext)), os.path.join(d, "data_simulated_%s%s" % (tag, ext)))
            run_case(ref, "example_" + tag, os.path.join(d, "data_simulated_%s.vcf.gz" % tag),
                     os.path.join(d, "data_simulated_%s.raxml.bestTree" % tag), default_w)
    if "--only-examples" in sys.argv:
        return
    v, t = make_synthetic("syn_a", 12, 240, 11, FN=0.12, FP=0.01, MS=0.1, log_rates=("0.010000", "0.240000"))
    run_case(ref, "syn_a", v, t, [100, 400, 1000])
    v, t = make_synthetic("syn_b", 20, 400, 12, FN=0.2, FP=0.005, MS=0.25, with_tg=False, reads="AD", chrom_prefix="chr",
                          sep="/", name_style="tumcell", rate_mult_clade=6.0, quirks=False)
    run_case(ref, "syn_b", v, t, [50, 1000], FN_in=0.2, FP_in=0.005)
    v, t = make_synthetic("syn_c", 9, 150, 13, FN=0.05, FP=0.02, MS=0.05, quirks=False, tree_suffix=".newick")
    run_case(ref, "syn_c", v, t, [1000], exclude="tumcell000[12]")
    run_case(ref, "syn_c_incl", v, t, [300], include="(tumcell000[1-7])|healthycell")
    # tree leaves that are not in the VCF: the reference deletes them (:311-313), collapsing their
    # parents.  (The opposite mismatch -- a VCF sample missing from the tree -- makes the reference
    # itself fail with a KeyError at :437 as soon as one of the padding rows of S wins an SNV.)
    v, t = make_synthetic("syn_e", 10, 160, 14, FN=0.1, FP=0.01, MS=0.1, quirks=False)
    from scsommerclock_b200.tree import Tree as _T
    tt = _T(open(t).read().strip(), format=1)
    leaves = tt.get_leaves()
    for k, (host, nm) in enumerate(((leaves[2], "cell0098"), (leaves[7], "cell0099"))):
        # graft an extra leaf next to an existing one: (host) -> (host, extra)
        up = host.up
        pos = up.children.index(host)
        joint = _T()
        joint.dist = host.dist / 2
        host.dist = host.dist / 2
        up.children[pos] = joint
        joint.up = up
        joint.children = [host, _T(name=nm, dist=0.02)] if k == 0 else [_T(name=nm, dist=0.02), host]
        for ch in joint.children:
            ch.up = joint
    t2 = t.replace("syn_e", "syn_e_mismatch")
    def _nw(node):
        if node.children:
            return "(" + ",".join(_nw(c) for c in node.children) + "):%.6f" % node.dist
        return "%s:%.6f" % (node.name, node.dist)
    with open(t2, "w") as f:
        f.write("(" + ",".join(_nw(c) for c in tt.children) + ");\n")
    os.remove(t)
    run_case(ref, "syn_e_mismatch", v, t2, [500])
    if "--skip-large" not in sys.argv:
        # example_data-sized synthetic pairs (30 cells + outgroup, ~4000 SNVs, default w_max list),
        # one simulated under the clock and one with a 5x rate in a subtree, like the reference's examples
        v, t = make_synthetic("syn_d_clock", 30, 4000, 21, FN=0.1, FP=0.005, MS=0.1, quirks=True, log_rates=("0.005000", "0.200000"))
        run_case(ref, "syn_d_clock", v, t, default_w)
        v, t = make_synthetic("syn_d_noclock", 30, 4000, 22, FN=0.1, FP=0.005, MS=0.1, quirks=False, log_rates=("0.005000", "0.200000"),
                              rate_mult_clade=5.0)
        run_case(ref, "syn_d_noclock", v, t, default_w)
    if "--skip-scite" not in sys.argv:
        make_scite_and_path_cases(ref)
    if "--skip-lrt-sweep" not in sys.argv:
        make_lrt_sweep(ref)
    make_percell_fn(ref)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
