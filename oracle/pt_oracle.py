"""CPU oracle for the Poisson Tree (PT) test hot path -- TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of the reference's algorithm (cbg-ethz/scSomMerClock,
run_PT_test.py); every function cites the reference lines it follows.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; the product (``scsommerclock_b200``) never does.

Pinning: parity is PINNED against the reference itself.  ``tests/golden/
make_golden.py`` imports the unmodified /root/reference/run_PT_test.py in the
build container and records its inputs, intermediates and outputs for
example_data and for small synthetic edge cases; ``tests/test_oracle_golden.py``
checks this oracle against those fixtures.  Two things differ from a stock
reference environment and are recorded in the fixtures' ``meta``: (1) ete3 is
not installable here, so the reference ran on ``scsommerclock_b200.tree.Tree``,
a restatement of the ete3 3.1.x TreeNode methods the reference calls; (2) NumPy
underflow trapping was disabled inside ``scipy.optimize.minimize`` only (the
reference's global ``np.seterr(all='raise')`` makes SciPy 1.18's own linear
algebra raise on harmless underflow).

The reference relies on ``np.seterr(all='raise')`` + ``except
FloatingPointError`` for control flow in two places (run_PT_test.py:370-377 and
:482-486); this oracle states those conditions explicitly instead of changing
the process-wide NumPy error state.
"""
import gzip
import os
import re

import numpy as np

LAMBDA_MIN = 1e-6          # run_PT_test.py:21
log_LAMBDA_MIN = -100      # run_PT_test.py:22
_EXP_UNDERFLOW = np.log(np.finfo(np.float64).tiny)   # exp(x) raises 'underflow' below this


def _tree_class():
    try:
        from ete3 import Tree  # the real thing, when present
        from ete3.parser.newick import NewickError
        return Tree, NewickError
    except Exception:
        from scsommerclock_b200.tree import Tree, NewickError
        return Tree, NewickError


# --------------------------------------------------------------------- VCF
def get_mut_df(vcf_file, exclude_pat="", include_pat=""):
    """run_PT_test.py:27-207 (filter=True).  Returns a dict:
    muts [n_kept, n_incl] float64 in {0,1,2,nan}; samples (included names);
    pos [(chrom, pos)] of kept rows; skipped = [no_alt, filtered] counters.
    Read counts (RC/AD) and true genotypes are parsed by the reference but not
    used by the PT test and are not reproduced (TG only feeds the skip rule)."""
    opener = gzip.open if vcf_file.endswith("gz") else open
    exclude = []
    sample_names = None
    rows, true_rows, pos_all, keep_flags = [], [], [], []
    skip = [0, 0]
    with opener(vcf_file, "rb") as f_in:
        for raw in f_in:
            line = raw.decode()
            if line.startswith("#"):
                if line.startswith("#CHROM"):                                   # :49-76
                    if line.count("tumcell") == 0:
                        line = re.sub(r"cell(?=\d+)", "tumcell", line)
                    if line.count("healthycell") == 0:
                        line = re.sub("outgcell", "healthycell", line)
                    sample_names = [i.strip() for i in line.split("\t")[9:]]
                    if exclude_pat != "":
                        for s in sample_names:
                            if re.fullmatch(exclude_pat, s):
                                exclude.append(s)
                    if include_pat != "":
                        for s in sample_names:
                            if not re.fullmatch(include_pat, s):
                                exclude.append(s)
                continue
            elif line.strip() == "":
                continue
            cols = line.strip().split("\t")                                     # :82
            n = len(sample_names)
            het = np.zeros(n)
            hom = np.zeros(n)
            tru = np.zeros(n)
            diff_muts = []
            elem_ids = cols[8].split(":")
            gt_i = elem_ids.index("GT")
            tg_i = elem_ids.index("TG") if "TG" in elem_ids else -1
            for s_i, s_rec in enumerate(cols[9:]):                              # :89-119
                els = s_rec.split(":")
                gt = els[gt_i].replace("/", "|")
                if tg_i >= 0:
                    true_gt = els[tg_i]
                else:
                    true_gt = ""
                    tru[s_i] = np.nan
                if true_gt.count("0") == 1:
                    tru[s_i] = 1
                elif true_gt in ("1|1", "2|2", "1|2", "2|1"):
                    tru[s_i] = 2
                if gt == ".|.":
                    het[s_i] = hom[s_i] = np.nan
                    continue
                elif gt == "0|0":
                    pass
                elif "0" in gt:
                    het[s_i] = int(gt[-1])
                    diff_muts.append(int(gt[-1]))
                else:
                    hom[s_i] = int(gt[-1])
                    diff_muts.append(int(gt[-1]))
            if not (np.nansum(het) > 0 or np.nansum(hom) > 0 or np.nansum(tru) > 0):   # :147
                skip[0] += 1
                continue
            alts, alt_counts = np.unique(diff_muts, return_counts=True)         # :152
            flt = cols[6]
            if alts.size == 0:                                                  # :155
                row = het + hom
            elif alts.size == 2 and all(alt_counts == 1):                       # :159
                row = np.where(np.isnan(het), np.nan, 0)
                flt = "ISA_violation"
            else:                                                               # :163
                ab_alt = alts[np.argmax(alt_counts)]
                h1 = np.where(het == ab_alt, 1, 0)
                h2 = np.where(hom == ab_alt, 2, 0)
                row = np.where(np.isnan(het), np.nan, h2 + h1)
            rows.append(row)
            pos_all.append((int(cols[0].replace("chr", "")), int(cols[1])))     # :172
            if "PASS" in flt or flt == "s50" or flt == "singleton":             # :175
                keep_flags.append(True)
            else:
                keep_flags.append(False)
                skip[1] += 1
    exclude.append("")                                                          # :187
    incl_idx = [i for i, s in enumerate(sample_names) if s not in exclude]
    keep = np.array(keep_flags, dtype=bool)
    kept_pos = [p for p, k in zip(pos_all, keep_flags) if k]
    if len(set(kept_pos)) != len(kept_pos):
        raise ValueError("duplicate (chrom, pos) among kept records: the reference's label-based "
                         ".loc selection (run_PT_test.py:192) would duplicate rows")
    data = np.array(rows, dtype=np.float64).reshape(len(rows), len(sample_names))
    muts = data[keep][:, incl_idx]
    return {"muts": muts, "samples": np.array([sample_names[i] for i in incl_idx]), "pos": kept_pos,
            "all_samples": sample_names, "skipped": skip}


# -------------------------------------------------------------------- tree
def read_tree(tree_file, samples=()):
    """run_PT_test.py:210-334.  Returns (tree, outgroup name, FP, FN)."""
    Tree, NewickError = _tree_class()
    samples = np.asarray(samples)
    with open(tree_file, "r") as f:
        tree_raw = f.read().strip()
    cellphy_ends = (".raxml.bestTree", "raxml.mutationMapTree", "raxml.supportFBP", "raxml.supportTBE")
    if tree_file.endswith(cellphy_ends) or "cellphy" in tree_file:             # :216-232
        tree_raw = re.sub(r"\[\d+\]", "", tree_raw)
        if tree_raw.count("tumcell") == 0:
            tree_raw = re.sub(r"cell(?=\d+)", "tumcell", tree_raw)
        if tree_raw.count("healthycell") == 0:
            tree_raw = re.sub("outgcell", "healthycell", tree_raw)
        log_file = ".".join(tree_file.split(".")[:-1]) + ".log"
        if os.path.exists(log_file):
            with open(log_file, "r") as f:
                log = f.read().strip()
            FP = float(re.search(r"SEQ_ERROR: (0.\d+(e-\d+)?)", log).group(1))
            FN = float(re.search(r"ADO_RATE: (0.\d+(e-\d+)?)", log).group(1))
        else:
            FP = LAMBDA_MIN
            FN = LAMBDA_MIN
    elif tree_file.endswith("_ml0.newick") or "scite" in tree_file:            # :233-280
        sem_count = tree_raw.count(";")
        if sem_count == 0:
            tree_raw += ";"
        elif sem_count > 1:
            tree_raw = tree_raw.split(";")[0].strip() + ";"
        nodes = [int(i) for i in re.findall(r"(?<=[\(\),])\d+(?=[,\)\(;])", tree_raw)]
        int_node_idx = 1
        for i, s_i in enumerate(sorted(nodes)):
            pat = r"(?<=[\(\),])%d(?=[,\)\(;)])" % s_i
            if i < samples.size:
                repl = "%s:1.0" % samples[i]
            else:
                repl = "Node%d:1.0" % int_node_idx
                int_node_idx += 1
            tree_raw = re.sub(pat, repl, tree_raw)
        tree_raw = tree_raw[tree_raw.index("("):]
        log_file = tree_file.replace("_ml0.newick", ".log")
        if log_file == tree_file:
            log_file = tree_file.replace(".newick", ".log")
        error_file = tree_file.replace("_ml0.newick", ".errors.csv")
        if os.path.exists(log_file):
            with open(log_file, "r") as f:
                log_raw = f.read()
            FN = float(re.search(r"best value for beta:\\t(\d.\d+(e-\d+)?)", log_raw).group(1))
            FP = float(re.search(r"best value for alpha:\\t(\d.\d+(e-\d+)?)", log_raw).group(1))
            FN *= 2
        elif os.path.exists(error_file):
            with open(error_file, "r") as f:
                errors_raw = f.read().strip().split("\n")
            FP, FN = [float(i) for i in errors_raw[1].split(",")]
            FN *= 2
        else:
            FP = LAMBDA_MIN
            FN = LAMBDA_MIN
    else:                                                                       # :281-294
        if tree_raw.count("tumcell") == 0:
            tree_raw = re.sub(r"cell(?=\d+)", "tumcell", tree_raw)
        if tree_raw.count("healthycell") == 0:
            tree_raw = re.sub("outgcell", "healthycell", tree_raw)
        try:
            FP = float(re.search(r"WGA0[\.\d,]*-0[\.\d]*-(0[\.\d]*)", tree_file).group(1))
            FN = float(re.search(r"WGA(0[\.\d]*)[,\.\d]*?-", tree_file).group(1))
        except AttributeError:
            FP = LAMBDA_MIN
            FN = LAMBDA_MIN
    outg_name = "healthycell"
    try:
        tree = Tree(tree_raw, format=2)
    except NewickError:
        tree = Tree(tree_raw, format=1)
    outg_node = tree & outg_name                                                # :301-308
    anc_node = outg_node.get_ancestors()[0]
    outg_node.support = 100 if anc_node.is_root() else anc_node.support
    tree.set_outgroup(outg_node)
    outg_node.delete()
    for node in tree.get_leaves():                                              # :311-313
        if node.name not in samples:
            node.delete()
    for node in tree.iter_descendants():                                        # :317-332
        node.add_features(in_length=node.dist, mut_no_soft=0, mut_no_true=-1,
                          name="+".join(sorted(node.get_leaf_names())),
                          weights=np.zeros(2), weights_norm=np.zeros(2))
    return tree, outg_name, FP, FN


def clade_matrix(tree, columns):
    """S of run_PT_test.py:389-397: one row per ``tree.iter_descendants()`` node
    (level order), padded with zero rows up to 2n-1, plus the final zero row."""
    n = len(columns)
    col = {c: i for i, c in enumerate(columns)}
    nodes = list(tree.iter_descendants())
    S = np.zeros((max(2 * n - 1, len(nodes)) + 1, n))
    for i, node in enumerate(nodes):
        for leaf in node.name.split("+"):
            S[i, col[leaf]] = 1
    return S, nodes


# --------------------------------------------------------------- placement
def normalize_log_probs_rows(probs):
    """_normalize_log_probs (run_PT_test.py:368-383, return_normal=True) applied
    to every row of ``probs``."""
    out = np.empty_like(probs)
    for i in range(probs.shape[0]):
        p = probs[i]
        max_i = np.nanargmax(p)
        others = np.arange(p.size) != max_i
        d = p[others] - p[max_i]
        if np.any(d < _EXP_UNDERFLOW):          # the reference's FloatingPointError branch (:373-377)
            e = np.exp(np.clip(d, log_LAMBDA_MIN, 0))
        else:
            e = np.exp(d)
        pn = p - p[max_i] - np.log1p(np.nansum(e))
        out[i] = np.exp(np.clip(pn, log_LAMBDA_MIN, 0))
    return out


def placement_log_probs(muts_vals, S, FP, FN):
    """probs of run_PT_test.py:408-427 for all SNVs at once.  ``FN``: a float (:408-412) or one rate per column of
    ``muts_vals`` / ``S`` (the pd.Series branch :413-418, ``FN_i = FN[muts_in.columns].values``: the same four
    products with log(1 - FN_i) / log(FN_i) broadcast over the columns of S)."""
    if not np.isscalar(FN):
        FN = np.asarray(FN, dtype=np.float64)
        assert FN.shape == (S.shape[1],), "one FN rate per cell"
    S_inv = 1 - S
    nans = np.isnan(muts_vals)
    muts = np.where(nans, np.nan, muts_vals.astype(bool).astype(float))        # :403
    mut1 = np.where(nans, 0.0, muts)
    inv1 = np.where(nans, 0.0, 1 - muts)
    TP_m = S * np.log(1 - FN)
    FP_m = S_inv * np.log(FP)
    TN_m = S_inv * np.log(1 - FP)
    FN_m = S * np.log(FN)
    return mut1 @ (TP_m + FP_m).T + inv1 @ (TN_m + FN_m).T


def map_mutations_gt(muts_vals, S, FP, FN, return_M=False, chunk=4096):
    """run_PT_test.py:386-448: soft_assigned = M.sum(axis=0) (one value per row of
    S, the last one being the 'FP' column)."""
    soft = np.zeros(S.shape[0])
    Ms = []
    for s in range(0, muts_vals.shape[0], chunk):
        probs = placement_log_probs(muts_vals[s:s + chunk], S, FP, FN)
        M = normalize_log_probs_rows(probs)
        soft += M.sum(axis=0)
        if return_M:
            Ms.append(M)
    if return_M:
        return soft, (np.vstack(Ms) if Ms else np.zeros((0, S.shape[0])))
    return soft


def map_mutations_gt_literal(muts_vals, S, FP, FN):
    """The per-SNV loop exactly as the reference writes it (:421-430); slow, used
    to cross-check the vectorised form and as the 'reference' CPU timing arm."""
    S_inv = 1 - S
    nans = np.isnan(muts_vals)
    muts = np.where(nans, np.nan, muts_vals.astype(bool).astype(float))
    muts_inv = 1 - muts
    TP_m = S * np.log(1 - FN)
    FP_m = S_inv * np.log(FP)
    TN_m = S_inv * np.log(1 - FP)
    FN_m = S * np.log(FN)
    M = np.zeros((muts.shape[0], S.shape[0]))
    for i, mut in enumerate(muts):
        probs = np.stack([np.nansum(TP_m * mut, axis=1), np.nansum(FP_m * mut, axis=1),
                          np.nansum(TN_m * muts_inv[i], axis=1), np.nansum(FN_m * muts_inv[i], axis=1)]).sum(axis=0)
        M[i] = normalize_log_probs_rows(probs[None, :])[0]
    return M.sum(axis=0), M


def clade_intervals(S):
    """Order of the cells in which every row of a tree-shaped (laminar) 0/1 matrix S is one interval.
    Rows by decreasing size, cells sorted by their membership signature over those rows; returns
    (perm, lo, hi) with row b == cells perm[lo[b]:hi[b]], or None when some row is not contiguous in
    that order (S is then not a laminar family).  Independent NumPy restatement of the product's host
    analysis (csrc/scs_place_interval.cuh: clade_intervals), used to check it and to give the tests an
    oracle that finishes at 10,000 cells."""
    S = np.asarray(S).astype(bool)
    size = S.sum(axis=1)
    order = np.argsort(-size, kind="stable")
    order = order[size[order] > 0]
    # lexsort: last key is the primary one -> feed the rows smallest first; descending on every key
    perm = np.lexsort(tuple(~S[r] for r in order[::-1])) if order.size else np.arange(S.shape[1])
    pos = np.empty(S.shape[1], dtype=np.int64)
    pos[perm] = np.arange(S.shape[1])
    lo = np.zeros(S.shape[0], dtype=np.int64)
    hi = np.zeros(S.shape[0], dtype=np.int64)
    for b in range(S.shape[0]):
        if size[b] == 0:
            continue
        p = pos[S[b]]
        lo[b], hi[b] = p.min(), p.max() + 1
        if hi[b] - lo[b] != size[b]:
            return None
    return perm, lo, hi


def map_mutations_gt_interval(muts_vals, S, FP, FN, chunk=2048):
    """map_mutations_gt (run_PT_test.py:386-448) for a tree-shaped S without the dense products: the
    terms of probs[i, b] that depend on b are log((1-FN)/FP) * C1[i, b] + log(FN/(1-FP)) * C0[i, b]
    (C1 / C0: cells of clade b observed mutated / wild type; everything else cancels in the
    normalisation :368-383), and for a laminar S both counts are differences of per-SNV prefix counts
    in the cell order of clade_intervals.  float64 throughout; agrees with map_mutations_gt to
    rounding (tests/test_oracle_golden.py) and runs at 10,000 cells x 20,000 rows in seconds."""
    iv = clade_intervals(S)
    if iv is None:
        raise ValueError("S is not a laminar family")
    perm, lo, hi = iv
    a1 = np.log(1 - FN) - np.log(FP)
    a0 = np.log(FN) - np.log(1 - FP)
    soft = np.zeros(S.shape[0])
    percell = not np.isscalar(FN)          # the pd.Series branch (:413-418): prefix sums of per-cell weights instead of counts
    for s in range(0, muts_vals.shape[0], chunk):
        v = muts_vals[s:s + chunk][:, perm]
        obs = ~np.isnan(v)
        mut = obs & (np.nan_to_num(v) != 0)
        wt = obs & ~mut
        if percell:
            z = np.zeros((v.shape[0], 1))
            pw = np.concatenate([z, np.cumsum(mut * a1[perm][None, :] + wt * a0[perm][None, :], axis=1)], axis=1)
            logit = pw[:, hi] - pw[:, lo]
            M = np.exp(logit - logit.max(axis=1, keepdims=True))
            soft += (M / M.sum(axis=1, keepdims=True)).sum(axis=0)
            continue
        z = np.zeros((v.shape[0], 1), dtype=np.int64)
        p1 = np.concatenate([z, np.cumsum(mut, axis=1)], axis=1)
        p0 = np.concatenate([z, np.cumsum(wt, axis=1)], axis=1)
        logit = a1 * (p1[:, hi] - p1[:, lo]) + a0 * (p0[:, hi] - p0[:, lo])
        M = np.exp(logit - logit.max(axis=1, keepdims=True))
        soft += (M / M.sum(axis=1, keepdims=True)).sum(axis=0)
    return soft


# ----------------------------------------------------------------- weights
def branch_weights(S_nodes_bool, l_TN, l_DO, w_max):
    """add_br_weights, run_PT_test.py:451-502 (per-cell MS branch).  S_nodes_bool
    [n_nodes, m] over the tree's leaves in sorted-name order.  Returns (weights
    [n_nodes-1, 2], weights_norm, root_id)."""
    n, m = S_nodes_bool.shape
    p_ADO = np.zeros(n)
    for i, br in enumerate(S_nodes_bool):
        x = l_DO[br].sum() + l_TN[~br].sum()
        p_ADO[i] = LAMBDA_MIN if x < _EXP_UNDERFLOW else max(np.exp(x), LAMBDA_MIN)   # :482-486
    p_noADO = 1 - p_ADO
    weights = np.zeros((n, 2))
    weights[:, 0] = np.clip(1 / (p_noADO * p_ADO), None, w_max)
    weights[:, 1] = np.clip(p_noADO / p_ADO, None, w_max)
    root_id = np.argwhere(S_nodes_bool.sum(axis=1) == m).flatten()[0]
    weights = np.delete(weights, root_id, axis=0)
    weights_norm = weights.shape[0] * weights / weights.sum(axis=0)
    return weights, weights_norm, root_id


def add_br_weights(tree, FP, FN, MS_by_name, w_max):
    """run_PT_test.py:451-515 on a tree: sets node.weights / node.weights_norm."""
    m = len(tree)
    nodes = list(tree.iter_descendants())
    leaf_names = sorted(i.name for i in tree.get_leaves())
    leaf_map = {j: i for i, j in enumerate(leaf_names)}
    S = np.zeros((2 * m - 1, m), dtype=bool)
    for i, node in enumerate(nodes):
        S[i, [leaf_map[c] for c in node.name.split("+")]] = True
    MS = np.array([MS_by_name[c] for c in leaf_names])
    l_TN = np.log((1 - MS) * (1 - FP) + MS)
    l_DO = np.log((1 - MS) * FN + MS)
    weights, weights_norm, root_id = branch_weights(S, l_TN, l_DO, w_max)
    nodes.pop(root_id)
    for i, node in enumerate(nodes):
        node.weights = weights[i]
        node.weights_norm = weights_norm[i]
    root = tree.get_tree_root().children[0]
    root.weights = np.full(2, -1)
    root.weights_norm = np.full(2, -1)
    return weights, weights_norm


def get_gt_tree(tree_file, mut, w_max, FN_fix=None, FP_fix=None):
    """run_PT_test.py:337-365.  ``mut`` is get_mut_df's dict.  Returns a dict with
    the tree, rates, the reduced matrix, S, and the soft assignment."""
    muts, cols = mut["muts"], list(mut["samples"])
    tree, outg, FP, FN = read_tree(tree_file, mut["samples"])
    if FP_fix and FN_fix:
        FP = max(FP_fix, LAMBDA_MIN)
        FN = max(FN_fix, LAMBDA_MIN)
    else:
        FP = max(FP, LAMBDA_MIN)
        FN /= 2
        FN = max(FN, LAMBDA_MIN)
    if outg in cols:                                                            # :350-354
        k = cols.index(outg)
        red = np.delete(muts, k, axis=1)
        red_cols = cols[:k] + cols[k + 1:]
        red = red[np.nansum(red, axis=1) > 0]
    else:
        red, red_cols = muts, cols
    MS_i = np.clip(np.isnan(red).mean(axis=0), LAMBDA_MIN, 1 - FN - 0.2)        # :358
    MS_by_name = dict(zip(red_cols, MS_i))
    S, nodes = clade_matrix(tree, red_cols)
    soft = map_mutations_gt(red, S, FP, FN)
    for i, node in enumerate(nodes):                                            # :443-445
        node.mut_no_soft = soft[i]
        node.dist = soft[i]
    weights, weights_norm = add_br_weights(tree, FP, FN, MS_by_name, w_max)
    return {"tree": tree, "FP": FP, "FN": FN, "muts_red": red, "columns": red_cols, "S": S,
            "soft": soft, "MS": MS_i, "weights": weights, "weights_norm": weights_norm, "n_muts": red.shape[0]}


# -------------------------------------------------------------- LRT model
def get_model_data(tree, pseudo_mut=0):
    """run_PT_test.py:518-635 (true_data=False).  Returns Y, constr, init,
    weights_norm, constr_cols."""
    leaf_no = len(tree)
    br_no = 2 * leaf_no - 2
    Y = np.zeros(br_no)
    init = np.zeros(br_no)
    weights_norm = np.zeros((br_no, tree.get_tree_root().children[0].weights.size))
    constr = np.zeros((leaf_no - 1, br_no), dtype=int)
    constr_cols = []
    br_cells = {}

    def get_traversal(start_node):
        traverse = np.zeros(br_no, dtype=int)
        traverse[br_cells[start_node]] = 1
        node = start_node
        while len(node.children) != 0:
            w0 = node.children[0].weights_norm[0]
            w1 = node.children[1].weights_norm[0]
            nxt = node.children[0] if w0 < w1 else node.children[1]
            traverse[br_cells[nxt]] = 1
            node = nxt
        return traverse

    node_idx = 0

    def add_node(node, l_idx, neighbor=None):
        br_cells[node] = node_idx
        Y[node_idx] = node.mut_no_soft + pseudo_mut
        weights_norm[node_idx] = node.weights_norm
        if neighbor is None:
            constr[l_idx, node_idx] = 1
        else:
            constr[l_idx] = get_traversal(neighbor) - get_traversal(node)

    int_nodes = [i for i in tree.iter_descendants("postorder") if not i.is_leaf()]
    for l_idx, int_node in enumerate(int_nodes):
        is_terminal = [i.is_leaf() for i in int_node.children]
        if all(is_terminal):                                                    # :574-583
            add_node(int_node.children[0], l_idx)
            init[node_idx] = max(Y[node_idx] + pseudo_mut, LAMBDA_MIN)
            constr_cols.append(int_node.children[0].name)
            node_idx += 1
            add_node(int_node.children[1], l_idx)
            constr[l_idx, node_idx] = -1
            constr_cols.append(int_node.children[1].name)
        elif any(is_terminal):                                                  # :586-597
            terminal = int_node.children[is_terminal.index(True)]
            add_node(terminal, l_idx)
            init[node_idx] = max(Y[node_idx] + pseudo_mut, LAMBDA_MIN)
            constr_cols.append(terminal.name)
            node_idx += 1
            internal = int_node.children[is_terminal.index(False)]
            add_node(internal, l_idx, neighbor=terminal)
            constr_cols.append(internal.name)
        else:                                                                   # :600-611
            shorter, longer = sorted(int_node.children, key=lambda x: x.mut_no_soft)
            add_node(shorter, l_idx)
            init[node_idx] = max(Y[node_idx] + pseudo_mut, LAMBDA_MIN)
            constr_cols.append(shorter.name)
            node_idx += 1
            add_node(longer, l_idx, neighbor=shorter)
            constr_cols.append(longer.name)
        init_val = np.dot(constr[l_idx], init)                                  # :613-622
        if init_val >= LAMBDA_MIN:
            init[node_idx] = init_val
        else:
            init[node_idx - 1] += -init_val + 0.1
            init[node_idx] = np.dot(constr[l_idx], init)
        node_idx += 1
    assert np.allclose(constr @ init, 0, atol=LAMBDA_MIN), "Constraints not fulfilled for x_0"
    assert (init >= LAMBDA_MIN).all(), "Init value smaller than min. distance"
    return Y, constr, init, weights_norm, np.array(constr_cols)


def lrt_scale_fac(ll_H1):
    """run_PT_test.py:640-643."""
    scale_fac = 0.0
    for i in [10000, 1000, 100, 10, 1, 0.1]:
        scale_fac = (ll_H1 // i) * i
        if scale_fac != 0:
            break
    return scale_fac


def lrt_pvalue(LR, dof, on_bound):
    """run_PT_test.py:687-695."""
    from scipy.stats.distributions import chi2
    from scipy.special import binom as binomCoeff
    if on_bound > 0:
        dof_diff = np.arange(on_bound + 1)
        with np.errstate(all="ignore"):
            p_vals = np.clip(chi2.sf(LR, dof - dof_diff), 1e-100, 1)
        p_val_weights = np.where(np.isnan(p_vals), 0, binomCoeff(on_bound, dof_diff))
        return float(np.average(np.nan_to_num(p_vals), weights=p_val_weights))
    return float(chi2.sf(LR, dof))


def get_LRT_poisson(Y, constr, init, weights):
    """run_PT_test.py:638-700 with alg='trust-constr' (same SciPy call, same
    options).  Returns LR, dof, on_bound, p_val, x_opt."""
    from scipy.optimize import minimize
    ll_H1 = np.sum((Y * np.log(np.where(Y > 0, Y, 1)) - Y) * weights)
    scale_fac = lrt_scale_fac(ll_H1)

    def fun_opt(l, Y):
        return -np.nansum((Y * np.log(np.where(l > LAMBDA_MIN, l, np.nan)) - l) * weights) / scale_fac

    def fun_jac(l, Y):
        return -(Y / np.clip(l, LAMBDA_MIN, None) - 1) * weights / scale_fac

    def fun_hess(l, Y):
        return np.identity(Y.size) * (Y / np.clip(l, LAMBDA_MIN, None) ** 2) * weights / scale_fac

    const = [{"type": "eq", "fun": lambda x: np.matmul(constr, x)}]
    init = np.clip(init, LAMBDA_MIN, None)
    bounds = np.full((Y.size, 2), (LAMBDA_MIN, Y.sum()))
    with np.errstate(under="ignore"):
        opt = minimize(fun_opt, init, args=(Y,), constraints=const, bounds=bounds, method="trust-constr",
                       jac=fun_jac, hess=fun_hess, options={"disp": False, "maxiter": 10000, "barrier_tol": 1e-5})
        if not opt.success:
            opt = minimize(fun_opt, init, args=(Y,), constraints=const, bounds=bounds, method="SLSQP", jac=fun_jac,
                           options={"disp": False, "maxiter": 10000})
            if not opt.success:
                return np.nan, np.nan, np.nan, np.nan, None
    ll_H0 = np.nansum((Y * np.log(opt.x) - opt.x) * weights)
    dof = round(weights.sum() - constr.shape[0])
    LR = -2 * (ll_H0 - ll_H1)
    on_bound = int(np.sum(opt.x <= LAMBDA_MIN ** 0.5))
    return LR, dof, on_bound, lrt_pvalue(LR, dof, on_bound), opt.x


def run_poisson_tree_test(vcf_file, tree_file, w_maxs, exclude="", include="", FN_in=None, FP_in=None):
    """run_PT_test.py:703-728: returns (tsv text, per-w result dicts)."""
    header_str = "muts\tFN\tFP"
    model_str = ""
    w_cols = ["-2logLR", "dof", "p-value", "hypothesis"]
    mut = get_mut_df(vcf_file, exclude, include)
    results = []
    for w_max in w_maxs:
        g = get_gt_tree(tree_file, mut, w_max, FN_in, FP_in)
        Y, constr, init, weights_norm, constr_cols = get_model_data(g["tree"])
        LR, dof, on_bound, p_val, x = get_LRT_poisson(Y, constr, init, weights_norm[:, 0])
        hyp = int(p_val < 0.05)
        header_str += "\t" + "\t".join([f"{i}_{w_max:.0f}" for i in w_cols])
        model_str += f"\t{LR:0>5.3f}\t{dof}\t{p_val}\tH{hyp}"
        results.append({"w_max": w_max, "LR": LR, "dof": dof, "on_bound": on_bound, "p_val": p_val, "x": x,
                        "Y": Y, "constr": constr, "init": init, "weights_norm": weights_norm[:, 0],
                        "constr_cols": constr_cols, "g": g})
    g = results[-1]["g"]
    model_str = f"{g['n_muts']}\t{g['FN']:.4f}\t{g['FP']:.4f}{model_str}"
    return f"{header_str}\n{model_str}", results
