#!/usr/bin/env python3
"""Poisson Tree (PT) test -- host side of the B200-native hot path.

Mirrors the reference module of the same name (cbg-ethz/scSomMerClock,
run_PT_test.py): same function names, argument meaning, command line and TSV
output, so it drops in for that path.  The work is done by hand-written sm_100a
CUDA behind the C ABI (include/scsommerclock_b200.h):

  get_mut_df        -> scs_vcf_decode          (VCF text -> packed SNV x cell matrix)
  get_gt_tree       -> scs_gt_reduce           (row filter, per-cell missing rates)
  map_mutations_gt  -> scs_placement_*         (two-pass tcgen05 GEMM + softmax epilogue)
  add_br_weights    -> scs_branch_weights
  get_LRT_poisson   -> scs_lrt_poisson_batch   (barrier-Newton MLE, LRT, p-value)

What stays on the host is what the reference does with ete3/regexes: reading
files, the #CHROM line, the Newick tree and its clade/branch bookkeeping.  There
is no CPU fallback for the numerical path: without the built library and a
B200 the functions raise.
"""
import argparse
import gzip
import os
import re

import numpy as np

from . import engine
from .tree import NewickError, Tree

LAMBDA_MIN = 1e-6          # run_PT_test.py:21
log_LAMBDA_MIN = -100      # run_PT_test.py:22

_CTX = {}


def get_context(device=None):
    """The process-wide scs context (one per device)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if "SCS_DEVICE" not in os.environ else int(os.environ["SCS_DEVICE"])
    if device not in _CTX:
        _CTX[device] = engine.Context(device)
    return _CTX[device]


# ------------------------------------------------------------------ containers
class GenotypeMatrix:
    """The reference's ``muts`` DataFrame (run_PT_test.py:183), kept on the device
    as the packed 2-bit SNV x cell matrix.  ``columns`` are the sample names,
    ``index`` the (chrom, pos) of every kept SNV, ``values`` materialises the
    {0,1,2,NaN} matrix on the host when somebody asks for it."""

    def __init__(self, ctx, decoded, columns):
        self.ctx = ctx
        self._dec = decoded
        self.columns = np.array(columns)
        self._host = None
        self._cache = {}

    @property
    def shape(self):
        return (self._dec.n_kept, len(self.columns))

    def _copy(self):
        if self._host is None:
            self._host = self._dec.copy()
        return self._host

    @property
    def index(self):
        return [tuple(int(v) for v in p) for p in self._copy()[1]]

    @property
    def codes(self):
        return engine.unpack_genotypes(self._copy()[0], len(self.columns))

    @property
    def values(self):
        c = self.codes.astype(np.float64)
        c[c == engine._cabi.GT_MISSING] = np.nan
        return c

    def device_tensor(self):
        """Zero-copy torch view (uint8 [n_snv, pitch]) of the packed matrix in device memory."""
        import torch

        class _View:
            pass
        v = _View()
        n, pitch = self._dec.n_kept, self._dec.pitch
        v.__cuda_array_interface__ = {"shape": (n, pitch), "typestr": "|u1", "data": (self._dec.device_ptr(), False), "version": 2}
        return torch.as_tensor(v, device=torch.device("cuda", self.ctx.device))

    def to_dataframe(self):
        import pandas as pd
        return pd.DataFrame(self.values, index=pd.MultiIndex.from_tuples(self.index), columns=self.columns)


class PlacementResult:
    """Stands in for the DataFrame ``M`` returned by map_mutations_gt
    (run_PT_test.py:447-448).  The SNV x branch matrix itself never exists: only
    its column sums (``soft``) and its shape are available."""

    def __init__(self, n_muts, columns, soft):
        self.shape = (int(n_muts), len(columns))
        self.columns = list(columns)
        self.soft = np.asarray(soft)

    def sum(self, axis=0):
        if axis != 0:
            raise NotImplementedError("only column sums of M exist on the GPU path")
        return self.soft


class ClockConstraints:
    """The equality constraints of get_model_data (run_PT_test.py:529, :562-565)
    in tree form.  ``child_int[k]`` is the internal node (post-order index) below
    branch k, or -1 for a leaf; branch k hangs below internal node k // 2.
    ``path_next[k]`` is the branch the reference's get_traversal continues with
    below branch k (-1 at a leaf).  ``toarray()`` builds the reference's dense
    integer matrix; the GPU solver never needs it."""

    def __init__(self, child_int, path_next):
        self.child_int = np.asarray(child_int, dtype=np.int32)
        self.path_next = np.asarray(path_next, dtype=np.int64)
        self.shape = (self.child_int.size // 2, self.child_int.size)

    def _traversal(self, k):
        out = []
        while k >= 0:
            out.append(k)
            k = self.path_next[k]
        return out

    def toarray(self):
        n_int, br = self.shape
        constr = np.zeros((n_int, br), dtype=int)
        for i in range(n_int):
            a, b = 2 * i, 2 * i + 1
            if self.child_int[a] < 0 and self.child_int[b] < 0:
                constr[i, a], constr[i, b] = 1, -1
            else:
                constr[i, self._traversal(a)] += 1
                constr[i, self._traversal(b)] -= 1
        return constr

    def __array__(self, dtype=None, copy=None):
        a = self.toarray()
        return a.astype(dtype) if dtype is not None else a

    def __matmul__(self, x):
        x = np.asarray(x, dtype=float)
        out = np.zeros(self.shape[0])
        for i in range(self.shape[0]):
            out[i] = x[self._traversal(2 * i)].sum() - x[self._traversal(2 * i + 1)].sum()
        return out

    @staticmethod
    def from_dense(constr):
        """Recover the tree form from the reference's dense matrix."""
        constr = np.asarray(constr)
        n_int, br = constr.shape
        child_int = np.full(br, -1, dtype=np.int32)
        path_next = np.full(br, -1, dtype=np.int64)
        for i in range(n_int):
            for k, sign in ((2 * i, 1), (2 * i + 1, -1)):
                path = [j for j in np.flatnonzero(constr[i] == sign) if j != k]
                # the chain below branch k visits nodes in decreasing post-order index
                path.sort(key=lambda j: -(j // 2))
                prev = k
                for j in path:
                    child_int[prev] = j // 2
                    path_next[prev] = j
                    prev = j
        # every internal node j sits on the recorded path of its parent's row (the branch
        # above j is followed by one of j's own branches), so child_int is complete
        return ClockConstraints(child_int, path_next)


# --------------------------------------------------------------------- VCF
def _read_text(vcf_file):
    if vcf_file.endswith("gz"):
        with gzip.open(vcf_file, "rb") as f:
            return f.read()
    with open(vcf_file, "rb") as f:
        return f.read()


def _prefetch_texts(files, workers=None, window=None):
    """Yield ``_read_text(f)`` for every file, in order, reading (and gunzipping -- zlib releases the GIL)
    up to ``window`` files ahead on ``workers`` threads: a sweep over thousands of small VCFs is bound by
    decompression on the host, not by the GPU."""
    from concurrent.futures import ThreadPoolExecutor
    files = list(files)
    workers = workers or max(1, min(8, (os.cpu_count() or 2) - 1))
    window = window or 2 * workers
    if workers == 1 or len(files) <= 1:
        for f in files:
            yield _read_text(f)
        return
    with ThreadPoolExecutor(max_workers=workers) as ex:
        futs, nxt = {}, 0
        for i in range(len(files)):
            while nxt < len(files) and nxt < i + window:
                futs[nxt] = ex.submit(_read_text, files[nxt])
                nxt += 1
            yield futs.pop(i).result()


def _header_samples(text):
    """Sample names from the (last) #CHROM line, renamed like run_PT_test.py:49-56."""
    sample_names = None
    pos = 0
    n = len(text)
    while pos < n:
        end = text.find(b"\n", pos)
        if end < 0:
            end = n
        if text[pos:pos + 1] == b"#":
            if text[pos:pos + 6] == b"#CHROM":
                line = text[pos:end + 1].decode()
                if line.count("tumcell") == 0:
                    line = re.sub(r"cell(?=\d+)", "tumcell", line)
                if line.count("healthycell") == 0:
                    line = re.sub("outgcell", "healthycell", line)
                sample_names = [i.strip() for i in line.split("\t")[9:]]
        elif text[pos:end].strip() != b"":
            break       # first record: the header block is over
        pos = end + 1
    if sample_names is None:
        raise ValueError("VCF has no #CHROM header line")
    return sample_names


def get_mut_df(vcf_file, exclude_pat, include_pat, filter=True, ctx=None, _text=None, shard_rows=False):
    """run_PT_test.py:27-207.  Returns ``(muts, true_muts, read_data)`` like the
    reference; ``muts`` is a device-resident GenotypeMatrix, the other two (true
    genotypes and read counts, which the PT test never touches) are None.
    ``_text``: the file's decompressed bytes when the caller has read them already.
    ``shard_rows``: opt-in, set by run_poisson_tree_test under torchrun only -- this rank decodes its own
    block of record lines and the caller all-reduces what depends on all rows (kept SNVs, missing counts,
    soft weights).  Replicate sweeps shard whole pairs instead and never set it."""
    if not filter:
        raise NotImplementedError("filter=False is not on the PT-test path")
    ctx = ctx or get_context()
    text = _text if _text is not None else _read_text(vcf_file)
    sample_names = _header_samples(text)
    exclude = []
    if exclude_pat != "":                                                       # :57-65
        print("Exluded:")
        for sample in sample_names:
            if re.fullmatch(exclude_pat, sample):
                exclude.append(sample)
                print("\t{}".format(sample))
        if len(exclude) == 0:
            print("\nWARNING: no samples with pattern {} in vcf!\n".format(exclude_pat))
    if include_pat != "":                                                       # :66-76
        print("Include:")
        for sample in sample_names:
            if not re.fullmatch(include_pat, sample):
                exclude.append(sample)
            elif not re.fullmatch(exclude_pat, sample):
                print("\t{}".format(sample))
        if len(exclude) == 0:
            print("\nWARNING: no samples with pattern {} in vcf!\n".format(include_pat))
    exclude.append("")                                                          # :187
    include_idx = [i for i, j in enumerate(sample_names) if j not in exclude]
    include_id = [sample_names[i] for i in include_idx]
    if not include_idx:
        raise ValueError("no sample left after -excl/-incl")
    sharded = False
    if shard_rows and _row_sharding_active():
        # one large dataset on several GPUs: every rank decodes and places its own block of record
        # lines; the branch weights are all-reduced later
        text = _shard_record_lines(text)
        sharded = True
    dec = engine.DecodedVcf(ctx, text, len(sample_names), include_idx)
    muts = GenotypeMatrix(ctx, dec, include_id)
    muts.row_sharded = sharded
    return muts, None, None


def _row_sharding_active():
    """True under torchrun with an initialised process group (and SCS_SHARD_ROWS != 0)."""
    if os.environ.get("SCS_SHARD_ROWS", "1") == "0":
        return False
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def _shard_record_lines(text):
    """This rank's share of the VCF: the header block plus a contiguous block of record lines
    (cut at line boundaries, sizes balanced by bytes)."""
    import torch.distributed as dist
    from . import sharding
    rank, world = dist.get_rank(), dist.get_world_size()
    pos, n = 0, len(text)
    while pos < n and (text[pos:pos + 1] == b"#" or text[pos:text.find(b"\n", pos) if text.find(b"\n", pos) >= 0 else n].strip() == b""):
        nl = text.find(b"\n", pos)
        pos = n if nl < 0 else nl + 1
    body0 = pos
    lo, hi = sharding.shard_range(n - body0, rank, world)

    def snap(off):                    # first line start at or after body0 + off
        if off <= 0:
            return body0
        if body0 + off >= n:
            return n
        nl = text.find(b"\n", body0 + off - 1)
        return n if nl < 0 else nl + 1

    return text[:body0] + text[snap(lo):snap(hi)]


# -------------------------------------------------------------------- tree
def read_tree(tree_file, samples=[]):
    """run_PT_test.py:210-334 (host work: regexes and tree surgery)."""
    samples = np.asarray(samples)
    with open(tree_file, "r") as f:
        tree_raw = f.read().strip()
    cellphy_ends = (".raxml.bestTree", "raxml.mutationMapTree", "raxml.supportFBP", "raxml.supportTBE")
    if tree_file.endswith(cellphy_ends) or "cellphy" in tree_file:
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
    elif tree_file.endswith("_ml0.newick") or "scite" in tree_file:
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
            FN *= 2     # SCITE's beta is per allele
        elif os.path.exists(error_file):
            with open(error_file, "r") as f:
                errors_raw = f.read().strip().split("\n")
            FP, FN = [float(i) for i in errors_raw[1].split(",")]
            FN *= 2
        else:
            FP = LAMBDA_MIN
            FN = LAMBDA_MIN
    else:
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
    outg_node = tree & outg_name
    anc_node = outg_node.get_ancestors()[0]
    outg_node.support = 100 if anc_node.is_root() else anc_node.support
    tree.set_outgroup(outg_node)
    outg_node.delete()
    sample_set = set(samples.tolist())
    for node in tree.get_leaves():
        if node.name not in sample_set:
            node.delete()
    weight_no = 2
    # name = '+'.join(sorted(leaf names)) for every node (:327), built bottom-up
    for node in tree.traverse("postorder"):
        node._leafset = [node.name] if node.is_leaf() else [x for ch in node.children for x in ch._leafset]
    for node in tree.iter_descendants():
        node.add_features(in_length=node.dist, muts_br=[[], [], []], muts_br_true=set([]), muts_br_false=set([]),
                          muts_br_missing=set([]), muts_node_true=set([]), mut_no_soft=0, mut_no_true=-1,
                          name="+".join(sorted(node._leafset)), weights=np.zeros(weight_no),
                          weights_norm=np.zeros(weight_no), weights_norm_z=np.zeros(weight_no), drivers=[])
    return tree, outg_name, FP, FN


def _clade_bits(tree, columns, active):
    """Bit mask of S (run_PT_test.py:389-397) over the packed matrix's columns:
    one row per ``tree.iter_descendants()`` node, zero rows up to 2n-1 (n = number
    of active columns), and the final all-zero 'FP' row."""
    col = {c: i for i, c in enumerate(columns)}
    nodes = list(tree.iter_descendants())
    n = int(np.sum(active))
    n_rows = max(2 * n - 1, len(nodes)) + 1
    words = (len(columns) + 31) // 32
    bits = np.zeros((n_rows, words), dtype=np.uint32)
    # leaves first, then parents as the OR of their children (reverse level order)
    index = {id(node): i for i, node in enumerate(nodes)}
    for i in range(len(nodes) - 1, -1, -1):
        node = nodes[i]
        if node.is_leaf():
            c = col[node._leafset[0]] if hasattr(node, "_leafset") else col[node.name]
            bits[i, c >> 5] |= np.uint32(1 << (c & 31))
        else:
            for ch in node.children:
                bits[i] |= bits[index[id(ch)]]
    return bits, words, nodes


def get_gt_tree(tree_file, call_data, w_max, FN_fix=None, FP_fix=None):
    """run_PT_test.py:337-365.  The placement does not depend on ``w_max``; it is
    computed once per (tree, error rates) and cached on the genotype matrix, so
    calling this in the reference's ``for w_max`` loop costs one GEMM, not ten."""
    muts = call_data[0]
    ctx = muts.ctx
    tree, outg, FP, FN = read_tree(tree_file, muts.columns)
    FP, FN = _error_rates(FP, FN, FN_fix, FP_fix)
    key = (os.path.abspath(tree_file), FP, FN)
    st = muts._cache.get(key)
    if st is None:
        st = _place_on_tree(ctx, muts, tree, outg, FP, FN)
        muts._cache[key] = st
    _apply_placement(tree, st)
    add_br_weights(tree, FP, FN, st["MS_i"], w_max, _state=st)
    M = PlacementResult(st["n_muts"], [n.name for n in tree.iter_descendants()] + ["FP"], st["soft"])
    return tree, FP, FN, M


def _place_on_tree(ctx, muts, tree, outg, FP, FN):
    import torch
    columns = list(muts.columns)
    active = np.array([c != outg for c in columns], dtype=np.uint8)
    has_outg = outg in columns
    n_snv, n_cols = muts.shape
    dev = torch.device("cuda", ctx.device)
    cur = torch.cuda.current_stream(dev)
    d_gt = muts._dec.device_ptr()
    bits, words, nodes = _clade_bits(tree, columns, active)
    sharded = getattr(muts, "row_sharded", False)
    y = torch.zeros(bits.shape[0], dtype=torch.float64, device=dev)
    n_kept, miss = 0, np.zeros(n_cols)
    if n_snv > 0 or not sharded:
        # row filter + per-cell missing counts (:350-358) and the placement (:362): the decoded matrix is read once on the
        # interval path (scs_placement_prepare), nothing comes back to the host before the soft weights do
        plan = engine.Placement(ctx, n_snv, n_cols, bits.shape[0])
        plan.set_clades_host(bits, words, stream=cur)
        plan.prepare(d_gt, muts._dec.pitch, active, has_outg, stream=cur)     # the decoded matrix outlives the run
        plan.run(FP, FN, y, stream=cur)
        n_kept, miss = plan.prepare_result(stream=cur)
        plan.close()
    if sharded:
        from . import sharding
        t = torch.cat([torch.tensor(np.concatenate([[float(n_kept)], miss * n_kept]), dtype=torch.float64, device=dev), y])
        sharding.allreduce_soft_weights(t)                 # the path's one exchange: kept SNVs, per-cell missing counts, 2n soft weights
        n_kept = int(round(t[0].item()))
        miss = (t[1:1 + n_cols] / max(n_kept, 1)).cpu().numpy()
        y = t[1 + n_cols:]
    soft = y.cpu().numpy()
    FN_cap = 1 - FN - 0.2
    MS_i = {c: float(np.clip(miss[i], LAMBDA_MIN, FN_cap)) for i, c in enumerate(columns) if active[i]}        # :358
    return {"soft": soft, "MS_i": MS_i, "n_muts": n_kept, "bits": bits, "words": words, "columns": columns,
            "weights": {}}


def _per_cell_rates(FN, columns, active):
    """``FN[muts_in.columns].values`` of run_PT_test.py:414 for a pd.Series / dict keyed by sample name, or a plain
    sequence with one rate per matrix column; columns outside the placement (the outgroup) may be absent."""
    out = np.full(len(columns), np.nan)
    if hasattr(FN, "keys"):                                  # pd.Series, dict
        for i, c in enumerate(columns):
            if c in FN.keys():
                out[i] = float(FN[c])
            elif active[i]:
                raise KeyError("no FN rate for cell %r" % (c,))
    else:
        arr = np.asarray(FN, dtype=np.float64).ravel()
        n_act = int(np.sum(active))
        if arr.size == len(columns):
            out[:] = arr
        elif arr.size == n_act:                              # rates for the placed cells only, in column order
            out[np.asarray(active, dtype=bool)] = arr
        else:
            raise ValueError("FN holds %d rates for %d cells" % (arr.size, n_act))
    return out


def _place_on_tree_percell(ctx, muts, tree, outg, FP, FN):
    """map_mutations_gt's per-cell-FN branch (:413-418) on the device: row filter (:350-353) by scs_gt_reduce, placement
    by scs_place_percell.  Under row sharding the soft weights are summed over the ranks like the scalar path's."""
    import torch
    columns = list(muts.columns)
    active = np.array([c != outg for c in columns], dtype=np.uint8)
    has_outg = outg in columns
    n_snv, n_cols = muts.shape
    dev = torch.device("cuda", ctx.device)
    cur = torch.cuda.current_stream(dev)
    bits, words, nodes = _clade_bits(tree, columns, active)
    rates = _per_cell_rates(FN, columns, active)
    y = torch.zeros(bits.shape[0], dtype=torch.float64, device=dev)
    n_kept = 0
    if n_snv > 0:
        d_gt = muts._dec.device_ptr()
        keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
        n_kept, _ = engine.gt_reduce(ctx, d_gt, muts._dec.pitch, n_snv, n_cols, active, has_outg, keep, stream=cur)
        engine.place_percell(ctx, d_gt, muts._dec.pitch, n_snv, n_cols, keep, bits, words, FP, np.where(active.astype(bool), rates, 0.5), y, stream=cur)     # (columns outside the tree carry no weight; a NaN rate of a placed cell is refused by the library)
    if getattr(muts, "row_sharded", False):
        from . import sharding
        t = torch.cat([torch.tensor([float(n_kept)], dtype=torch.float64, device=dev), y])
        sharding.allreduce_soft_weights(t)
        n_kept = int(round(t[0].item()))
        y = t[1:]
    return {"soft": y.cpu().numpy(), "n_muts": n_kept, "bits": bits, "words": words, "columns": columns}


def _apply_placement(tree, st):
    for i, node in enumerate(tree.iter_descendants()):                          # :443-445
        node.mut_no_soft = st["soft"][i]
        node.dist = st["soft"][i]


def map_mutations_gt(tree, muts_in, FP, FN):
    """run_PT_test.py:386-448 for callers that drive the steps themselves: places
    every SNV of ``muts_in`` (a GenotypeMatrix) on ``tree`` and sets
    ``node.mut_no_soft`` / ``node.dist``."""
    if not np.isscalar(FP):
        raise TypeError("FP is one rate for all cells (run_PT_test.py:409-418 has no per-cell FP branch)")
    if not np.isscalar(FN):
        # -- Taking different FN value per cell into account -- (:413-418): its own kernel (csrc/scs_place_percell.cuh),
        # the logits are prefix sums of per-cell weights, not functions of two integer counts
        st = _place_on_tree_percell(muts_in.ctx, muts_in, tree, "healthycell", FP, FN)
        _apply_placement(tree, st)
        return PlacementResult(st["n_muts"], [n.name for n in tree.iter_descendants()] + ["FP"], st["soft"])
    st = _place_on_tree(muts_in.ctx, muts_in, tree, "healthycell", FP, FN)
    _apply_placement(tree, st)
    return PlacementResult(st["n_muts"], [n.name for n in tree.iter_descendants()] + ["FP"], st["soft"])


def add_br_weights(tree, FP, FN, MS, w_max, _state=None):
    """run_PT_test.py:451-515.  ``MS``: dict / Series of per-cell missing rates."""
    ctx = get_context()
    nodes = list(tree.iter_descendants())
    m = len(tree)
    if len(nodes) != 2 * m - 1:
        raise ValueError("the PT test needs a binary tree: %d leaves but %d nodes" % (m, len(nodes)))
    if _state is not None:
        columns, bits, words = _state["columns"], _state["bits"][:len(nodes)], _state["words"]
    else:
        columns = sorted(l.name for l in tree.get_leaves())
        for node in tree.traverse("postorder"):
            node._leafset = [node.name] if node.is_leaf() else [x for ch in node.children for x in ch._leafset]
        bits, words, _ = _clade_bits(tree, columns, np.ones(len(columns), dtype=np.uint8))
        bits = bits[:len(nodes)]
    leaves = set(l.name for l in tree.get_leaves())
    rate = np.array([float(MS[c]) if c in leaves else -1.0 for c in columns])
    cache = _state["weights"] if _state is not None else {}
    if w_max not in cache:
        w, wn, root = engine.branch_weights(ctx, bits, words, len(columns), rate, FP, FN, [w_max])
        cache[w_max] = (w[0], wn[0], root)
    w, wn, root = cache[w_max]
    # second weight (odds ratio, :494) is not used by the PT test; it is filled with the same
    # clipping rule on the host only so that node.weights keeps the reference's shape
    for i, node in enumerate(nodes):
        node.weights = np.array([w[i], np.nan])
        node.weights_norm = np.array([wn[i], np.nan])
        node.weights_norm_z = np.array([np.nan, np.nan])
    root_node = tree.get_tree_root().children[0]
    root_node.weights = np.full(2, -1)
    root_node.weights_norm = np.full(2, -1)
    root_node.weights_norm_z = np.ones(2)


# -------------------------------------------------------------- LRT model
def get_model_data(tree, pseudo_mut=0, true_data=False):
    """run_PT_test.py:518-635.  Same outputs (Y, constr, init, weights_norm,
    constr_cols); ``constr`` is a ClockConstraints (tree form, ``toarray()`` gives
    the reference's dense matrix)."""
    if true_data:
        raise NotImplementedError("true_data=True needs the TG genotypes, which are not on the PT-test path")
    leaf_no = len(tree)
    br_no = 2 * leaf_no - 2
    Y = np.zeros(br_no, dtype=float)
    init = np.zeros(br_no, dtype=float)
    weights_norm = np.zeros((br_no, 2))
    child_int = np.full(br_no, -1, dtype=np.int32)
    path_next = np.full(br_no, -1, dtype=np.int64)
    constr_cols = []
    int_nodes = [i for i in tree.iter_descendants("postorder") if not i.is_leaf()]
    int_index = {id(n): i for i, n in enumerate(int_nodes)}
    br_of = {}

    def path_sum(k, vec):
        s = 0.0
        while k >= 0:
            s += vec[k]
            k = path_next[k]
        return s

    def add(node, k):
        br_of[id(node)] = k
        Y[k] = node.mut_no_soft + pseudo_mut
        weights_norm[k] = node.weights_norm
        constr_cols.append(node.name)
        if not node.is_leaf():
            if len(node.children) != 2:
                raise ValueError("the PT test needs a binary tree")
            child_int[k] = int_index[id(node)]
            # get_traversal (:533-551): follow the child with the higher weight
            w0 = node.children[0].weights_norm[0]
            w1 = node.children[1].weights_norm[0]
            nxt = node.children[0] if w0 < w1 else node.children[1]
            path_next[k] = br_of[id(nxt)]

    node_idx = 0
    for l_idx, int_node in enumerate(int_nodes):
        if len(int_node.children) != 2:
            raise ValueError("the PT test needs a binary tree")
        is_terminal = [i.is_leaf() for i in int_node.children]
        if all(is_terminal):                                                    # :574-583
            first, second = int_node.children
        elif any(is_terminal):                                                  # :586-597
            first = int_node.children[is_terminal.index(True)]
            second = int_node.children[is_terminal.index(False)]
        else:                                                                   # :600-611
            first, second = sorted(int_node.children, key=lambda x: x.mut_no_soft)
        add(first, node_idx)
        init[node_idx] = max(Y[node_idx] + pseudo_mut, LAMBDA_MIN)
        node_idx += 1
        add(second, node_idx)
        # init of the second branch: constr[l_idx] @ init with its own entry still zero (:613-622)
        init_val = path_sum(node_idx - 1, init) - path_sum(node_idx, init)
        if init_val >= LAMBDA_MIN:
            init[node_idx] = init_val
        else:
            init[node_idx - 1] += -init_val + 0.1
            init[node_idx] = path_sum(node_idx - 1, init) - path_sum(node_idx, init)
        node_idx += 1
    constr = ClockConstraints(child_int, path_next)
    viol = constr @ init
    assert np.allclose(viol, 0, atol=LAMBDA_MIN), "Constraints not fulfilled for x_0"
    assert (init >= LAMBDA_MIN).all(), f"Init value smaller than min. distance: {init.min()}"
    return Y, constr, init, weights_norm, np.array(constr_cols)


def _branch_order(tree):
    """The w_max-independent part of get_model_data (:553-611): the branches in the order the
    reference numbers them and, per branch, the index of the internal node it leads to (-1 for a
    leaf) -- all the LRT kernel needs besides Y and the weights.  Used by the sweep driver so
    that a tree is walked once, not once per w_max."""
    int_nodes = [i for i in tree.iter_descendants("postorder") if not i.is_leaf()]
    int_index = {id(n): i for i, n in enumerate(int_nodes)}
    order = []
    for int_node in int_nodes:
        if len(int_node.children) != 2:
            raise ValueError("the PT test needs a binary tree")
        a, b = int_node.children
        ta, tb = a.is_leaf(), b.is_leaf()
        if ta == tb:
            if not ta and b.mut_no_soft < a.mut_no_soft:      # two internal children: fewer mutations first (stable)
                a, b = b, a
        elif tb:                                              # one leaf: the leaf first
            a, b = b, a
        order.append(a)
        order.append(b)
    child_int = np.array([-1 if n.is_leaf() else int_index[id(n)] for n in order], dtype=np.int32)
    return order, child_int


def get_LRT_poisson(Y, constr, init, weights=np.array([]), alg="trust-constr"):
    """run_PT_test.py:638-700 on the GPU (``alg`` is accepted for signature
    compatibility; the solver is the barrier-Newton kernel, see csrc/scs_lrt.cu).
    Returns (LR, dof, on_bound, p_val, zip(x, deviance, weights))."""
    res = get_LRT_poisson_batch([(Y, constr, weights)])[0]
    return res


def get_LRT_poisson_batch(problems, ctx=None):
    """Many LRTs in one kernel launch: ``problems`` is a list of (Y, constr,
    weights).  Returns a list of the reference's 5-tuples."""
    ctx = ctx or get_context()
    Ys, ws, chs = [], [], []
    for Y, constr, weights in problems:
        if not isinstance(constr, ClockConstraints):
            constr = ClockConstraints.from_dense(constr)
        Ys.append(np.asarray(Y, dtype=np.float64))
        ws.append(np.asarray(weights, dtype=np.float64))
        chs.append(constr.child_int)
    out, xs = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=True)
    results = []
    for i, (Y, constr, weights) in enumerate(problems):
        LR, dof, on_bound, p_val, ll_H0, ll_H1, iters, status = out[i]
        if status != 0 or not np.isfinite(LR):
            print("\nFAILED POISSON OPTIMIZATION\n")
            if status in (1, 3):          # (the library has already made its second attempt, scs_lrt_poisson_batch: :671-678)
                print("\nFAILED POISSON OPTIMIZATION, TWICE\n")
            results.append((np.nan, np.nan, np.nan, np.nan, np.nan, np.nan))
            continue
        x = xs[i]
        Yv, wv = Ys[i], ws[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            devi = ((Yv * np.log(np.where(Yv > 0, Yv, 1)) - Yv) * wv) - ((Yv * np.log(x) - x) * wv)
        results.append((float(LR), int(dof), int(on_bound), float(p_val), zip(x, devi, wv)))
    return results


def run_poisson_tree_test(vcf_file, tree_file, out_file, w_maxs, exclude, include, FN_in, FP_in):
    """run_PT_test.py:703-728: same console text, same TSV."""
    print("\nRunning the PT test:\n\t" f"- vcf: {vcf_file}\n\t- tree: {tree_file}\n\t- output: {out_file}")
    header_str = "muts\tFN\tFP"
    model_str = ""
    w_cols = ["-2logLR", "dof", "p-value", "hypothesis"]
    w_idx = 0
    call_data = get_mut_df(vcf_file, exclude, include, shard_rows=_row_sharding_active())
    problems = []
    for w_max in w_maxs:
        tree, FP, FN, M = get_gt_tree(tree_file, call_data, w_max, FN_in, FP_in)
        Y, constr, init, weights_norm, constr_cols = get_model_data(tree)
        problems.append((Y, constr, weights_norm[:, w_idx]))
    results = get_LRT_poisson_batch(problems)      # all w_max values in one launch
    for w_max, (LR, dof, on_bound, p_val, opt_vals) in zip(w_maxs, results):
        hyp = int(p_val < 0.05)
        header_str += "\t" + "\t".join([f"{i}_{w_max:.0f}" for i in w_cols])
        model_str += f"\t{LR:0>5.3f}\t{dof}\t{p_val}\tH{hyp}"
    model_str = f"{M.shape[0]}\t{FN:.4f}\t{FP:.4f}{model_str}"
    if not getattr(call_data[0], "row_sharded", False) or int(os.environ.get("RANK", "0")) == 0:
        with open(out_file, "w") as f_out:
            f_out.write(f"{header_str}\n{model_str}")
    print("Running successful\n")


def _error_rates(FP, FN, FN_fix, FP_fix):
    """run_PT_test.py:341-347."""
    if FP_fix and FN_fix:
        return max(FP_fix, LAMBDA_MIN), max(FN_fix, LAMBDA_MIN)
    return max(FP, LAMBDA_MIN), max(FN / 2, LAMBDA_MIN)


def _tree_source(tree_file):
    """The file-level part of read_tree (:210-231, :281-294) without building the tree: (kind, text, FP, FN) with kind
    'cellphy' (``[n]`` tags are dropped), 'plain', or 'scite' (integer node labels: handled by read_tree only)."""
    with open(tree_file, "r") as f:
        tree_raw = f.read().strip()
    cellphy_ends = (".raxml.bestTree", "raxml.mutationMapTree", "raxml.supportFBP", "raxml.supportTBE")
    if tree_file.endswith(cellphy_ends) or "cellphy" in tree_file:
        log_file = ".".join(tree_file.split(".")[:-1]) + ".log"
        if os.path.exists(log_file):
            with open(log_file, "r") as f:
                log = f.read().strip()
            FP = float(re.search(r"SEQ_ERROR: (0.\d+(e-\d+)?)", log).group(1))
            FN = float(re.search(r"ADO_RATE: (0.\d+(e-\d+)?)", log).group(1))
        else:
            FP = FN = LAMBDA_MIN
        return "cellphy", tree_raw, FP, FN
    if tree_file.endswith("_ml0.newick") or "scite" in tree_file:
        return "scite", tree_raw, None, None
    try:
        FP = float(re.search(r"WGA0[\.\d,]*-0[\.\d]*-(0[\.\d]*)", tree_file).group(1))
        FN = float(re.search(r"WGA(0[\.\d]*)[,\.\d]*?-", tree_file).group(1))
    except AttributeError:
        FP = FN = LAMBDA_MIN
    return "plain", tree_raw, FP, FN


def _tsv_header(w_maxs):
    w_cols = ["-2logLR", "dof", "p-value", "hypothesis"]
    return "muts\tFN\tFP" + "".join("\t" + "\t".join([f"{c}_{w_max:.0f}" for c in w_cols]) for w_max in w_maxs)


def run_poisson_tree_test_batch(vcf_files, tree_files, out_files, w_maxs, exclude="", include="", FN_in=None, FP_in=None,
                                ctx=None, shard=True):
    """Many independent (VCF, tree) pairs -- a simulation sweep (BASELINE configs[4]).  Same TSV per pair as
    run_poisson_tree_test.  Under torchrun the pairs are dealt to the ranks (no communication, no row sharding).  On each
    rank the pairs that share their sample columns go through engine.ReplicateSweep: tree parsing on host threads in native
    code, then ONE call each for the row filters, the placement (block-diagonal, one launch per pass), branch weights +
    LRT branch order, and all (pair, w_max) clock tests -- no Python loop over pairs touches a tree or waits for the
    device.  infSCITE trees, trees that lack some of the VCF's cells and matrices wider than 512 cells take the per-pair
    path (``_run_pairs``).  Returns the list of (pair index, [(LR, dof, on_bound, p), ...]) this rank handled."""
    import torch
    from . import sharding
    ctx = ctx or get_context()
    rank, world, _ = sharding.env_world()
    mine = sharding.replicate_shard(len(vcf_files), rank, world) if shard else list(range(len(vcf_files)))
    w_maxs = [float(w) for w in w_maxs]
    pairs = []
    for i, text in zip(mine, _prefetch_texts([vcf_files[i] for i in mine])):    # files are read / gunzipped ahead on host threads
        muts = get_mut_df(vcf_files[i], exclude, include, ctx=ctx, _text=text, shard_rows=False)[0]   # whole pairs are sharded, not rows
        if muts.shape[0] == 0:
            raise ValueError("%s: no SNV passes the filters" % vcf_files[i])
        kind, raw, FP, FN = _tree_source(tree_files[i])
        pairs.append(dict(i=i, muts=muts, kind=kind, raw=raw, FP=FP, FN=FN))
    results = {}
    slow = []
    fast = {}
    force_pairs = os.environ.get("SCS_SWEEP_PATH", "") == "pairs"      # (tests: the per-pair path on inputs the batched path would take)
    for p in pairs:
        if force_pairs or p["kind"] == "scite" or p["muts"].shape[1] > 512 or p["muts"].shape[1] < 3:
            slow.append(p)
        else:
            fast.setdefault((tuple(p["muts"].columns), p["kind"]), []).append(p)
    header = _tsv_header(w_maxs)
    for (columns, kind), grp in fast.items():
        rates = [_error_rates(p["FP"], p["FN"], FN_in, FP_in) for p in grp]
        sweep = engine.ReplicateSweep(ctx, [p["muts"].shape[0] for p in grp], list(columns), w_maxs)
        try:
            sweep.parse_trees([p["raw"] for p in grp], strip_tags=(kind == "cellphy"))
        except engine._cabi.ScsError:
            slow.extend(grp)                      # let the per-pair path name the offending file
            sweep.close()
            continue
        packed = torch.cat([p["muts"].device_tensor() for p in grp], dim=0) if len(grp) > 1 else grp[0]["muts"].device_tensor()
        sweep.enqueue(packed.contiguous(), np.array([r[0] for r in rates]), np.array([r[1] for r in rates]))
        out, n_kept, status = sweep.results()
        rows = sweep.rows()
        for k, p in enumerate(grp):
            if status[k] != 0:
                slow.append(p)                    # e.g. a tree that lacks some of the VCF's cells
                continue
            if np.isin(out[k, :, 7], (1, 3)).any():
                slow.append(p)                    # a failed solve: the per-pair path makes the second attempt (:671-678)
                continue
            if (out[k, :, 7] != 0).any():
                print("\nFAILED POISSON OPTIMIZATION\n")
            with open(out_files[p["i"]], "w") as f_out:
                f_out.write(f"{header}\n{rows[k]}")
            results[p["i"]] = [(float(o[0]), int(o[1]), int(o[2]), float(o[3])) if o[7] == 0 and np.isfinite(o[0])
                               else (np.nan, np.nan, np.nan, np.nan) for o in out[k]]
        sweep.close()
    if slow:
        results.update(_run_pairs(ctx, slow, tree_files, out_files, w_maxs, FN_in, FP_in))
    return [(p["i"], results[p["i"]]) for p in pairs]


def _run_pairs(ctx, pairs, tree_files, out_files, w_maxs, FN_in, FP_in):
    """The per-pair form of the sweep: every tree goes through read_tree and the host-side model set-up; pairs of equal
    shape still share one placement launch per pass and all clock tests one LRT launch."""
    import torch
    dev = torch.device("cuda", ctx.device)
    items = []
    for p in pairs:
        i, muts = p["i"], p["muts"]
        tree, outg, FP, FN = read_tree(tree_files[i], muts.columns)
        FP, FN = _error_rates(FP, FN, FN_in, FP_in)
        columns = list(muts.columns)
        active = np.array([c != outg for c in columns], dtype=np.uint8)
        n_snv, n_cols = muts.shape
        d_gt = muts.device_tensor()
        keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
        n_kept, miss = engine.gt_reduce(ctx, d_gt, muts._dec.pitch, n_snv, n_cols, active, outg in columns, keep)
        MS_i = {c: float(np.clip(miss[k], LAMBDA_MIN, 1 - FN - 0.2)) for k, c in enumerate(columns) if active[k]}
        bits, words, nodes = _clade_bits(tree, columns, active)
        items.append(dict(i=i, muts=muts, tree=tree, FP=FP, FN=FN, columns=columns, d_gt=d_gt, keep=keep, n_kept=n_kept, MS_i=MS_i,
                          bits=bits, words=words, n_cols=n_cols, n_rows=bits.shape[0], pitch=muts._dec.pitch))
    # ---- placement: one batched plan per (cells, clade rows) shape
    shapes = {}
    for it in items:
        shapes.setdefault((it["n_cols"], it["n_rows"], it["pitch"]), []).append(it)
    for (n_cols, n_rows, pitch), grp in shapes.items():
        plan = engine.Placement(ctx, [g["muts"].shape[0] for g in grp], n_cols, n_rows)
        plan.set_genotypes(torch.cat([g["d_gt"] for g in grp], dim=0).contiguous(), pitch, torch.cat([g["keep"] for g in grp]).contiguous())
        plan.set_clades_host(np.concatenate([g["bits"] for g in grp], axis=0), grp[0]["words"])
        soft = np.atleast_2d(plan.run_host(np.array([g["FP"] for g in grp]), np.array([g["FN"] for g in grp])).reshape(len(grp), n_rows))
        plan.close()
        for g, sw in zip(grp, soft):
            g["st"] = {"soft": sw, "MS_i": g["MS_i"], "n_muts": g["n_kept"], "bits": g["bits"], "words": g["words"],
                       "columns": g["columns"], "weights": {}}
    # ---- LRT: every (pair, w_max) in one launch.  Per pair the tree is walked once (branch order and
    # topology do not depend on w_max) and the branch weights of all w_max come from one call.
    Ys, ws, chs = [], [], []
    for it in items:
        tree = it["tree"]
        _apply_placement(tree, it["st"])
        nodes = list(tree.iter_descendants())
        if len(nodes) != 2 * len(tree) - 1:
            raise ValueError("the PT test needs a binary tree: %d leaves but %d nodes" % (len(tree), len(nodes)))
        leaves = set(l.name for l in tree.get_leaves())
        rate = np.array([float(it["MS_i"][c]) if c in leaves else -1.0 for c in it["columns"]])
        _, wn, _ = engine.branch_weights(ctx, it["bits"][:len(nodes)], it["words"], len(it["columns"]), rate, it["FP"], it["FN"],
                                         list(w_maxs))
        order, child_int = _branch_order(tree)
        row_of = {id(n): k for k, n in enumerate(nodes)}
        rows = np.array([row_of[id(n)] for n in order], dtype=np.int64)
        Y = np.asarray(it["st"]["soft"], dtype=np.float64)[rows]
        for wi in range(len(w_maxs)):
            Ys.append(Y)
            ws.append(wn[wi][rows])
            chs.append(child_int)
    results = []
    if Ys:
        lrt_out, _ = engine.lrt_poisson_batch(ctx, Ys, ws, chs, want_x=False)
        for LR, dof, on_bound, p_val, _h0, _h1, _it, status in lrt_out:
            if status != 0 or not np.isfinite(LR):
                print("\nFAILED POISSON OPTIMIZATION\n")
                results.append((np.nan, np.nan, np.nan, np.nan))
            else:
                results.append((float(LR), int(dof), int(on_bound), float(p_val)))
    out = {}
    header_str = _tsv_header(w_maxs)
    for k, it in enumerate(items):
        model_str = ""
        res_i = []
        for w_max, r in zip(w_maxs, results[k * len(w_maxs):(k + 1) * len(w_maxs)]):
            LR, dof, on_bound, p_val = r[0], r[1], r[2], r[3]
            hyp = int(p_val < 0.05)
            model_str += f"\t{LR:0>5.3f}\t{dof}\t{p_val}\tH{hyp}"
            res_i.append((LR, dof, on_bound, p_val))
        model_str = f"{it['n_kept']}\t{it['FN']:.4f}\t{it['FP']:.4f}{model_str}"
        with open(out_files[it["i"]], "w") as f_out:
            f_out.write(f"{header_str}\n{model_str}")
        out[it["i"]] = res_i
    return out


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("vcf", type=str, help="SNP file in vcf format.")
    parser.add_argument("tree", type=str, help="Tree file in newick format.")
    parser.add_argument("-o", "--output", type=str, default="",
                        help="Output file. Default = <VCF_FILE>.poissonTree_LRT.tsv.")
    parser.add_argument("-excl", "--exclude", type=str, default="",
                        help="Regex pattern for samples to exclude from LRT test. Default = none.")
    parser.add_argument("-incl", "--include", type=str, default="",
                        help="Regex pattern for samples to include from LRT test. Default = all.")
    parser.add_argument("-w", "--w_max", type=float, nargs="+",
                        default=[100, 200, 300, 400, 500, 600, 700, 800, 900, 1000],
                        help="Maximum weight value. Defaut = 100, 200, ..., 1000")
    parser.add_argument("-FN", "--FN_rate", type=float, default=None,
                        help="False negative rate. Default = inferred from .log for CellPhy/infSCITE.")
    parser.add_argument("-FP", "--FP_rate", type=float, default=None,
                        help="False positive rate. Default = inferred from .log for CellPhy/infSCITE.")
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    if not args.output:
        args.output = args.vcf + ".poissonTree_LRT.tsv"
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # launched with torchrun: one process per GPU, SNV rows sharded over the ranks
        from . import sharding
        sharding.init_distributed()
    run_poisson_tree_test(vcf_file=args.vcf, tree_file=args.tree, out_file=args.output, w_maxs=args.w_max,
                          exclude=args.exclude, include=args.include, FN_in=args.FN_rate, FP_in=args.FP_rate)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
