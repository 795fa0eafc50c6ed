"""Synthetic coalescent-shaped PT-test inputs (tests and bench.py).

The shapes follow what CellCoal produces for the reference's example_data: a
Kingman coalescent over the tumour cells plus an outgroup, SNVs thrown on
branches proportional to their length, then false negatives (allelic dropout),
false positives and missing calls applied per cell.
"""
import numpy as np


class SynthTree:
    """Binary tree over ``n`` cells. Node ids: leaves 0..n-1, internal n..2n-2
    (2n-2 is the root).  ``lo/hi`` give each node's clade as a half-open interval
    of DFS leaf ranks and ``leaf_rank[c]`` the rank of cell c."""

    def __init__(self, left, right, height, n):
        self.n = n
        self.left, self.right, self.height = left, right, height
        n_nodes = 2 * n - 1
        self.parent = np.full(n_nodes, -1, dtype=np.int64)
        for v in range(n, n_nodes):
            self.parent[left[v - n]] = v
            self.parent[right[v - n]] = v
        # DFS ranks (iterative)
        self.lo = np.zeros(n_nodes, dtype=np.int64)
        self.hi = np.zeros(n_nodes, dtype=np.int64)
        self.leaf_rank = np.zeros(n, dtype=np.int64)
        rank = 0
        stack = [(n_nodes - 1, 0)]
        while stack:
            v, state = stack.pop()
            if v < n:
                self.lo[v], self.hi[v] = rank, rank + 1
                self.leaf_rank[v] = rank
                rank += 1
                continue
            if state == 0:
                self.lo[v] = rank
                stack.append((v, 1))
                stack.append((right[v - n], 0))
                stack.append((left[v - n], 0))
            else:
                self.hi[v] = rank
        self.branch_len = np.zeros(n_nodes)
        for v in range(n_nodes - 1):
            self.branch_len[v] = height[self.parent[v]] - height[v]

    def clade_matrix(self, order=None):
        """0/1 matrix [len(order)+1, n]: one row per node in ``order`` (default all
        nodes, root last) plus the final all-zero row of map_mutations_gt."""
        n = self.n
        if order is None:
            order = np.arange(2 * n - 1)
        S = np.zeros((len(order) + 1, n), dtype=np.uint8)
        inv = np.empty(n, dtype=np.int64)
        inv[self.leaf_rank] = np.arange(n)  # rank -> cell
        for r, v in enumerate(order):
            S[r, inv[self.lo[v]:self.hi[v]]] = 1
        return S

    def newick(self, names, outgroup=None, scale=1.0):
        n = self.n

        def rec(v):
            if v < n:
                return names[v]
            a, b = self.left[v - n], self.right[v - n]
            return "(%s:%.6f,%s:%.6f)" % (rec(a), max(self.branch_len[a] * scale, 1e-6), rec(b),
                                          max(self.branch_len[b] * scale, 1e-6))

        import sys
        sys.setrecursionlimit(max(10000, 4 * n))
        body = rec(2 * n - 2)
        if outgroup is not None:
            return "(%s:%.6f,%s:%.6f);" % (body, 0.01, outgroup, 0.3)
        return body + ";"


def coalescent_tree(n, rng):
    """Kingman coalescent: merge uniformly random pairs, exponential waiting times."""
    left = np.zeros(n - 1, dtype=np.int64)
    right = np.zeros(n - 1, dtype=np.int64)
    height = np.zeros(2 * n - 1)
    active = list(range(n))
    t = 0.0
    for k in range(n - 1):
        m = len(active)
        t += rng.exponential(2.0 / (m * (m - 1)))
        i, j = rng.choice(m, size=2, replace=False)
        a, b = active[i], active[j]
        v = n + k
        left[k], right[k], height[v] = a, b, t
        for idx in sorted((i, j), reverse=True):
            active.pop(idx)
        active.append(v)
    return SynthTree(left, right, height, n)


def throw_mutations(tree, n_snv, rng, rate_mult=None):
    """Branch (node id below the branch) carrying each SNV, proportional to length."""
    w = tree.branch_len.copy()
    if rate_mult is not None:
        w = w * rate_mult
    w[-1] = 0.0
    return rng.choice(len(w), size=n_snv, p=w / w.sum())


def observe(tree, snv_branch, FN, FP, MS, rng):
    """Genotype codes [n_snv, n] with errors: 0 wt, 1 het, 3 missing."""
    n = tree.n
    ranks = tree.leaf_rank[None, :]
    lo = tree.lo[snv_branch][:, None]
    hi = tree.hi[snv_branch][:, None]
    truth = (ranks >= lo) & (ranks < hi)
    u = rng.random((len(snv_branch), n))
    called = np.where(truth, u >= FN, u < FP)
    miss = rng.random((len(snv_branch), n)) < MS
    return np.where(miss, 3, called.astype(np.uint8)).astype(np.uint8)


def observe_packed_torch(lo, hi, leaf_rank, n_snv, FN, FP, MS, device, seed=0, chunk=8192, rate=None, extra_cols=0):
    """Same model, generated on the GPU straight into the packed 2-bit layout
    (bench-sized inputs: 10^10 entries never exist on the host).  ``lo/hi`` are the
    per-branch clade intervals (torch int64 on ``device``), ``rate`` optional branch
    sampling weights.  Returns (packed uint8 [n_snv, pitch], pitch, row_keep uint8)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = leaf_rank.numel()
    pitch = ((n + extra_cols + 3) // 4 + 15) // 16 * 16
    packed = torch.zeros((n_snv, pitch), dtype=torch.uint8, device=device)
    keep = torch.zeros(n_snv, dtype=torch.uint8, device=device)
    if rate is None:
        rate = torch.ones(lo.numel(), device=device)
    ranks = leaf_rank.to(device)[None, :]
    shifts = torch.tensor([0, 2, 4, 6], dtype=torch.uint8, device=device)
    for s in range(0, n_snv, chunk):
        m = min(chunk, n_snv - s)
        br = torch.multinomial(rate, m, replacement=True, generator=g)
        truth = (ranks >= lo[br][:, None]) & (ranks < hi[br][:, None])
        u = torch.rand((m, n), device=device, generator=g)
        called = torch.where(truth, u >= FN, u < FP)
        miss = torch.rand((m, n), device=device, generator=g) < MS
        codes = torch.where(miss, torch.full_like(called, 3, dtype=torch.uint8), called.to(torch.uint8))
        keep[s:s + m] = ((codes == 1).any(dim=1)).to(torch.uint8)
        full = torch.zeros((m, pitch * 4), dtype=torch.uint8, device=device)
        full[:, :n] = codes
        q = full.view(m, pitch, 4)
        packed[s:s + m] = (q << shifts).sum(dim=2, dtype=torch.int32).to(torch.uint8)
    return packed, pitch, keep
