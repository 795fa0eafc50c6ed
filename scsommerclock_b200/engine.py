"""Host-side handles over the C ABI: context, placement plan, packing helpers.

Only plumbing lives here (numpy packing, pointer passing, stream selection);
every number the PT test reports is produced by the CUDA library.
"""
import ctypes as C

import time

import numpy as np

from . import _cabi
from ._cabi import check, lib, ptr


# ------------------------------------------------------------------ packing
def pack_genotypes(muts):
    """{0,1,2,NaN} SNV x cell matrix (the reference's ``muts`` values,
    run_PT_test.py:156-167) -> 2-bit codes, 4 cells per byte, row pitch padded to
    16 bytes.  Returns (packed uint8 [n_snv, pitch], pitch)."""
    muts = np.asarray(muts)
    n_snv, n_cells = muts.shape
    if muts.dtype.kind == "f":
        codes = np.where(np.isnan(muts), _cabi.GT_MISSING, np.nan_to_num(muts, nan=0.0)).astype(np.uint8)
    else:
        codes = muts.astype(np.uint8)
    if codes.size and codes.max() > 3:
        raise ValueError("genotype codes must be in {0,1,2,3}")
    pitch = ((n_cells + 3) // 4 + 15) // 16 * 16
    padded = np.zeros((n_snv, pitch * 4), dtype=np.uint8)
    padded[:, :n_cells] = codes
    q = padded.reshape(n_snv, pitch, 4)
    packed = (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)
    return np.ascontiguousarray(packed), pitch


def unpack_genotypes(packed, n_cells):
    """Inverse of pack_genotypes -> uint8 codes [n_snv, n_cells]."""
    packed = np.asarray(packed, dtype=np.uint8)
    n_snv = packed.shape[0]
    out = np.empty((n_snv, packed.shape[1] * 4), dtype=np.uint8)
    for i in range(4):
        out[:, i::4] = (packed >> (2 * i)) & 3
    return out[:, :n_cells]


def pack_clades(S):
    """0/1 clade matrix [n_rows, n_cells] (S of run_PT_test.py:389-397) -> uint32
    bit mask [n_rows, words], bit c%32 of word c//32."""
    S = np.asarray(S).astype(bool)
    n_rows, n_cells = S.shape
    words = (n_cells + 31) // 32
    padded = np.zeros((n_rows, words * 32), dtype=np.uint8)
    padded[:, :n_cells] = S
    bits = np.packbits(padded.reshape(n_rows, words, 32), axis=2, bitorder="little")
    return np.ascontiguousarray(bits.view("<u4").reshape(n_rows, words)), words


def unpack_clades(bits, n_cells):
    bits = np.ascontiguousarray(bits, dtype="<u4")
    n_rows = bits.shape[0]
    u8 = bits.view(np.uint8).reshape(n_rows, -1)
    return np.unpackbits(u8, axis=1, bitorder="little")[:, :n_cells].astype(bool)


# ------------------------------------------------------------------ handles
def clade_intervals(bits, n_cells):
    """scs_clade_intervals (host only): (laminar, perm [n_cells] uint16, lo [n_rows], hi [n_rows])."""
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    n_rows, words = bits.shape
    perm = np.zeros(n_cells, dtype=np.uint16)
    lohi = np.zeros(n_rows, dtype=np.uint32)
    lam = C.c_int32()
    check(lib().scs_clade_intervals(ptr(bits), int(words), int(n_rows), int(n_cells), ptr(perm), ptr(lohi), C.byref(lam)))
    return bool(lam.value), perm, (lohi & 0xFFFF).astype(np.int64), (lohi >> 16).astype(np.int64)


def _stream_handle(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)  # torch.cuda.Stream


def _dev_ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())  # torch tensor on the device


class Context:
    """One per process/GPU (scs_ctx)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib().scs_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        info = (C.c_int64 * 6)()
        check(lib().scs_ctx_info(self._h, info, 6))
        self.sm_count, self.cc_major, self.cc_minor, self.total_mem_mib, self.l2_bytes = [int(v) for v in info[:5]]

    def close(self):
        if self._h:
            lib().scs_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Placement:
    """Device-resident placement plan (scs_placement) for one problem shape."""

    def __init__(self, ctx, n_snv, n_cells, n_rows):
        """``n_snv``: an int (one dataset) or a sequence (a batch of datasets -- simulation
        replicates -- over the same number of cells, placed block-diagonally in one launch)."""
        self.ctx = ctx
        self.group_snvs = np.atleast_1d(np.asarray(n_snv, dtype=np.int64)).copy()
        self.n_groups = int(self.group_snvs.size)
        self.n_snv, self.n_cells, self.n_rows = int(self.group_snvs.sum()), int(n_cells), int(n_rows)
        h = C.c_void_p()
        check(lib().scs_placement_create_batch(ctx._h, self.n_groups, ptr(self.group_snvs), self.n_cells, self.n_rows, C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            lib().scs_placement_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        v = (C.c_int64 * 21)()
        check(lib().scs_placement_info(self._h, v, 21))
        keys = ("M_pad", "N_pad", "K_pad", "tasks", "run_len", "superblock_width", "grid", "launches", "runs", "smem_bytes",
                "mode", "k_chunks", "cta_pair", "groups", "pass2_mode", "algo", "interval_grid", "interval_threads",
                "interval_smem_bytes", "interval_runs", "interval_variant")
        return dict(zip(keys, (int(x) for x in v)))

    # host buffers (numpy)
    def set_genotypes_host(self, packed, pitch, row_keep=None, stream=None):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        assert packed.shape == (self.n_snv, pitch)
        if row_keep is not None:
            row_keep = np.ascontiguousarray(row_keep, dtype=np.uint8)
            assert row_keep.shape == (self.n_snv,)
        check(lib().scs_placement_set_genotypes_host(self._h, ptr(packed), int(pitch), ptr(row_keep), _stream_handle(stream)))

    def set_clades_host(self, bits, words, stream=None):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        assert bits.shape == (self.n_groups * self.n_rows, words)
        check(lib().scs_placement_set_clades_host(self._h, ptr(bits), int(words), _stream_handle(stream)))

    # device buffers (torch tensors or raw addresses)
    def set_genotypes(self, d_packed, pitch, d_row_keep=None, stream=None):
        check(lib().scs_placement_set_genotypes(self._h, _dev_ptr(d_packed), int(pitch), _dev_ptr(d_row_keep), _stream_handle(stream)))

    def bind_genotypes(self, d_packed, pitch, d_row_keep=None, stream=None):
        """set_genotypes without a copy on the interval path: ``d_packed`` must stay alive and
        unchanged until the run that uses it has finished (call after the clades are set)."""
        check(lib().scs_placement_bind_genotypes(self._h, _dev_ptr(d_packed), int(pitch), _dev_ptr(d_row_keep), _stream_handle(stream)))

    def prepare(self, d_packed, pitch, col_active, filter_rows, accumulate=False, stream=None):
        """scs_placement_prepare: row filter + per-cell missing counts + the plan's genotype set-up in one
        asynchronous call (one pass over ``d_packed`` on the interval path).  ``d_packed`` must stay alive
        and unchanged until the run has finished; the numbers come back through ``prepare_result``."""
        col_active = np.ascontiguousarray(col_active, dtype=np.uint8)
        assert col_active.shape == (self.n_cells,)
        check(lib().scs_placement_prepare(self._h, _dev_ptr(d_packed), int(pitch), ptr(col_active), int(bool(filter_rows)),
                                          int(bool(accumulate)), _stream_handle(stream)))

    def prepare_result(self, stream=None):
        """(kept SNVs, missing fraction per cell) of the prepare call(s) so far; waits for ``stream``."""
        n_kept = C.c_int64()
        miss = np.zeros(self.n_cells, dtype=np.float64)
        check(lib().scs_placement_prepare_result(self._h, C.byref(n_kept), ptr(miss), _stream_handle(stream)))
        return int(n_kept.value), miss

    def set_clades(self, d_bits, words, stream=None):
        check(lib().scs_placement_set_clades(self._h, _dev_ptr(d_bits), int(words), _stream_handle(stream)))

    def set_trees(self, d_bits, words, d_children, d_info, stream=None):
        """scs_placement_set_trees: clade rows given as trees (device copies of scs_tree_parse_batch's outputs); batches of
        up to 512 cells then take the interval path for small trees instead of the GEMM."""
        check(lib().scs_placement_set_trees(self._h, _dev_ptr(d_bits), int(words), _dev_ptr(d_children), _dev_ptr(d_info), _stream_handle(stream)))

    def _rates(self, FP, FN):
        fp = np.ascontiguousarray(np.broadcast_to(np.asarray(FP, dtype=np.float64), (self.n_groups,)))
        fn = np.ascontiguousarray(np.broadcast_to(np.asarray(FN, dtype=np.float64), (self.n_groups,)))
        return fp, fn

    def run(self, FP, FN, d_out=None, stream=None):
        """Asynchronous; result in ``d_out`` (device, float64[n_groups * n_rows]) or in the plan.
        ``FP`` / ``FN``: scalars or one value per group."""
        fp, fn = self._rates(FP, FN)
        check(lib().scs_placement_run_batch(self._h, ptr(fp), ptr(fn), _dev_ptr(d_out), _stream_handle(stream)))

    def run_host(self, FP, FN, stream=None):
        out = np.empty(self.n_groups * self.n_rows, dtype=np.float64)
        if self.n_groups == 1:
            check(lib().scs_placement_run_host(self._h, float(np.ravel(FP)[0]), float(np.ravel(FN)[0]), ptr(out), _stream_handle(stream)))
            return out
        import torch
        d = torch.empty(out.size, dtype=torch.float64, device=torch.device("cuda", self.ctx.device))
        self.run(FP, FN, d, stream=stream)
        return d.cpu().numpy().reshape(self.n_groups, self.n_rows)

    def last_times(self):
        """Device ms of (GEMM pass 1, merge, GEMM pass 2, reduce) of the last run."""
        ms = (C.c_float * 4)()
        check(lib().scs_placement_last_times(self._h, ms))
        return [float(v) for v in ms]

    def prepare_time(self):
        """Device ms of the last ``prepare`` call's kernels."""
        ms = C.c_float()
        check(lib().scs_placement_prepare_time(self._h, C.byref(ms)))
        return float(ms.value)

    def device_result_ptr(self):
        return int(lib().scs_placement_device_result(self._h))

    def debug_counts(self, stream=None):
        out = np.empty((self.n_snv, self.n_rows, 2), dtype=np.int32)
        check(lib().scs_placement_debug_counts(self._h, ptr(out), _stream_handle(stream)))
        return out


def map_mutations(ctx, packed, pitch, clade_bits, words, n_cells, FP, FN, row_keep=None):
    """One-shot host call (scs_map_mutations): soft branch weights float64[n_rows]."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    clade_bits = np.ascontiguousarray(clade_bits, dtype=np.uint32)
    n_snv, n_rows = packed.shape[0], clade_bits.shape[0]
    if row_keep is not None:
        row_keep = np.ascontiguousarray(row_keep, dtype=np.uint8)
    out = np.empty(n_rows, dtype=np.float64)
    check(lib().scs_map_mutations(ctx._h, ptr(packed), int(pitch), ptr(row_keep), n_snv, int(n_cells),
                                  ptr(clade_bits), int(words), n_rows, float(FP), float(FN), ptr(out)))
    return out


def place_percell(ctx, d_gt, pitch, n_snv, n_cells, d_row_keep, clade_bits, words, FP, FN_cells, d_soft, stream=None):
    """scs_place_percell: map_mutations_gt with one false-negative rate per cell (the pd.Series branch of
    run_PT_test.py:413-418).  ``d_gt`` / ``d_row_keep`` / ``d_soft`` are device buffers (torch tensors or raw
    pointers; ``d_row_keep`` may be None), ``clade_bits`` and ``FN_cells`` (one rate per matrix column) host arrays.
    Returns the device time of the kernels in ms; ``d_soft`` holds the soft weights when the call returns."""
    clade_bits = np.ascontiguousarray(clade_bits, dtype=np.uint32)
    FN_cells = np.ascontiguousarray(FN_cells, dtype=np.float64)
    if FN_cells.shape != (int(n_cells),):
        raise ValueError("FN_cells must hold one rate per matrix column (%d), got shape %r" % (n_cells, FN_cells.shape))
    if stream is None and hasattr(d_gt, "device"):
        import torch
        stream = torch.cuda.current_stream(d_gt.device)
    ms = C.c_float()
    check(lib().scs_place_percell(ctx._h, _dev_ptr(d_gt), int(pitch), int(n_snv), int(n_cells),
                                  _dev_ptr(d_row_keep) if d_row_keep is not None else None, ptr(clade_bits), int(words),
                                  int(clade_bits.shape[0]), float(FP), ptr(FN_cells), _dev_ptr(d_soft), C.byref(ms),
                                  _stream_handle(stream)))
    return float(ms.value)


class DecodedVcf:
    """Device-resident result of scs_vcf_decode (get_mut_df, run_PT_test.py:27-207)."""

    def __init__(self, ctx, text, n_samples, incl_idx):
        self.ctx = ctx
        incl_idx = np.ascontiguousarray(incl_idx, dtype=np.int32)
        h = C.c_void_p()
        text = bytes(text)
        check(lib().scs_vcf_decode(ctx._h, C.cast(C.c_char_p(text), C.c_void_p), len(text), int(n_samples),
                                   ptr(incl_idx), int(incl_idx.size), C.byref(h)))
        self._h = h
        info = (C.c_int64 * 6)()
        check(lib().scs_vcf_info(self._h, info, 6))
        self.n_kept, self.n_incl, self.pitch, self.n_lines, self.n_noalt, self.n_filtered = [int(v) for v in info]

    def device_ptr(self):
        return int(lib().scs_vcf_device_genotypes(self._h) or 0)

    def copy(self):
        packed = np.zeros((self.n_kept, self.pitch), dtype=np.uint8)
        pos = np.zeros((self.n_kept, 2), dtype=np.int64)
        check(lib().scs_vcf_copy(self._h, ptr(packed), ptr(pos)))
        return packed, pos

    def close(self):
        if self._h:
            lib().scs_vcf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gt_reduce(ctx, d_gt, pitch, n_snv, n_cells, col_active, filter_rows, d_row_keep, stream=None):
    """scs_gt_reduce: fills the device buffer ``d_row_keep`` and returns
    (n_kept, missing fraction per cell).  Ordered on ``stream`` (default: torch's current stream when the
    buffers are torch tensors, else the legacy default stream)."""
    if stream is None and hasattr(d_gt, "device"):
        import torch
        stream = torch.cuda.current_stream(d_gt.device)
    col_active = np.ascontiguousarray(col_active, dtype=np.uint8)
    n_kept = C.c_int64()
    miss = np.zeros(n_cells, dtype=np.float64)
    check(lib().scs_gt_reduce(ctx._h, _dev_ptr(d_gt), int(pitch), int(n_snv), int(n_cells), ptr(col_active),
                              int(bool(filter_rows)), _dev_ptr(d_row_keep), C.byref(n_kept), ptr(miss), _stream_handle(stream)))
    return int(n_kept.value), miss


def branch_weights(ctx, node_bits, words, n_cells, missing_rate, FP, FN, w_maxs):
    """scs_branch_weights: (weights [n_w, n_nodes], weights_norm [n_w, n_nodes], root row)."""
    node_bits = np.ascontiguousarray(node_bits, dtype=np.uint32)
    n_nodes = node_bits.shape[0]
    missing_rate = np.ascontiguousarray(missing_rate, dtype=np.float64)
    w_maxs = np.ascontiguousarray(w_maxs, dtype=np.float64)
    weights = np.zeros((w_maxs.size, n_nodes))
    weights_norm = np.zeros((w_maxs.size, n_nodes))
    root = C.c_int32()
    check(lib().scs_branch_weights(ctx._h, ptr(node_bits), int(words), n_nodes, int(n_cells), ptr(missing_rate),
                                   float(FP), float(FN), ptr(w_maxs), int(w_maxs.size), ptr(weights), ptr(weights_norm),
                                   C.byref(root)))
    return weights, weights_norm, int(root.value)


def lrt_poisson_batch(ctx, Ys, ws, child_ints, want_x=False):
    """scs_lrt_poisson_batch over a list of problems.  Returns (out [n, 8], x list or None);
    out columns: LR, dof, on_bound, p_value, ll_H0, ll_H1, iterations, status."""
    n = len(Ys)
    sizes = np.array([len(y) for y in Ys], dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    Y = np.ascontiguousarray(np.concatenate(Ys), dtype=np.float64)
    w = np.ascontiguousarray(np.concatenate(ws), dtype=np.float64)
    ch = np.ascontiguousarray(np.concatenate(child_ints), dtype=np.int32)
    assert Y.size == w.size == ch.size == off[-1]
    out = np.zeros((n, 8), dtype=np.float64)
    x = np.zeros(Y.size, dtype=np.float64) if want_x else None
    check(lib().scs_lrt_poisson_batch(ctx._h, n, ptr(off), ptr(Y), ptr(w), ptr(ch), ptr(out), ptr(x)))
    xs = [x[off[i]:off[i + 1]] for i in range(n)] if want_x else None
    return out, xs


class LrtPlan:
    """Reusable device-resident LRT batch (scs_lrt_plan): fixed topologies, weights
    set once, counts fed per run from the host or straight from device memory."""

    def __init__(self, ctx, child_ints):
        self.ctx = ctx
        sizes = np.array([len(c) for c in child_ints], dtype=np.int64)
        self.off = np.zeros(len(child_ints) + 1, dtype=np.int64)
        self.off[1:] = np.cumsum(sizes)
        self.n_prob = len(child_ints)
        ch = np.ascontiguousarray(np.concatenate(child_ints), dtype=np.int32)
        h = C.c_void_p()
        check(lib().scs_lrt_plan_create(ctx._h, self.n_prob, ptr(self.off), ptr(ch), C.byref(h)))
        self._h = h
        self.launches = 0

    def set_weights(self, ws):
        w = np.ascontiguousarray(np.concatenate(ws), dtype=np.float64)
        assert w.size == self.off[-1]
        check(lib().scs_lrt_plan_set_weights(self._h, ptr(w)))

    def set_gather(self, index, src_len):
        index = np.ascontiguousarray(index, dtype=np.int32)
        assert index.size == self.off[-1]
        check(lib().scs_lrt_plan_set_gather(self._h, ptr(index), int(src_len)))

    def run(self, d_Y, d_out=None, stream=None):
        check(lib().scs_lrt_plan_run(self._h, _dev_ptr(d_Y), _dev_ptr(d_out), None, _stream_handle(stream)))
        self.launches += 1

    def run_host(self, Y, want_x=False, stream=None):
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        out = np.zeros((self.n_prob, 8))
        x = np.zeros(int(self.off[-1])) if want_x else None
        check(lib().scs_lrt_plan_run_host(self._h, ptr(Y), ptr(out), ptr(x), _stream_handle(stream)))
        self.launches += 1
        return out, x

    def close(self):
        if self._h:
            lib().scs_lrt_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StreamingPlacement:
    """Placement of a host-resident packed genotype matrix in SNV-row chunks.

    The soft branch weights are additive over SNV rows, so a dataset that sits in (pinned)
    host memory is walked in chunks: the H2D copy of chunk c+1 runs on a side stream while
    chunk c goes through scs_placement_prepare (row filter, missing counts, cell-order pass: one
    read of the chunk) and the placement kernel on the compute stream.  One chunk-sized plan is
    reused (plus a smaller one for a short last chunk), device memory stays bounded by the chunk
    size, nothing is read back per chunk -- the counters accumulate on the device -- and the copy
    disappears behind the kernels.  Returns the same numbers a single plan gives (to summation order)."""

    CHUNK_BYTES = 96 << 20      # default chunk size: large enough for full-rate H2D copies, small enough that the first,
                                # exposed copy is a few per cent of the step

    def __init__(self, ctx, n_snv, n_cells, n_rows, n_chunks=None, pitch=None):
        import torch
        self.ctx, self.n_snv, self.n_cells, self.n_rows = ctx, int(n_snv), int(n_cells), int(n_rows)
        if n_chunks is None:
            pitch = pitch or ((n_cells + 3) // 4 + 15) // 16 * 16
            n_chunks = int(max(1, min(64, (self.n_snv * pitch) // self.CHUNK_BYTES)))
        rows = (self.n_snv + n_chunks - 1) // n_chunks
        self.chunk_rows = min(self.n_snv, max(256, (rows + 255) // 256 * 256))
        self.n_chunks = (self.n_snv + self.chunk_rows - 1) // self.chunk_rows
        self.tail_rows = self.n_snv - (self.n_chunks - 1) * self.chunk_rows
        self.plan = Placement(ctx, self.chunk_rows, n_cells, n_rows)
        self.plan_tail = Placement(ctx, self.tail_rows, n_cells, n_rows) if self.tail_rows != self.chunk_rows else None
        self.dev = torch.device("cuda", ctx.device)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self._bufs = None
        self.y_chunk = torch.zeros(n_rows, dtype=torch.float64, device=self.dev)

    def set_clades_host(self, bits, words):
        self.plan.set_clades_host(bits, words)
        if self.plan_tail is not None:
            self.plan_tail.set_clades_host(bits, words)

    def close(self):
        self.plan.close()
        if self.plan_tail is not None:
            self.plan_tail.close()

    def run(self, host_packed, pitch, col_active, filter_rows, FP, FN, y_dev):
        """``host_packed``: torch uint8 [n_snv, pitch] in pinned host memory.  Accumulates the soft
        weights into ``y_dev`` (zeroed here) and returns (kept SNVs, per-cell missing fraction)."""
        import torch
        cur = torch.cuda.current_stream(self.dev)
        if self._bufs is None or self._bufs[0].shape[1] != pitch:
            self._bufs = [torch.empty((self.chunk_rows, pitch), dtype=torch.uint8, device=self.dev) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
            for e in self._free:
                e.record(cur)
        y_dev.zero_()

        def enqueue_copy(c):
            b = c & 1
            r0 = c * self.chunk_rows
            r1 = min(self.n_snv, r0 + self.chunk_rows)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self._free[b])          # the previous user of this buffer is done
                self._bufs[b][: r1 - r0].copy_(host_packed[r0:r1], non_blocking=True)
                self._ready[b].record(self.copy_stream)

        enqueue_copy(0)
        for c in range(self.n_chunks):
            b = c & 1
            if c + 1 < self.n_chunks:
                enqueue_copy(c + 1)
            cur.wait_event(self._ready[b])
            tail = c == self.n_chunks - 1 and self.plan_tail is not None
            plan = self.plan_tail if tail else self.plan
            plan.prepare(self._bufs[b], pitch, col_active, filter_rows, accumulate=(c > 0 and not tail), stream=cur)   # the buffer outlives the run (_free[b])
            plan.run(FP, FN, self.y_chunk, stream=cur)
            y_dev.add_(self.y_chunk)
            self._free[b].record(cur)
        n_kept, miss = self.plan.prepare_result(stream=cur) if (self.plan_tail is None or self.n_chunks > 1) else (0, np.zeros(len(col_active)))
        if self.plan_tail is not None:
            k2, m2 = self.plan_tail.prepare_result(stream=cur)
            miss = (miss * n_kept + m2 * k2) / max(n_kept + k2, 1)
            n_kept += k2
        return n_kept, miss


# ------------------------------------------------------------------ replicate sweeps
def _concat_strings(items):
    """bytes buffer + int64 offsets [n + 1] of a list of str / bytes."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in items]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in bs])
    return b"".join(bs), off


def tree_parse_batch(newicks, sample_names, n_rows, strip_tags=False, outgroup="healthycell", threads=0):
    """scs_tree_parse_batch (host only): read_tree + the clade rows for many Newick texts at once.
    Returns (bits uint32 [n, n_rows, words], words, children int32 [n, n_rows, 2], info int32 [n, 4])."""
    n = len(newicks)
    text, off = _concat_strings(newicks)
    names, noff = _concat_strings(sample_names)
    words = (len(sample_names) + 31) // 32
    bits = np.empty((n, n_rows, words), dtype=np.uint32)
    children = np.empty((n, n_rows, 2), dtype=np.int32)
    info = np.zeros((n, 4), dtype=np.int32)
    check(lib().scs_tree_parse_batch(n, text, ptr(off), int(bool(strip_tags)), len(sample_names), names, ptr(noff), outgroup.encode(),
                                     int(n_rows), int(words), ptr(bits), ptr(children), ptr(info), int(threads)))
    return bits, words, children, info


def format_pt_rows(out, n_kept, FN, FP):
    """scs_format_pt_rows: the TSV model line of every replicate (list of str)."""
    out = np.ascontiguousarray(out, dtype=np.float64)
    n_groups, n_w = out.shape[0], out.shape[1]
    n_kept = np.ascontiguousarray(n_kept, dtype=np.int64)
    FN = np.ascontiguousarray(np.broadcast_to(np.asarray(FN, dtype=np.float64), (n_groups,)))
    FP = np.ascontiguousarray(np.broadcast_to(np.asarray(FP, dtype=np.float64), (n_groups,)))
    cap = n_groups * (40 + 64 * n_w)
    buf = C.create_string_buffer(cap)
    row_off = np.zeros(n_groups + 1, dtype=np.int64)
    check(lib().scs_format_pt_rows(n_groups, n_w, ptr(out), ptr(n_kept), ptr(FN), ptr(FP), C.cast(buf, C.c_void_p), cap, ptr(row_off)))
    raw = buf.raw
    return [raw[row_off[g]:row_off[g + 1]].decode() for g in range(n_groups)]


class ReplicateSweep:
    """A batch of (genotype matrix, tree) pairs over the same sample columns -- a simulation sweep -- through the whole PT
    test with a fixed number of calls, whatever the number of pairs:

        host threads   scs_tree_parse_batch      read_tree + clade rows of every tree (native, ~30 us per 100-leaf tree)
        device         scs_gt_reduce_groups      row filter + per-cell missing counts of every pair
                       scs_placement_* (batch)   all pairs placed block-diagonally, one launch per pass
                       scs_sweep_model           branch weights + LRT branch order of every pair, into the LRT plan
                       scs_lrt_plan_run          all (pair, w_max) clock tests in one launch
        host           scs_format_pt_rows        the TSV model lines

    Nothing is read back between the upload of the matrices / masks and the download of the results.  Every tree must
    hold all non-outgroup columns as leaves (the usual case); pairs whose tree does not are flagged in ``status`` and left
    to the per-pair path."""

    def __init__(self, ctx, group_snvs, columns, w_maxs, outgroup="healthycell"):
        import torch
        self.ctx = ctx
        self.dev = torch.device("cuda", ctx.device)
        self.group_snvs = np.ascontiguousarray(group_snvs, dtype=np.int64)
        self.G = int(self.group_snvs.size)
        self.columns = list(columns)
        self.outgroup = outgroup
        self.n_cols = len(self.columns)
        self.active = np.array([c != outgroup for c in self.columns], dtype=np.uint8)
        self.has_outg = outgroup in self.columns
        n_active = int(self.active.sum())
        if n_active < 2:
            raise ValueError("a sweep needs at least two cells besides the outgroup")
        self.n_rows = 2 * n_active                  # max(2 n - 1, nodes) + 1 with nodes = 2 n - 1 (:389, :397)
        self.n_br = 2 * (n_active - 1)
        self.w_maxs = np.ascontiguousarray(w_maxs, dtype=np.float64)
        self.n_w = int(self.w_maxs.size)
        self.words = (self.n_cols + 31) // 32
        self.pitch = ((self.n_cols + 3) // 4 + 15) // 16 * 16
        self.total = int(self.group_snvs.sum())
        self.place = Placement(ctx, self.group_snvs, self.n_cols, self.n_rows)
        h = C.c_void_p()
        check(lib().scs_lrt_plan_create_uniform(ctx._h, self.G * self.n_w, self.n_br, C.byref(h)))
        self._lrt = h
        G, d = self.G, self.dev
        row0 = np.zeros(G + 1, dtype=np.int64)
        row0[1:] = np.cumsum(self.group_snvs)
        self.d_row0 = torch.as_tensor(row0, device=d)
        self.d_keep = torch.empty(max(self.total, 1), dtype=torch.uint8, device=d)
        self.d_nkept = torch.zeros(G, dtype=torch.int64, device=d)
        self.d_missing = torch.zeros((G, self.n_cols), dtype=torch.int32, device=d)
        self.d_bits = torch.empty((G, self.n_rows, self.words), dtype=torch.int32, device=d)
        self.d_children = torch.empty((G, self.n_rows, 2), dtype=torch.int32, device=d)
        self.d_info = torch.empty((G, 4), dtype=torch.int32, device=d)
        self.d_y = torch.zeros(G * self.n_rows, dtype=torch.float64, device=d)
        self.d_out = torch.zeros((G * self.n_w, 8), dtype=torch.float64, device=d)
        self.d_status = torch.zeros(G, dtype=torch.int32, device=d)
        amask = np.zeros((self.n_cols + 15) // 16, dtype=np.uint32)
        for c in range(self.n_cols):
            if self.active[c]:
                amask[c >> 4] |= np.uint32(1 << (2 * (c & 15)))
        self.d_amask = torch.as_tensor(amask.view(np.int32), device=d)
        self.d_work = torch.empty(int(lib().scs_sweep_model_workspace(G, self.n_rows, self.n_w)) // 8, dtype=torch.float64, device=d)
        # pinned staging for the small host <-> device transfers (asynchronous copies need page-locked memory)
        self.h_bits = torch.empty((G, self.n_rows, self.words), dtype=torch.int32, pin_memory=True)
        self.h_children = torch.empty((G, self.n_rows, 2), dtype=torch.int32, pin_memory=True)
        self.h_info = torch.empty((G, 4), dtype=torch.int32, pin_memory=True)
        self.h_out = torch.empty((G * self.n_w, 8), dtype=torch.float64, pin_memory=True)
        self.h_nkept = torch.empty(G, dtype=torch.int64, pin_memory=True)
        self.h_status = torch.empty(G, dtype=torch.int32, pin_memory=True)
        self.d_gt = None
        self.done = torch.cuda.Event()
        self.marks = None

    def close(self):
        if self._lrt:
            lib().scs_lrt_plan_destroy(self._lrt)
            self._lrt = None
        self.place.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def parse_trees(self, newicks, strip_tags=False, threads=0):
        """Host part (no device work): trees -> masks / topology in the pinned staging buffers.  ``threads`` = 0: the
        host cores divided by the ranks on this node (torchrun's LOCAL_WORLD_SIZE), at most 16."""
        n = len(newicks)
        assert n == self.G
        if threads <= 0:
            import os
            local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
            threads = max(1, min(16, (os.cpu_count() or 1) // local_world))
        text, off = _concat_strings(newicks)
        names, noff = _concat_strings(self.columns)
        check(lib().scs_tree_parse_batch(n, text, ptr(off), int(bool(strip_tags)), self.n_cols, names, ptr(noff), self.outgroup.encode(),
                                         self.n_rows, self.words, C.c_void_p(self.h_bits.data_ptr()), C.c_void_p(self.h_children.data_ptr()),
                                         C.c_void_p(self.h_info.data_ptr()), int(threads)))

    def enqueue(self, packed, FP, FN, stream=None):
        """Device part, asynchronous on ``stream`` (default: torch's current stream): ``packed`` is a torch uint8
        [sum(group_snvs), pitch] tensor on the device, or in pinned host memory (then it is copied first).  The trees
        must have been parsed (``parse_trees``).  ``done`` is recorded at the end; ``results()`` waits for it."""
        import torch
        st = stream or torch.cuda.current_stream(self.dev)
        FP = np.ascontiguousarray(np.broadcast_to(np.asarray(FP, dtype=np.float64), (self.G,)))
        FN = np.ascontiguousarray(np.broadcast_to(np.asarray(FN, dtype=np.float64), (self.G,)))
        self.FP, self.FN = FP, FN
        marks = self.marks                  # tools only: a list collects (stage, event) pairs of this call
        def mark(what):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(st)
                marks.append((what, ev))
        with torch.cuda.stream(st):
            mark("start")
            if packed.is_cuda:
                d_gt = packed
            else:
                if self.d_gt is None:
                    self.d_gt = torch.empty((self.total, self.pitch), dtype=torch.uint8, device=self.dev)
                self.d_gt.copy_(packed, non_blocking=True)
                d_gt = self.d_gt
            sh = _stream_handle(st)
            # masks / topology / node counts: uploaded by a kernel that reads the pinned staging buffers, not by the copy
            # engine, whose one queue may be busy with another sub-batch's matrix for milliseconds (scs_upload_async)
            for dst, src in ((self.d_bits, self.h_bits), (self.d_children, self.h_children), (self.d_info, self.h_info)):
                check(lib().scs_upload_async(_dev_ptr(dst), C.c_void_p(src.data_ptr()), int(src.numel() * src.element_size()), sh))
            mark("uploads")
            check(lib().scs_gt_reduce_groups(self.ctx._h, _dev_ptr(d_gt), int(self.pitch), self.G, _dev_ptr(self.d_row0), int(self.group_snvs.max()),
                                             self.n_cols, _dev_ptr(self.d_amask), int(self.has_outg), _dev_ptr(self.d_keep), _dev_ptr(self.d_nkept),
                                             _dev_ptr(self.d_missing), sh))
            mark("gt_reduce")
            self.place.set_trees(self.d_bits, self.words, self.d_children, self.d_info, stream=st)
            self.place.set_genotypes(d_gt, self.pitch, self.d_keep, stream=st)
            mark("tables")
            self.place.run(FP, FN, self.d_y, stream=st)
            mark("placement")
            check(lib().scs_sweep_model(self.ctx._h, self.G, self.n_rows, self.n_cols, int(self.words), _dev_ptr(self.d_bits), _dev_ptr(self.d_children),
                                        _dev_ptr(self.d_info), _dev_ptr(self.d_nkept), _dev_ptr(self.d_missing), ptr(FP), ptr(FN), ptr(self.w_maxs),
                                        self.n_w, _dev_ptr(self.d_y), self.n_br, self._lrt, _dev_ptr(self.d_status), _dev_ptr(self.d_work), sh))
            mark("model")
            check(lib().scs_lrt_plan_run(self._lrt, _dev_ptr(self.d_y), _dev_ptr(self.d_out), None, sh))
            mark("lrt")
            self.h_out.copy_(self.d_out, non_blocking=True)
            self.h_nkept.copy_(self.d_nkept, non_blocking=True)
            self.h_status.copy_(self.d_status, non_blocking=True)
            mark("download")
            self.done.record(st)

    def results(self):
        """(out float64 [G, n_w, 8], n_kept int64 [G], status int32 [G]) of the last ``enqueue``."""
        self.done.synchronize()
        return self.h_out.numpy().reshape(self.G, self.n_w, 8), self.h_nkept.numpy(), self.h_status.numpy()

    def rows(self):
        """The TSV model line of every pair (``results`` formatted by scs_format_pt_rows)."""
        out, n_kept, status = self.results()
        return format_pt_rows(out, n_kept, self.FN, self.FP)


def sub_batch_cuts(n_groups, n_sub):
    """Boundaries [0, ..., n_groups] of at most ``n_sub`` non-empty sub-batches: a short first one (the device idles until
    its trees are parsed and its kernels enqueued) and a short last one (its result read-back and line formatting are not
    hidden behind anything) -- weights 1, 2, 2, ..., 2, 1."""
    n_groups, n_sub = int(n_groups), max(1, min(int(n_sub), int(n_groups)))
    wts = np.array([1.0] + [2.0] * (n_sub - 2) + [1.0]) if n_sub >= 3 else np.ones(n_sub)
    edges = np.concatenate([[0.0], np.cumsum(wts)]) / wts.sum()
    return sorted(set(int(round(n_groups * e)) for e in edges))


class ReplicateSweepStream:
    """A large sweep whose packed genotype matrices sit in (pinned) HOST memory and whose trees are Newick TEXT, through
    ``ReplicateSweep`` in ``n_sub`` sub-batches so that the three resources work at the same time:

        copy engine     every sub-batch's matrix copy is queued up front on a copy stream;
        host threads    a worker thread parses the trees of sub-batch k + 1, k + 2, ... (native code, the GIL is released)
                        while the main thread enqueues sub-batch k's kernels the moment its trees are parsed and formats the
                        TSV lines of the sub-batches that have finished (native, on host threads);
        SMs             sub-batch k's kernels start when its copy has landed.

    Measured before this class existed (tools/sweep_stages.py, 1250 pairs of 100 cells x 10k SNVs in 4 sub-batches): copy
    1.8 ms + kernels 2.4 ms per sub-batch, but 1.8 ms of tree parsing and 1.9 ms of line formatting per sub-batch in series on
    the one host thread -- the device idled half the time."""

    def __init__(self, ctx, group_snvs, columns, w_maxs, n_sub=4, outgroup="healthycell"):
        import torch
        self.ctx = ctx
        self.dev = torch.device("cuda", ctx.device)
        group_snvs = np.ascontiguousarray(group_snvs, dtype=np.int64)
        G = int(group_snvs.size)
        self.cuts = sub_batch_cuts(G, n_sub)
        n_sub = len(self.cuts) - 1
        self.row_cuts = [int(group_snvs[:c].sum()) for c in self.cuts]
        self.subs = [ReplicateSweep(ctx, group_snvs[self.cuts[k]:self.cuts[k + 1]], columns, w_maxs, outgroup=outgroup) for k in range(n_sub)]
        for sub in self.subs:
            sub.d_gt = torch.empty((sub.total, sub.pitch), dtype=torch.uint8, device=self.dev)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.compute_stream = torch.cuda.Stream(device=self.dev)
        self.copied = [torch.cuda.Event() for _ in self.subs]
        self.n_sub = n_sub
        self.trace = None              # set to a list to collect (what, sub-batch, host ms) records of the next run

    def close(self):
        for sub in self.subs:
            sub.close()

    def run(self, host_packed, newicks, FP, FN, strip_tags=False):
        """``host_packed``: torch uint8 [sum(group_snvs), pitch] in pinned host memory; ``newicks``: one text per pair;
        ``FP`` / ``FN``: scalars or one value per pair.  Returns (TSV model lines, status int32 per pair)."""
        import threading
        import torch
        G = self.cuts[-1]
        FP = np.ascontiguousarray(np.broadcast_to(np.asarray(FP, dtype=np.float64), (G,)))
        FN = np.ascontiguousarray(np.broadcast_to(np.asarray(FN, dtype=np.float64), (G,)))
        cur = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(cur)
        self.compute_stream.wait_stream(cur)

        trace = self.trace
        t00 = time.perf_counter()
        if trace is not None:
            self.trace_events = []
            self.trace_t0 = torch.cuda.Event(enable_timing=True)
            self.trace_t0.record(cur)
        # every sub-batch's matrix copy is queued up front on the copy stream: the link is busy from the first microsecond.
        # (That only works because the small uploads of a sub-batch -- masks, topology, per-group constants -- do NOT go
        # through the copy engine, scs_upload_async: the device's one H2D queue serves all streams, and with copy-engine
        # uploads the first kernels started after the LAST matrix had landed.)
        with torch.cuda.stream(self.copy_stream):
            for k, sub in enumerate(self.subs):
                self.copy_stream.wait_event(sub.done)          # the previous run's kernels on this buffer (a no-op the first time)
                sub.d_gt.copy_(host_packed[self.row_cuts[k]:self.row_cuts[k + 1]], non_blocking=True)
                self.copied[k].record(self.copy_stream)
        parsed = [threading.Event() for _ in self.subs]
        failure = []

        def note(what, k):
            if trace is not None:
                trace.append((what, k, 1e3 * (time.perf_counter() - t00)))

        def parse_all():
            try:
                for k, sub in enumerate(self.subs):
                    sub.parse_trees(newicks[self.cuts[k]:self.cuts[k + 1]], strip_tags=strip_tags)
                    note("parsed", k)
                    parsed[k].set()
            except BaseException as exc:          # handed to the caller's thread
                failure.append(exc)
                for ev in parsed:
                    ev.set()

        worker = threading.Thread(target=parse_all, daemon=True)
        worker.start()
        rows, status, formatted = [], [], 0
        for k, sub in enumerate(self.subs):
            parsed[k].wait()
            if failure:
                break
            self.compute_stream.wait_event(self.copied[k])
            if trace is not None:                  # device time stamps around the sub-batch's kernels (tools/sweep_stream_trace.py)
                e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e_a.record(self.compute_stream)
            sub.enqueue(sub.d_gt, FP[self.cuts[k]:self.cuts[k + 1]], FN[self.cuts[k]:self.cuts[k + 1]], stream=self.compute_stream)
            if trace is not None:
                e_b.record(self.compute_stream)
                self.trace_events.append((k, e_a, e_b))
            note("enqueued", k)
            while formatted < k and self.subs[formatted].done.query():      # lines of sub-batches that are already back
                rows.extend(self.subs[formatted].rows())
                status.append(self.subs[formatted].h_status.numpy().copy())
                note("formatted", formatted)
                formatted += 1
        worker.join()
        if failure:
            torch.cuda.synchronize(self.dev)
            raise failure[0]
        for k in range(formatted, len(self.subs)):
            sub = self.subs[k]
            sub.done.synchronize()
            note("done", k)
            rows.extend(sub.rows())
            status.append(sub.h_status.numpy().copy())
            note("formatted", k)
        cur.wait_stream(self.compute_stream)
        return rows, np.concatenate(status)
