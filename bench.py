#!/usr/bin/env python3
"""bench.py -- the PT-test hot path on synthetic coalescent-shaped data.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        (the reference's CPU algorithm, rank 0 only)

Workload (BASELINE.json configs[3]): ONE dataset of 10,000 cells x 1,000,000 SNVs, its
SNV rows block-partitioned over the ranks (strong scaling; the 1M-SNVs-per-GPU variant
is measured beside it as `weak_scaling`), one all-reduce of the 2n branch weights,
then the clock LRT for the ten default w_max values.  One "step" is one pass of
the hot path over the resident packed genotype matrix:
    row filter + missing rates (scs_gt_reduce) -> placement -> [all-reduce] -> LRT batch -> result to host,
where the placement of a tree-shaped clade matrix is ONE kernel (interval path: per-SNV prefix
counts in shared memory, counts of a clade = one difference) and that of a general 0/1 matrix
(or with --algo gemm) is operand planes -> tcgen05 GEMM pass 1 -> merge -> pass 2 -> reduce.
`value` times the step with CUDA events, inputs already in HBM; `e2e` adds the H2D copy of the
packed matrix from pinned host memory every step; `gemm_path` is the same workload through the
GEMM, measured in the same run.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: if the environment turns NCCL's debug output on (it goes to
# stdout by default, e.g. the version banner), send it to stderr instead
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

W_MAXS = [100.0, 200.0, 300.0, 400.0, 500.0, 600.0, 700.0, 800.0, 900.0, 1000.0]
FN_TRUE, FP_TRUE, MS_TRUE = 0.10, 0.01, 0.15
METRIC = "snv_branch_placements_per_s"

def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=10000)
    ap.add_argument("--snvs", type=int, default=1000000, help="SNVs of the dataset (block-partitioned over the GPUs)")
    ap.add_argument("--no-weak-leg", action="store_true", help="N > 1: skip the extra weak-scaling measurement (--snvs rows on every rank)")
    ap.add_argument("--no-small-leg", action="store_true", help="N = 1: skip the BASELINE configs[2] leg (1,000 cells x 100k SNVs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="SNV-row chunks of the host-buffer (e2e) leg (0: sized from the matrix, ~96 MB per chunk)")
    ap.add_argument("--algo", default="auto", choices=["auto", "gemm", "interval"],
                    help="auto: tree-shaped clade masks take the interval path (prefix counts, no GEMM); gemm: force the tcgen05 GEMM")
    ap.add_argument("--no-gemm-leg", action="store_true", help="skip the extra GEMM-path measurement of the same workload")
    ap.add_argument("--no-percell-leg", action="store_true", help="N = 1: skip the per-cell-FN placement of the same dataset (scs_place_percell)")
    ap.add_argument("--replicates", type=int, default=1250,
                    help="replicate-sweep leg, replicates per GPU and step (BASELINE configs[4]: 10,000 pairs of 100 cells x 10k SNVs "
                         "over 8 GPUs = 1250 per GPU); 0 disables it")
    ap.add_argument("--rep-sub", type=int, default=4, help="sub-batches of the replicate leg's end-to-end pipeline")
    ap.add_argument("--rep-cells", type=int, default=100)
    ap.add_argument("--rep-snvs", type=int, default=10000)
    return ap.parse_args()


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def make_tree(n_cells, seed=1234):
    from scsommerclock_b200 import synth
    rng = np.random.default_rng(seed)
    return synth.coalescent_tree(n_cells, rng)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        while not self.stop_flag.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        p.terminate()

    def summary(self):
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        power = []
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                power.append(float(r[2]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]        # upper half: samples taken under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": float(np.median(sorted(power)[len(power) // 2:])) if power else None}


# ----------------------------------------------------------------- CPU arms
def cpu_sample_inputs(tree, n_cells, n_sample, seed=99):
    from scsommerclock_b200 import synth
    rng = np.random.default_rng(seed)
    br = synth.throw_mutations(tree, n_sample, rng)
    codes = synth.observe(tree, br, FN_TRUE, FP_TRUE, MS_TRUE, rng)
    return codes


def cpu_port_baseline(tree, S, n_cells, seconds):
    """The plain-C/OpenMP oracle (oracle/place_oracle.c) on a bounded sample."""
    from oracle import c_oracle
    cores = os.cpu_count() or 1
    n_rows = S.shape[0]
    # grow the sample until one call takes at least ~2/3 of the budget (the first call also pays thread start-up and first touch)
    n_sample = max(cores, 8)
    while True:
        codes = cpu_sample_inputs(tree, n_cells, n_sample)
        t0 = time.perf_counter()
        c_oracle.place(codes, S, FP_TRUE, FN_TRUE, cores)
        dt = time.perf_counter() - t0
        if dt >= 0.66 * seconds or n_sample >= 200000:
            break
        n_sample = int(min(200000, max(2 * n_sample, n_sample * 0.9 * seconds / max(dt, 1e-3)))) // cores * cores
    return {"value": n_sample * n_rows / dt, "unit": "placements/s", "cores": cores, "kind": "port",
            "sample": "%d SNVs x %d clade rows x %d cells of the same workload, oracle/place_oracle.c (OpenMP), %.1f s" % (n_sample, n_rows, n_cells, dt)}


def reference_arm(args):
    """The reference's own CPU implementation of the path on this box's host cores, rank 0 only.

    kind "reference": the UNMODIFIED reference module (oracle/_ref/run_PT_test.py, a byte-for-byte copy made by
    oracle/ref_module.make() in the build container; ete3 served by the tree.py stand-in) -- its read_tree (:210-334) parses
    the bench's Newick file, its map_mutations_gt (:386-448) is called on a pandas frame of a bounded sample of SNVs of
    the bench's 10,000-cell workload.  map_mutations_gt rebuilds its 2n x n clade matrices on every call (the reference
    calls it once per w_max); that fixed cost stays outside the timed steps and is reported separately, so `value` is the
    marginal per-SNV rate -- the figure that extrapolates to 1M SNVs, and the one that flatters the reference.
    kind "port" (only when oracle/_ref is missing): the same per-SNV loop as restated in oracle/pt_oracle.py."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_module
    tree_s = make_tree(args.cells)
    n_rows = 2 * args.cells            # 2n - 1 clade rows of the n + 1 sample columns ... + the reference's extra "FP" row
    per_snv = 2.4e-8 * n_rows * args.cells * 1.0
    n_sample = int(max(1, min(64, round(120.0 / (args.steps + args.warmup) / max(per_snv, 1e-3)))))
    codes = cpu_sample_inputs(tree_s, args.cells, n_sample * (args.steps + args.warmup))
    times = []      # (about two minutes of per-SNV work whatever K and W, plus the reference's one-off set-up)
    if ref_module.available():
        import pandas as pd
        ref = ref_module.load()
        names = ["tumcell%05d" % (i + 1) for i in range(args.cells)]
        tmp = tempfile.mkdtemp(prefix="scs_ref_")
        tree_file = os.path.join(tmp, "bench.newick")
        with open(tree_file, "w") as f:
            f.write(tree_s.newick(names, "healthycell", scale=0.05))
        t0 = time.perf_counter()
        tree, outg, _, _ = ref.read_tree(tree_file, np.array(names + ["healthycell"]))
        t_tree = time.perf_counter() - t0
        # one extra SNV closes the last step's interval
        codes = np.concatenate([codes, cpu_sample_inputs(tree_s, args.cells, 1, seed=98)], axis=0)
        muts = pd.DataFrame(np.where(codes == 3, np.nan, codes.astype(float)), columns=names)      # (the outgroup column is dropped before the call, :350)
        # map_mutations_gt rebuilds its 2n x n clade matrices on every call (65 s at 10,000 cells), so the K + W steps are ONE
        # call over (K + W) * n_sample SNVs; step boundaries are time stamps taken where its per-SNV loop enters
        # _normalize_log_probs (:431) -- a stamp on a module attribute, the loop body itself is untouched.
        stamps = []
        inner = ref._normalize_log_probs

        def stamped(*a, **k):
            stamps.append(time.perf_counter())
            return inner(*a, **k)

        ref._normalize_log_probs = stamped
        t0 = time.perf_counter()
        M = ref.map_mutations_gt(tree, muts, FP_TRUE, FN_TRUE / 2)
        t_setup = stamps[0] - t0
        ref._normalize_log_probs = inner
        n_rows = M.shape[1]
        for it in range(args.warmup, args.warmup + args.steps):
            times.append(stamps[(it + 1) * n_sample] - stamps[it * n_sample])
        kind = "reference"
        sample = ("%d SNV(s) per step: ONE call of the unmodified reference's map_mutations_gt (run_PT_test.py:386-448, oracle/_ref) over %d SNVs "
                  "on its own read_tree (%.1f s, untimed), steps cut at its per-SNV loop's entry into _normalize_log_probs; the call's build of "
                  "the %d x %d clade matrices (%.1f s, the reference repeats it per w_max) is outside the steps; single process (the loop is "
                  "single-threaded NumPy and holds ~%.0f GB)"
                  % (n_sample, len(muts), t_tree, n_rows, args.cells, t_setup, 6 * 8 * n_rows * args.cells / 1e9))
    else:
        from oracle import pt_oracle as O
        S = tree_s.clade_matrix().astype(np.float64)
        n_rows = S.shape[0]
        muts = np.where(codes == 3, np.nan, codes.astype(float))
        for it in range(args.warmup + args.steps):
            chunk = muts[it * n_sample:(it + 1) * n_sample]
            t0 = time.perf_counter()
            O.map_mutations_gt_literal(chunk, S, FP_TRUE, FN_TRUE / 2)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        kind = "port"
        sample = ("%d SNVs per step: map_mutations_gt's per-SNV NumPy loop (run_PT_test.py:421-430) as restated in oracle/pt_oracle.py "
                  "(oracle/_ref missing); single process" % n_sample)
    ms = 1e3 * float(np.mean(times))
    value = n_sample * n_rows / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "placements/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "coalescent %d cells x %d SNVs (BASELINE configs[3]); reference arm runs a bounded sample" % (args.cells, args.snvs),
                       "cells": args.cells, "snvs": args.snvs, "clade_rows": n_rows, "w_max": W_MAXS},
            "cpu_baseline": {"value": value, "unit": "placements/s", "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "placements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------- replicate-sweep leg
def sweep_cpu_baseline(tree_s, names, codes, tmp, seconds):
    """PT tests/s of the reference's per-(pair, w_max) loop body (run_PT_test.py:713-718: get_gt_tree -- placement and
    branch weights --, get_model_data, get_LRT_poisson with SciPy trust-constr) on ONE replicate of the sweep's shape and as
    many w_max values as fit the time budget; single process like the reference.  The unmodified reference module
    (oracle/_ref, kind "reference") when it is there, else its restatement in oracle/pt_oracle.py (kind "port")."""
    from oracle import ref_module
    tf = os.path.join(tmp, "cpu_rep.raxml.bestTree")
    with open(tf, "w") as f:
        f.write(tree_s.newick(names[:-1], "healthycell", scale=0.05))
    with open(os.path.join(tmp, "cpu_rep.raxml.log"), "w") as f:
        f.write("SEQ_ERROR: %.6f,  ADO_RATE: %.6f\n" % (FP_TRUE, FN_TRUE))
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    muts = np.concatenate([muts, np.zeros((muts.shape[0], 1))], axis=1)
    done = 0
    if ref_module.available():
        import pandas as pd
        ref = ref_module.load()
        call_data = (pd.DataFrame(muts, columns=list(names)),)
        t0 = time.perf_counter()
        for w_max in W_MAXS:
            tree, FP, FN, M = ref.get_gt_tree(tf, call_data, w_max)
            Y, constr, init, wn, cols = ref.get_model_data(tree)
            ref.get_LRT_poisson(Y, constr, init, wn[:, 0])
            done += 1
            if time.perf_counter() - t0 > seconds:
                break
        dt = time.perf_counter() - t0
        np.seterr(all="warn")          # the reference module sets a process-wide np.seterr(all='raise') on import
        kind, what = "reference", "the unmodified reference's get_gt_tree / get_model_data / get_LRT_poisson (oracle/_ref)"
    else:
        from oracle import pt_oracle as O
        mut = {"muts": muts, "samples": np.array(names)}
        t0 = time.perf_counter()
        for w_max in W_MAXS:
            r = O.get_gt_tree(tf, mut, w_max)
            Y, constr, init, wn, cols = O.get_model_data(r["tree"])
            with np.errstate(all="ignore"):
                O.get_LRT_poisson(Y, constr, init, wn[:, 0])
            done += 1
            if time.perf_counter() - t0 > seconds:
                break
        dt = time.perf_counter() - t0
        kind, what = "port", "oracle/pt_oracle.py"
    return {"value": done / dt, "unit": "PT tests/s", "cores": 1, "kind": kind,
            "sample": "%d (pair, w_max) tests of one replicate (%d cells x %d SNVs): placement + branch weights + model data + SciPy "
                      "trust-constr LRT per w_max like run_PT_test.py:713-718, %s, %.1f s" % (done, len(names) - 1, codes.shape[0], what, dt)}


def _max_rel_dev_lr(rows_a, rows_b):
    """Largest relative difference of the -2logLR fields of two lists of TSV model lines (sub-batches whose GEMM schedule
    differs from the whole batch's sum the same soft weights in another order)."""
    worst = 0.0
    for a, b in zip(rows_a, rows_b):
        fa, fb = a.split("\t"), b.split("\t")
        for i in range(3, min(len(fa), len(fb)), 4):
            try:
                x, y = float(fa[i]), float(fb[i])
            except ValueError:
                continue
            if np.isfinite(x) and np.isfinite(y):
                worst = max(worst, abs(x - y) / max(1.0, abs(y)))
    return worst


def replicate_sweep(ctx, dev, rank, world, R, n_cells, n_snv, steps, warmup, n_sub=4, cpu_seconds=0.0):
    """BASELINE configs[4]: many independent (VCF, tree) pairs, sharded over ranks with no communication.  Every replicate
    has its OWN random tree and its own genotypes.  Two timings of engine.ReplicateSweep (CUDA events, max over ranks):
      * device-resident: packed matrices in HBM, trees parsed beforehand; timed = upload of the clade masks / topology,
        per-group row filters, operand expansion, both placement passes, branch weights + LRT branch order, all R x 10
        clock tests in one launch, result read-back;
      * e2e: packed matrices in pinned HOST memory + Newick TEXT -> TSV rows through engine.ReplicateSweepStream: n_sub
        sub-batches, every H2D copy queued up front on a copy stream, trees parsed natively on a worker thread, kernels of
        sub-batch k as soon as its copy and its trees are there, TSV lines formatted natively (scs_format_pt_rows)."""
    import torch
    from scsommerclock_b200 import engine, sharding, synth
    names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
    columns = names + ["healthycell"]
    tmp = tempfile.mkdtemp(prefix="scs_rep_")
    FP, FN = FP_TRUE, FN_TRUE / 2
    trees = [make_tree(n_cells, seed=777 + 100000 * rank + r) for r in range(R)]
    newicks = [ts.newick(names, "healthycell", scale=0.05) for ts in trees]
    packs = []
    pitch = None
    for r, ts in enumerate(trees):
        lo = torch.as_tensor(ts.lo[:-1], device=dev)
        hi = torch.as_tensor(ts.hi[:-1], device=dev)
        rate = torch.as_tensor(ts.branch_len[:-1], device=dev, dtype=torch.float32)
        pk, pitch, _ = synth.observe_packed_torch(lo, hi, torch.as_tensor(ts.leaf_rank), n_snv, FN_TRUE, FP_TRUE, MS_TRUE, dev,
                                                  seed=5000 + 100000 * rank + r, rate=rate, extra_cols=1, chunk=n_snv)
        packs.append(pk)
    packed = torch.cat(packs, dim=0).contiguous()
    del packs
    sweep = engine.ReplicateSweep(ctx, [n_snv] * R, columns, W_MAXS)
    assert sweep.pitch == pitch
    t0 = time.perf_counter()
    sweep.parse_trees(newicks)
    parse_ms = 1e3 * (time.perf_counter() - t0)

    def step():
        sweep.enqueue(packed, FP, FN)
        return sweep.results()

    def timed(fn, n_steps, n_warm):
        for _ in range(n_warm):
            res = fn()
        torch.cuda.synchronize()
        sharding.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            res = fn()
        e1.record()
        torch.cuda.synchronize()
        return sharding.max_over_ranks(e0.elapsed_time(e1) / n_steps, dev), res

    ms, (out, n_kept, status) = timed(step, steps, max(1, warmup))
    gemm = sweep.place.last_times()
    ok = int((out[:, :, 7] == 0).sum())
    ref_rows = sweep.rows()
    result = {"replicates_per_gpu": R, "cells": n_cells, "snvs": n_snv, "w_max_per_replicate": len(W_MAXS), "distinct_trees": R,
              "ms_per_step": ms, "pt_tests_per_s": R * len(W_MAXS) * world / (ms / 1e3),
              "placements_per_s": float(R) * n_snv * sweep.n_rows * world / (ms / 1e3),
              "pass1_ms": gemm[0], "merge_ms": gemm[1], "pass2_ms": gemm[2], "tree_parse_ms_untimed": parse_ms,
              "lrt_converged": ok, "lrt_problems": int(out.shape[0] * out.shape[1]), "groups_flagged": int((status != 0).sum()),
              "median_newton_iterations": float(np.median(out[:, :, 6])), "tiling": sweep.place.info()}
    # ---- end to end: host matrices + Newick text -> TSV rows
    pinned = torch.empty(packed.shape, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(packed)
    torch.cuda.synchronize()
    n_sub = max(1, min(n_sub, R))
    pipe = engine.ReplicateSweepStream(ctx, [n_snv] * R, columns, W_MAXS, n_sub=n_sub)
    host_ms = []

    def step_e2e_timed():
        t0 = time.perf_counter()
        rows, _ = pipe.run(pinned, newicks, FP, FN)
        host_ms.append(1e3 * (time.perf_counter() - t0))
        return rows

    ms_e2e, rows = timed(step_e2e_timed, steps, 1)
    result["e2e"] = {"value": R * len(W_MAXS) * world / (ms_e2e / 1e3), "unit": "PT tests/s", "ms_per_step": ms_e2e,
                     "h2d_bytes_per_step": int(pinned.numel() + R * sweep.n_rows * (sweep.words + 2) * 4 + R * 16) * world,
                     "d2h_bytes_per_step": int(R * len(W_MAXS) * 64 + R * 12) * world, "sub_batches": pipe.n_sub,
                     "host_wall_ms_per_step": float(np.mean(host_ms[-steps:])),
                     "input": "packed genotype matrices in pinned host memory + Newick text of %d distinct trees" % R,
                     "output": "TSV model lines (scs_format_pt_rows)", "rows_equal_device_leg": bool(rows == ref_rows),
                     "max_rel_dev_LR_vs_device_leg": _max_rel_dev_lr(rows, ref_rows),
                     "vs_device_resident": ms_e2e / ms}
    if rank == 0 and cpu_seconds > 0:
        codes = synth.observe(trees[0], synth.throw_mutations(trees[0], n_snv, np.random.default_rng(3)), FN_TRUE, FP_TRUE, MS_TRUE,
                              np.random.default_rng(4))
        try:
            result["cpu_baseline"] = sweep_cpu_baseline(trees[0], columns, codes, tmp, cpu_seconds)
        except Exception as exc:          # (a reported extra: never lose the line over it)
            result["cpu_baseline"] = {"error": repr(exc)}
    pipe.close()
    sweep.close()
    return result


# ------------------------------------------------------------------ GPU arm
class SingleDataset:
    """One tree + this rank's shard of SNV rows, ready to be timed: the plan, the LRT batch for the ten w_max, the resident
    step and the host-buffer (e2e) step.  Every rank must construct it at the same time (one all-reduce inside)."""

    def __init__(self, ctx, dev, rank, world, n_cells, n_local, algo, e2e=True, e2e_chunks=0, tree_seed=1234, seed=1000):
        import torch
        from scsommerclock_b200 import engine, run_PT_test as P, sharding, synth
        self.ctx, self.dev, self.world, self.n_cells, self.n_local = ctx, dev, world, n_cells, n_local
        # ---- the tree (same on every rank), through the product's own tree code
        self.tree_s = tree_s = make_tree(n_cells, seed=tree_seed)
        names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
        self.columns = columns = names + ["healthycell"]
        tmp = tempfile.mkdtemp(prefix="scs_bench_")
        tree_file = os.path.join(tmp, "bench_r%d.newick" % rank)
        with open(tree_file, "w") as f:
            f.write(tree_s.newick(names, "healthycell", scale=0.05))
        tree, outg, _, _ = P.read_tree(tree_file, np.array(columns))
        self.active = active = np.array([c != outg for c in columns], dtype=np.uint8)
        self.bits, self.words, nodes = P._clade_bits(tree, columns, active)
        self.n_rows = n_rows = self.bits.shape[0]
        self.n_cols = n_cols = len(columns)
        # ---- this rank's shard of SNV rows, generated on the device in the packed layout
        lo = torch.as_tensor(tree_s.lo[:-1], device=dev)
        hi = torch.as_tensor(tree_s.hi[:-1], device=dev)
        rate = torch.as_tensor(tree_s.branch_len[:-1], device=dev, dtype=torch.float32)
        self.packed, self.pitch, _ = synth.observe_packed_torch(lo, hi, torch.as_tensor(tree_s.leaf_rank), n_local, FN_TRUE, FP_TRUE, MS_TRUE, dev,
                                                                seed=seed + rank, rate=rate, extra_cols=1)
        torch.cuda.synchronize()
        self.FP, self.FN = FP_TRUE, FN_TRUE / 2        # what get_gt_tree would pass for a CellPhy log with these rates
        os.environ["SCS_PLACE_ALGO"] = algo
        self.plan = engine.Placement(ctx, n_local, n_cols, n_rows)
        self.plan.set_clades_host(self.bits, self.words)
        self.y_dev = torch.zeros(n_rows, dtype=torch.float64, device=dev)
        self.stream = torch.cuda.current_stream()
        # ---- one untimed pass fixes the LRT model (branch order, weights) for the timed steps
        self.place_step()
        n_kept, miss = self.plan.prepare_result(stream=self.stream)
        torch.cuda.synchronize()
        soft = self.y_dev.cpu().numpy()
        if world > 1:
            t = torch.tensor(np.concatenate([[float(n_kept)], miss * n_kept]), dtype=torch.float64, device=dev)
            sharding.allreduce_soft_weights(t)
            n_kept = float(t[0].item())
            miss = (t[1:] / n_kept).cpu().numpy()
        FP, FN = self.FP, self.FN
        MS_i = {c: float(np.clip(miss[i], P.LAMBDA_MIN, 1 - FN - 0.2)) for i, c in enumerate(columns) if active[i]}
        st = {"soft": soft, "MS_i": MS_i, "n_muts": n_kept, "bits": self.bits, "words": self.words, "columns": columns, "weights": {}}
        P._apply_placement(tree, st)
        name_row = {n.name: i for i, n in enumerate(nodes)}
        childs, weights, gathers = [], [], []
        for w_max in W_MAXS:
            P.add_br_weights(tree, FP, FN, MS_i, w_max, _state=st)
            Y, constr, init, wn, cols = P.get_model_data(tree)
            childs.append(constr.child_int)
            weights.append(wn[:, 0].copy())
            gathers.append(np.array([name_row[c] for c in cols], dtype=np.int32))
        self.lrt = engine.LrtPlan(ctx, childs)
        self.lrt.set_weights(weights)
        self.lrt.set_gather(np.concatenate(gathers), n_rows)
        self.lrt_out = torch.zeros((len(W_MAXS), 8), dtype=torch.float64, device=dev)
        self.pinned = self.streamer = None
        if e2e:
            self.pinned = torch.empty(self.packed.shape, dtype=torch.uint8, pin_memory=True)
            self.pinned.copy_(self.packed)
            torch.cuda.synchronize()
            # host-buffer path: the packed matrix is streamed from pinned host memory in SNV-row chunks
            # (H2D of chunk c+1 under the kernels of chunk c), see engine.StreamingPlacement
            self.streamer = engine.StreamingPlacement(ctx, n_local, n_cols, n_rows, n_chunks=e2e_chunks or None, pitch=self.pitch)
            self.streamer.set_clades_host(self.bits, self.words)

    def place_step(self):
        from scsommerclock_b200 import sharding
        # row filter + per-cell missing counts + (interval path) the cell-order pass: one read of the packed matrix, counters
        # stay on the device; GEMM path: the two reduction kernels + the operand plane(s)
        self.plan.prepare(self.packed, self.pitch, self.active, True, stream=self.stream)
        self.plan.run(self.FP, self.FN, self.y_dev, stream=self.stream)
        sharding.allreduce_soft_weights(self.y_dev)

    def step(self):
        self.place_step()
        self.lrt.run(self.y_dev, self.lrt_out, stream=self.stream)
        return self.lrt_out.cpu()             # the step's result, read back to the host

    def step_e2e(self):
        from scsommerclock_b200 import sharding
        self.streamer.run(self.pinned, self.pitch, self.active, True, self.FP, self.FN, self.y_dev)
        sharding.allreduce_soft_weights(self.y_dev)
        self.lrt.run(self.y_dev, self.lrt_out, stream=self.stream)
        return self.lrt_out.cpu()

    def d2h_bytes(self):
        return int(self.lrt_out.numel() * 8 + self.n_cols * 4 + 8)

    def close(self):
        import torch
        if self.streamer is not None:
            self.streamer.close()
        self.plan.close()
        self.lrt.close()
        self.packed = self.pinned = self.streamer = None
        torch.cuda.empty_cache()


def timed(fn, steps, warmup, plan, dev, sampler=None, lrt_launches_per_step=1, prep=True):
    """W untimed + K timed steps between barriers and synchronisations; CUDA events; max over ranks.
    Returns (total ms, per-step plan kernel times, last result, this library's launches inside the timed region)."""
    import torch
    from scsommerclock_b200 import sharding
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    sharding.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    l0 = plan.info()["launches"]
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    times = []
    e0.record()
    for _ in range(steps):
        res = fn()
        times.append(plan.last_times() + [plan.prepare_time() if prep else 0.0])
    e1.record()
    torch.cuda.synchronize()
    sharding.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.stop_flag.set()
    ms = e0.elapsed_time(e1)
    # this library's kernels inside the timed region: plan launches (the small upload kernels of the active-cell mask and
    # the error-model constants, pre-pass, placement, reduce / operand expansion, two GEMM passes, merge, reduce) + the LRT
    # kernel of every step
    n_launch = plan.info()["launches"] - l0 + steps * lrt_launches_per_step
    return sharding.max_over_ranks(ms, dev), times, res, n_launch


def ncu_constants():
    """Per-launch figures that only a profiler can give (DRAM bytes, executed warp instructions), read from the summary
    the capture script wrote under profiles/ -- never pasted into this file.  {} when no capture of the current kernel exists."""
    try:
        return json.load(open(os.path.join(REPO, "profiles", "r02_ncu_constants.json")))
    except Exception:
        return {}


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    # stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL's
    # version banner does, whatever NCCL_DEBUG_FILE says), so everything that is not the result line is sent to
    # stderr at the descriptor level and the line itself goes to a private copy of the original stdout.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    from scsommerclock_b200 import engine, sharding
    rank, world, local = sharding.init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = engine.Context(local)
    n_cells, n_total = args.cells, args.snvs
    pk, pk_kind = peaks()
    ncu = ncu_constants()

    # ---- BASELINE configs[3]: ONE dataset of n_total SNVs, its rows block-partitioned over the ranks (strong scaling)
    r_lo, r_hi = sharding.shard_range(n_total, rank, world)
    n_snv = r_hi - r_lo
    ds = SingleDataset(ctx, dev, rank, world, n_cells, n_snv, args.algo, e2e=not args.no_e2e, e2e_chunks=args.e2e_chunks)
    n_rows, n_cols, pitch = ds.n_rows, ds.n_cols, ds.pitch
    plan = ds.plan

    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, gemm_times, res, launches = timed(ds.step, args.steps, args.warmup, plan, dev, sampler)
    plan_info = plan.info()
    ms_step = ms_total / args.steps
    placements = float(n_total) * n_rows
    value = placements / (ms_step / 1e3)

    e2e = None
    if not args.no_e2e:
        y_ref = ds.y_dev.clone()
        ms_e2e, _, _, _ = timed(ds.step_e2e, args.steps, min(args.warmup, 1), plan, dev)
        e2e_dev = float((ds.y_dev - y_ref).abs().max() / y_ref.abs().max())     # streamed == resident, to summation order
        e2e = {"value": placements / (ms_e2e / args.steps / 1e3), "unit": "placements/s",
               "h2d_bytes_per_step": int(n_total) * pitch, "d2h_bytes_per_step": ds.d2h_bytes() * world,
               "ms_per_step": ms_e2e / args.steps, "chunks": ds.streamer.n_chunks, "max_rel_dev_vs_resident": e2e_dev}

    # ---- the same workload through the tcgen05 GEMM (north_star's formulation), when the headline ran the interval path
    gemm_leg = None
    if plan_info["algo"] == 2 and not args.no_gemm_leg and world == 1:
        os.environ["SCS_PLACE_ALGO"] = "gemm"
        plan_g = engine.Placement(ctx, n_snv, n_cols, n_rows)
        plan_g.set_clades_host(ds.bits, ds.words)
        os.environ["SCS_PLACE_ALGO"] = args.algo
        y_g = torch.zeros_like(ds.y_dev)
        y_iv = ds.y_dev.clone()

        def step_gemm():
            plan_g.prepare(ds.packed, pitch, ds.active, True, stream=ds.stream)
            plan_g.run(ds.FP, ds.FN, y_g, stream=ds.stream)
            sharding.allreduce_soft_weights(y_g)
            ds.lrt.run(y_g, ds.lrt_out, stream=ds.stream)
            return ds.lrt_out.cpu()

        ms_g, g_times, _, _ = timed(step_gemm, 3, 3, plan_g, dev)
        gemm_leg = {"ms_per_step": ms_g / 3, "value": placements / (ms_g / 3 / 1e3), "times": np.array(g_times), "info": plan_g.info(),
                    "max_rel_dev_vs_interval": float((y_g - y_iv).abs().max() / y_iv.abs().max())}
        plan_g.close()
        del plan_g, y_g, y_iv
    # ---- the same dataset placed with ONE FALSE-NEGATIVE RATE PER CELL (the pd.Series branch of map_mutations_gt, :413-418):
    # scs_place_percell, float64 prefix sums of per-cell weights.  A side leg -- the reference's command line never takes it.
    percell = None
    if world == 1 and not args.no_percell_leg:
        try:
            rates = np.clip(np.random.default_rng(5).normal(ds.FN, 0.03, n_cols), 0.01, 0.4)
            keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
            engine.gt_reduce(ctx, ds.packed, pitch, n_snv, n_cols, ds.active, True, keep, stream=ds.stream)
            y_pc = torch.zeros_like(ds.y_dev)
            kms = [engine.place_percell(ctx, ds.packed, pitch, n_snv, n_cols, keep, ds.bits, ds.words, ds.FP, rates, y_pc, stream=ds.stream)
                   for _ in range(4)][1:]
            engine.place_percell(ctx, ds.packed, pitch, n_snv, n_cols, keep, ds.bits, ds.words, ds.FP, np.full(n_cols, ds.FN), y_pc, stream=ds.stream)
            percell = {"workload": "the headline dataset (%d cells x %d SNVs), per-cell FN rates ~ N(%.3g, 0.03)" % (n_cells, n_snv, ds.FN),
                       "kernel_ms": float(np.mean(kms)), "launches_timed": len(kms), "value": placements / (float(np.mean(kms)) / 1e3),
                       "unit": "placements/s", "kernel": "scs_place_percell2_kernel<R> (1,024..12,287 cells; else scs_place_percell_kernel) + scs_percell_reduce_kernel (device time; the call also "
                       "analyses the clade masks on the host and uploads three tables)",
                       "max_rel_dev_equal_rates_vs_count_kernels": float((y_pc - ds.y_dev).abs().max() / ds.y_dev.abs().max())}
            del keep, y_pc
        except Exception as exc:                        # (a side leg: never lose the line over it)
            percell = {"error": repr(exc)}
    lrt_newton = res.numpy().copy()
    S_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        S_cpu = np.ascontiguousarray(engine.unpack_clades(ds.bits, n_cols)[:, :n_cells].astype(np.uint8))
    tree_main = ds.tree_s
    ds.close()
    del ds

    # ---- weak-scaling variant beside it (N > 1): n_total SNVs on EVERY rank, same exchange
    weak = None
    if world > 1 and not args.no_weak_leg:
        dw = SingleDataset(ctx, dev, rank, world, n_cells, n_total, args.algo, e2e=False)
        ms_w, tw, _, _ = timed(dw.step, 3, 3, dw.plan, dev)
        weak = {"snvs_per_gpu": n_total, "ms_per_step": ms_w / 3, "steps": 3, "warmup": 3,
                "value": float(n_total) * world * n_rows / (ms_w / 3 / 1e3), "unit": "placements/s",
                "placement_kernel_ms": float(np.mean([t[2] for t in tw])),
                "note": "every rank holds its own %d SNVs (the round-1 headline); value = all ranks' placements / max-over-ranks time" % n_total}
        dw.close()
        del dw

    # ---- BASELINE configs[2]: 1,000 cells x 100k SNVs, a single PT test on one B200
    small = None
    if world == 1 and not args.no_small_leg:
        d2 = SingleDataset(ctx, dev, rank, world, 1000, 100000, "auto", e2e=not args.no_e2e, e2e_chunks=1, tree_seed=4321, seed=2000)
        ms_2, t2, r2, l2 = timed(d2.step, 20, 5, d2.plan, dev)
        info2 = d2.plan.info()
        small = {"workload": "coalescent 1000 cells x 100000 SNVs, single PT test (10 w_max) on one GPU (BASELINE configs[2])",
                 "ms_per_step": ms_2 / 20, "steps": 20, "warmup": 5, "value": 100000.0 * d2.n_rows / (ms_2 / 20 / 1e3), "unit": "placements/s",
                 "algorithm": "interval" if info2["algo"] == 2 else "tcgen05 GEMM", "gpu_launches": int(l2),
                 "kernel_ms": {"pass1_or_prepass": float(np.mean([t[0] for t in t2])), "merge": float(np.mean([t[1] for t in t2])),
                               "pass2_or_placement": float(np.mean([t[2] for t in t2])), "reduce": float(np.mean([t[3] for t in t2]))},
                 "lrt_status": [int(v) for v in r2.numpy()[:, 7]], "newton_iterations": [int(v) for v in r2.numpy()[:, 6]]}
        if not args.no_e2e:
            ms_2e, _, _, _ = timed(d2.step_e2e, 20, 3, d2.plan, dev)
            small["e2e"] = {"ms_per_step": ms_2e / 20, "value": 100000.0 * d2.n_rows / (ms_2e / 20 / 1e3), "unit": "placements/s",
                            "h2d_bytes_per_step": 100000 * d2.pitch, "d2h_bytes_per_step": d2.d2h_bytes(), "chunks": d2.streamer.n_chunks}
        d2.close()
        del d2

    rep = None
    if args.replicates > 0:
        torch.cuda.empty_cache()
        rep = replicate_sweep(ctx, dev, rank, world, args.replicates, args.rep_cells, args.rep_snvs, max(2, args.steps), 1, n_sub=args.rep_sub,
                              cpu_seconds=(args.cpu_seconds if (world == 1 and not args.no_cpu_baseline) else 0.0))

    if rank == 0:
        g = np.array(gemm_times)                         # [steps][pass1, merge, pass2, reduce]
        fp8_yard = None                                  # sustained cuBLASLt FP8 GEMM measured on this pool (tools/measure_fp8_peak.py)
        try:
            fp8_yard = json.load(open(os.path.join(REPO, "profiles", "r01_fp8_yardstick.json")))["fp8_scaled_mm_8192"]["sustained_tflops"]
        except Exception:
            pass
        peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
        full_size = n_cells == 10000 and n_snv == 1000000

        def gemm_roofline(g, info):
            cached = info.get("pass2_mode") == 2                             # pass 2 = sweep over the cached integer counts, not a GEMM
            t_gemm = float(np.mean(g[:, 0] if cached else np.concatenate([g[:, 0], g[:, 2]])))   # average launch of the dominant kernel
            flop = 2.0 * n_snv * n_rows * n_cols                             # algorithmic: SNV x cells x clade rows, 2 flop per MAC
            achieved = flop / (t_gemm / 1e3) / 1e12
            traffic = None
            if full_size and info["mode"] == 1:
                traffic = (ncu.get("gemm_cache_pass1") or {}).get("dram_bytes") if cached else (ncu.get("gemm_pass") or {}).get("dram_bytes")
            return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "peak_source": pk_kind + " cuBLAS bf16, sustained (kernel timed inside a long step)",
                    "kernel": "scs_place_kernel (%s), %s" % (
                        "tcgen05 kind::f8f6f4, one radix-packed e5m2 plane" if info["mode"] == 1 else "tcgen05 kind::i8, two u8 planes",
                        "pass 1 launches (pass 2 sweeps the cached integer counts: HBM bound, see pass2_*)" if cached else "avg of pass 1 and pass 2 launches"),
                    "peak_8bit_equiv": 2 * peak, "frac_8bit_equiv": achieved / (2 * peak),
                    "fp8_cublaslt_sustained_tflops": fp8_yard, "frac_vs_fp8_cublaslt_sustained": (achieved / fp8_yard) if fp8_yard else None,
                    "why_frac_exceeds_1": "MEASURED_PEAKS.json holds a bf16 figure only; this GEMM runs 8-bit operands (FP8 e5m2 / u8), whose tensor rate is twice bf16: against 2 x the measured bf16 rate the kernel sits at frac_8bit_equiv",
                    "algorithmic_flop_per_launch": flop, "ms_per_launch": t_gemm,
                    "pass1_ms": float(g[:, 0].mean()), "merge_ms": float(g[:, 1].mean()), "pass2_ms": float(g[:, 2].mean()),
                    "reduce_ms": float(g[:, 3].mean()),
                    "pass2_kind": "count-cache sweep (scs_colsum_cached_kernel)" if cached else "second GEMM pass",
                    "pass2_hbm_GBps": (4.0 * info["M_pad"] * n_rows / 1e9 / (float(g[:, 2].mean()) / 1e3)) if cached else None,
                    "count_cache_GB": (4.0 * info["M_pad"] * info["N_pad"] / 1e9) if cached else 0.0,
                    "executed_tensor_TOPs": achieved * (1 if info["mode"] == 1 else 2),
                    "note": "radix mode: one exact FP8 GEMM per contraction (counts packed as C1 + 4096*C0); planes mode: two exact u8 GEMMs"}

        if plan_info["algo"] == 2:
            # interval path: the packed 2-bit genotypes are read once by the pre-pass (row filter, missing counts, cell order) and the
            # reordered copy once by the placement kernel; nothing else touches HBM but the per-run partial sums.  The kernel is bound
            # by instruction issue, not by HBM or the tensor pipe -- of the two rooflines the contract names, HBM is the one its only
            # off-chip traffic belongs to.
            t_k = float(g[:, 2].mean())                                      # the placement kernel (g[:, 0]: the pre-pass before it)
            alg_bytes = float(n_snv) * pitch
            hbm_peak = pk.get("hbm_gbs_sustained", pk.get("hbm_gbs"))
            iv = ncu.get("interval_kernel") or {}
            scale = float(n_snv) / float(iv["snvs"]) if iv.get("snvs") else None
            roofline = {"bound": "hbm", "achieved": alg_bytes / 1e9 / (t_k / 1e3), "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / 1e9 / (t_k / 1e3) / hbm_peak,
                        "traffic": (iv["dram_bytes"] * scale) if (scale and n_cells == 10000 and iv.get("dram_bytes")) else None,
                        "traffic_source": iv.get("source") if scale else None,
                        "traffic_note": ("the timed pair reads the reordered rows twice (placement kernel, then the leaf kernel): ~2 x the algorithmic "
                                         "bytes by design -- the second read is 0.4 ms of HBM time, the leaf accumulators it moves out of the "
                                         "placement kernel's shared memory were worth 3.1 ms there") if n_cells >= 2048 else None,
                        "peak_source": pk_kind + " HBM copy bandwidth",
                        "kernel": "scs_place_iv%d_kernel (prefix counts in shared memory, no GEMM)%s" % (
                            plan_info["interval_variant"], " + scs_iv4_leaf_kernel (leaf accumulators, from 2,048 cells on): one timed pair" if n_cells >= 2048 else ""),
                        "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": t_k, "prepass_ms": float(g[:, 4].mean()),
                        "prepass_GBps": 2.0 * alg_bytes / 1e9 / (float(g[:, 4].mean()) / 1e3) if g[:, 4].mean() > 0 else None,
                        "prepass_note": "scs_permute_reduce4_kernel: row filter + per-cell missing counts + rows rewritten in depth-first cell order, one read and one write of the packed matrix",
                        "reduce_ms": float(g[:, 3].mean()),
                        "entries_per_s": float(n_snv) * n_rows / (t_k / 1e3),
                        "actual_limiter": "instruction issue / shared-memory latency: one 640-thread CTA per SM (5 warps per scheduler); HBM traffic is "
                                          "~0.1 TB/s because the algorithm needs O(cells + clade rows) work per SNV instead of the GEMM's "
                                          "O(cells x clade rows) -- see issue_roofline, and gemm_path for the tensor-pipe roofline of the same "
                                          "workload through the tcgen05 GEMM",
                        "smem_bytes": plan_info["interval_smem_bytes"], "threads": plan_info["interval_threads"], "grid": plan_info["interval_grid"]}
            if scale and n_cells == 10000 and iv.get("warp_instructions"):
              try:
                # the roofline that does bound it: issue slots (4 schedulers per SM, one warp instruction per cycle each)
                sm_count = getattr(ctx, "sm_count", 148)
                clk = ((sampler.summary().get("sm_mhz") if sampler else None) or pk.get("sm_max_mhz", 1965.0)) * 1e6
                inst = iv["warp_instructions"] * scale
                roofline["issue_roofline"] = {"warp_instructions_per_launch": inst, "achieved_per_s": inst / (t_k / 1e3),
                                              "peak_per_s": sm_count * 4 * clk, "frac": inst / (t_k / 1e3) / (sm_count * 4 * clk),
                                              "note": "instruction count: smsp__inst_executed.sum of %s, scaled by SNV count" % iv.get("source")}
              except Exception as exc:          # (an explanatory extra: never lose the line over it)
                roofline["issue_roofline"] = {"error": repr(exc)}
        else:
            roofline = gemm_roofline(g, plan_info)
        gemm_path = None
        if gemm_leg is not None:
            gemm_path = {"ms_per_step": gemm_leg["ms_per_step"], "value": gemm_leg["value"], "unit": "placements/s", "steps": 3, "warmup": 3,
                         "max_rel_dev_soft_weights_vs_interval": gemm_leg["max_rel_dev_vs_interval"],
                         "roofline": gemm_roofline(gemm_leg["times"], gemm_leg["info"]), "tiling": gemm_leg["info"]}
        out = lrt_newton
        line = {
            "metric": METRIC, "value": value, "unit": "placements/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": ("u16 prefix counts (exact integers), f32 softmax terms on exact integer differences (LRT f64)" if plan_info["algo"] == 2 else
                      "fp8_e5m2 operands holding exact integer counts, f32 accumulate (epilogue f32, LRT f64)" if plan_info["mode"] == 1 else
                      "u8 operands, s32 accumulate (epilogue f32, LRT f64)"),
            "data": "synthetic",
            "config": {"workload": "coalescent %d cells x %d SNVs (BASELINE configs[3]): one dataset, SNV rows block-partitioned over the GPUs, "
                                   "one all-reduce of the branch weights, LRT for 10 w_max" % (n_cells, n_total),
                       "cells": n_cells, "snvs": n_total, "snvs_per_gpu": n_snv, "clade_rows": n_rows, "w_max": W_MAXS, "FN": FN_TRUE, "FP": FP_TRUE,
                       "missing": MS_TRUE,
                       "algorithm": "interval path (tree-shaped clade masks: counts = differences of per-SNV prefix counts)" if plan_info["algo"] == 2 else "tcgen05 GEMM",
                       "l2": ("inputs larger than L2 (packed genotypes %.2f GB per GPU, read once per step)" % (float(n_snv) * pitch / 1e9)) if plan_info["algo"] == 2 else
                             ("inputs larger than L2 (operand plane(s) %.1f GB per GPU)" % ((1 if plan_info["mode"] == 1 else 2) * float(plan_info["M_pad"]) * plan_info["K_pad"] / 1e9)),
                       "tiling": plan_info},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "gemm_path": gemm_path,
            "weak_scaling": weak,
            "single_test_1k_cells": small,
            "per_cell_fn": percell,
            "pt_tests": {"per_step": len(W_MAXS), "LR": [float(v) for v in out[:, 0]], "p_value": [float(v) for v in out[:, 3]],
                         "newton_iterations": [int(v) for v in out[:, 6]], "status": [int(v) for v in out[:, 7]]},
            "clocks": sampler.summary() if sampler else None,
            "replicate_sweep": rep,
        }
        if S_cpu is not None:
            line["cpu_baseline"] = cpu_port_baseline(tree_main, S_cpu, n_cells, args.cpu_seconds)
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    sharding.barrier()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
