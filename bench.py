#!/usr/bin/env python3
"""bench.py -- the PT-test hot path on synthetic coalescent-shaped data.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        (the reference's CPU algorithm, rank 0 only)

Workload (BASELINE.json configs[3]): 10,000 cells x 1,000,000 SNVs PER GPU, SNV
rows sharded over ranks (weak scaling), one all-reduce of the 2n branch weights,
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
# dram__bytes_read.sum + dram__bytes_write.sum of one scs_place_kernel launch at the default
# workload, from the ncu --set full capture under profiles/ (None until a capture exists)
# (profiles/r01_ncu_place_radix_summary.txt: 287.1 GB read + 2.5 GB written, pass 1, radix mode;
# earlier captures of the same kernel read 280-310 GB -- it depends on how the L2 halves fill)
NCU_TRAFFIC_BYTES = 289.6e9
# pass 1 that also fills the count cache (profiles/r01_ncu_place_cache_summary.txt, 250k-SNV capture x 4:
# 222.8 GB read + 82.7 GB written) and the interval kernel (profiles/r01_ncu_interval_summary.txt, captured at
# the full size: 2.52 GB read = the permuted genotypes once, 0.30 GB of per-run partials written)
NCU_TRAFFIC_BYTES_CACHE = 305.5e9
NCU_TRAFFIC_BYTES_INTERVAL = 2.82e9
# warp instructions the interval kernel executes per SNV at 10,001 columns x 20,000 clade rows
# (smsp__inst_executed.sum of profiles/r01_ncu_interval_final_250k_summary.txt / 250,000 SNVs)
NCU_INTERVAL_WARP_INST_PER_SNV = 5282607802.0 / 250000.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=10000)
    ap.add_argument("--snvs", type=int, default=1000000, help="SNVs per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="SNV-row chunks of the host-buffer (e2e) leg (0: sized from the matrix, ~96 MB per chunk)")
    ap.add_argument("--algo", default="auto", choices=["auto", "gemm"],
                    help="auto: tree-shaped clade masks take the interval path (prefix counts, no GEMM); gemm: force the tcgen05 GEMM")
    ap.add_argument("--no-gemm-leg", action="store_true", help="skip the extra GEMM-path measurement of the same workload")
    ap.add_argument("--replicates", type=int, default=1250,
                    help="replicate-sweep leg, replicates per GPU and step (BASELINE configs[4]: 10,000 pairs of 100 cells x 10k SNVs "
                         "over 8 GPUs = 1250 per GPU); 0 disables it")
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
    probe = max(cores, 8)
    codes = cpu_sample_inputs(tree, n_cells, probe)
    t0 = time.perf_counter()
    c_oracle.place(codes, S, FP_TRUE, FN_TRUE, cores)
    dt = time.perf_counter() - t0
    n_sample = int(min(max(probe, seconds / max(dt / probe, 1e-9)), 200000))
    n_sample = max(cores, n_sample // cores * cores)
    codes = cpu_sample_inputs(tree, n_cells, n_sample)
    t0 = time.perf_counter()
    c_oracle.place(codes, S, FP_TRUE, FN_TRUE, cores)
    dt = time.perf_counter() - t0
    return {"value": n_sample * n_rows / dt, "unit": "placements/s", "cores": cores, "kind": "port",
            "sample": "%d SNVs x %d clade rows x %d cells of the same workload, oracle/place_oracle.c (OpenMP), %.1f s" % (n_sample, n_rows, n_cells, dt)}


def reference_arm(args):
    """The reference's own algorithm for the path -- map_mutations_gt's per-SNV NumPy
    loop (run_PT_test.py:421-430), restated in oracle/pt_oracle.py because the Python
    reference cannot travel to the GPU box -- on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pt_oracle as O
    tree = make_tree(args.cells)
    S = tree.clade_matrix().astype(np.float64)
    n_rows = S.shape[0]
    # a bounded sample: a few SNVs per step (each costs ~8 passes over a 2n x n float64 matrix), sized so
    # that the whole --steps/--warmup run stays around two minutes whatever K and W the caller asks for
    per_snv = 2.4e-8 * n_rows * args.cells * 1.0
    n_sample = int(max(1, min(64, round(120.0 / (args.steps + args.warmup) / max(per_snv, 1e-3)))))
    codes = cpu_sample_inputs(tree, args.cells, n_sample * (args.steps + args.warmup))
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    times = []
    for it in range(args.warmup + args.steps):
        chunk = muts[it * n_sample:(it + 1) * n_sample]
        t0 = time.perf_counter()
        O.map_mutations_gt_literal(chunk, S, FP_TRUE, FN_TRUE)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_sample * n_rows / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "placements/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "coalescent %d cells x %d SNVs per GPU (BASELINE configs[3]); reference arm runs a bounded sample" % (args.cells, args.snvs),
                       "cells": args.cells, "snvs_per_gpu": args.snvs, "clade_rows": n_rows, "w_max": W_MAXS},
            "cpu_baseline": {"value": value, "unit": "placements/s", "cores": 1, "kind": "port",
                             "sample": "%d SNVs per step: map_mutations_gt's per-SNV NumPy loop (run_PT_test.py:421-430) as restated in "
                                       "oracle/pt_oracle.py; single process (the loop is single-threaded NumPy and needs ~%.0f GB at this size)"
                                       % (n_sample, 6 * 8 * n_rows * args.cells / 1e9)},
            "e2e": {"value": value, "unit": "placements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------- replicate-sweep leg
def sweep_cpu_baseline(tree_s, names, codes, tmp, seconds):
    """PT tests/s of the reference's per-(pair, w_max) loop body (run_PT_test.py:713-718: get_gt_tree -- placement and
    branch weights --, get_model_data, get_LRT_poisson with SciPy trust-constr) as restated in oracle/pt_oracle.py, on ONE
    replicate of the sweep's shape and as many w_max values as fit the time budget; single process like the reference."""
    from oracle import pt_oracle as O
    tf = os.path.join(tmp, "cpu_rep.raxml.bestTree")
    with open(tf, "w") as f:
        f.write(tree_s.newick(names[:-1], "healthycell", scale=0.05))
    with open(os.path.join(tmp, "cpu_rep.raxml.log"), "w") as f:
        f.write("SEQ_ERROR: %.6f,  ADO_RATE: %.6f\n" % (FP_TRUE, FN_TRUE))
    muts = np.where(codes == 3, np.nan, codes.astype(float))
    muts = np.concatenate([muts, np.zeros((muts.shape[0], 1))], axis=1)
    mut = {"muts": muts, "samples": np.array(names)}
    t0 = time.perf_counter()
    done = 0
    for w_max in W_MAXS:
        r = O.get_gt_tree(tf, mut, w_max)
        Y, constr, init, wn, cols = O.get_model_data(r["tree"])
        with np.errstate(all="ignore"):
            O.get_LRT_poisson(Y, constr, init, wn[:, 0])
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "PT tests/s", "cores": 1, "kind": "port",
            "sample": "%d (pair, w_max) tests of one replicate (%d cells x %d SNVs): placement + branch weights + model data + SciPy "
                      "trust-constr LRT per w_max like run_PT_test.py:713-718, oracle/pt_oracle.py, %.1f s" % (done, len(names) - 1, codes.shape[0], dt)}


def replicate_sweep(ctx, dev, rank, world, R, n_cells, n_snv, steps, warmup, n_sub=4, cpu_seconds=0.0):
    """BASELINE configs[4]: many independent (VCF, tree) pairs, sharded over ranks with no communication.  Every replicate
    has its OWN random tree and its own genotypes.  Two timings of engine.ReplicateSweep (CUDA events, max over ranks):
      * device-resident: packed matrices in HBM, trees parsed beforehand; timed = upload of the clade masks / topology,
        per-group row filters, operand expansion, both placement passes, branch weights + LRT branch order, all R x 10
        clock tests in one launch, result read-back;
      * e2e: packed matrices in pinned HOST memory + Newick TEXT -> TSV rows.  The batch is cut into n_sub sub-batches on
        their own streams so that the native tree parsing (host threads) and the H2D copy of sub-batch k + 1 run under the
        kernels of sub-batch k; the TSV lines are formatted natively (scs_format_pt_rows)."""
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
    cuts = [R * k // n_sub for k in range(n_sub + 1)]
    subs = [engine.ReplicateSweep(ctx, [n_snv] * (cuts[k + 1] - cuts[k]), columns, W_MAXS) for k in range(n_sub)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_sub)]
    host_ms = []

    def step_e2e():
        t0 = time.perf_counter()
        for k, sub in enumerate(subs):
            sub.parse_trees(newicks[cuts[k]:cuts[k + 1]])
            sub.enqueue(pinned[cuts[k] * n_snv:cuts[k + 1] * n_snv], FP, FN, stream=streams[k])
        rows = []
        for sub in subs:
            rows.extend(sub.rows())
        host_ms.append(1e3 * (time.perf_counter() - t0))
        return rows

    cur = torch.cuda.current_stream(dev)

    def step_e2e_timed():
        for st in streams:
            st.wait_stream(cur)                       # the timing events are recorded on the current stream
        rows = step_e2e()
        for st in streams:
            cur.wait_stream(st)
        return rows

    ms_e2e, rows = timed(step_e2e_timed, steps, 1)
    result["e2e"] = {"value": R * len(W_MAXS) * world / (ms_e2e / 1e3), "unit": "PT tests/s", "ms_per_step": ms_e2e,
                     "h2d_bytes_per_step": int(pinned.numel() + R * sweep.n_rows * (sweep.words + 2) * 4 + R * 16) * world,
                     "d2h_bytes_per_step": int(R * len(W_MAXS) * 64 + R * 12) * world, "sub_batches": n_sub,
                     "host_wall_ms_per_step": float(np.mean(host_ms[-steps:])),
                     "input": "packed genotype matrices in pinned host memory + Newick text of %d distinct trees" % R,
                     "output": "TSV model lines (scs_format_pt_rows)", "rows_equal_device_leg": bool(rows == ref_rows),
                     "vs_device_resident": ms_e2e / ms}
    if rank == 0 and cpu_seconds > 0:
        codes = synth.observe(trees[0], synth.throw_mutations(trees[0], n_snv, np.random.default_rng(3)), FN_TRUE, FP_TRUE, MS_TRUE,
                              np.random.default_rng(4))
        try:
            result["cpu_baseline"] = sweep_cpu_baseline(trees[0], columns, codes, tmp, cpu_seconds)
        except Exception as exc:          # (a reported extra: never lose the line over it)
            result["cpu_baseline"] = {"error": repr(exc)}
    for sub in subs:
        sub.close()
    sweep.close()
    return result


# ------------------------------------------------------------------ GPU arm
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
    from scsommerclock_b200 import engine, run_PT_test as P, sharding, synth
    rank, world, local = sharding.init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = engine.Context(local)
    n_cells, n_snv = args.cells, args.snvs
    pk, pk_kind = peaks()

    # ---- the tree (same on every rank), through the product's own tree code
    tree_s = make_tree(n_cells)
    names = ["tumcell%05d" % (i + 1) for i in range(n_cells)]
    columns = names + ["healthycell"]
    tmp = tempfile.mkdtemp(prefix="scs_bench_")
    tree_file = os.path.join(tmp, "bench_r%d.newick" % rank)
    with open(tree_file, "w") as f:
        f.write(tree_s.newick(names, "healthycell", scale=0.05))
    tree, outg, _, _ = P.read_tree(tree_file, np.array(columns))
    active = np.array([c != outg for c in columns], dtype=np.uint8)
    bits, words, nodes = P._clade_bits(tree, columns, active)
    n_rows = bits.shape[0]
    n_cols = len(columns)

    # ---- this rank's shard of SNV rows, generated on the device in the packed layout
    lo = torch.as_tensor(tree_s.lo[:-1], device=dev)
    hi = torch.as_tensor(tree_s.hi[:-1], device=dev)
    rate = torch.as_tensor(tree_s.branch_len[:-1], device=dev, dtype=torch.float32)
    packed, pitch, _ = synth.observe_packed_torch(lo, hi, torch.as_tensor(tree_s.leaf_rank), n_snv, FN_TRUE, FP_TRUE, MS_TRUE, dev,
                                                  seed=1000 + rank, rate=rate, extra_cols=1)
    torch.cuda.synchronize()
    FP, FN = FP_TRUE, FN_TRUE / 2        # what get_gt_tree would pass for a CellPhy log with these rates

    os.environ["SCS_PLACE_ALGO"] = args.algo
    plan = engine.Placement(ctx, n_snv, n_cols, n_rows)
    plan.set_clades_host(bits, words)
    row_keep = torch.empty(n_snv, dtype=torch.uint8, device=dev)
    y_dev = torch.zeros(n_rows, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def place_step():
        # row filter + per-cell missing counts + (interval path) the cell-order pass: one read of the packed matrix, counters
        # stay on the device; GEMM path: the two reduction kernels + the operand plane(s)
        plan.prepare(packed, pitch, active, True, stream=stream)
        plan.run(FP, FN, y_dev, stream=stream)
        sharding.allreduce_soft_weights(y_dev)

    # ---- one untimed pass fixes the LRT model (branch order, weights) for the timed steps
    place_step()
    n_kept, miss = plan.prepare_result(stream=stream)
    torch.cuda.synchronize()
    soft = y_dev.cpu().numpy()
    if world > 1:
        t = torch.tensor(np.concatenate([[float(n_kept)], miss * n_kept]), dtype=torch.float64, device=dev)
        sharding.allreduce_soft_weights(t)
        n_kept_all = float(t[0].item())
        miss = (t[1:] / n_kept_all).cpu().numpy()
    MS_i = {c: float(np.clip(miss[i], P.LAMBDA_MIN, 1 - FN - 0.2)) for i, c in enumerate(columns) if active[i]}
    st = {"soft": soft, "MS_i": MS_i, "n_muts": n_kept, "bits": bits, "words": words, "columns": columns, "weights": {}}
    P._apply_placement(tree, st)
    row_of = {id(n): i for i, n in enumerate(nodes)}
    name_row = {n.name: i for i, n in enumerate(nodes)}
    childs, weights, gathers = [], [], []
    for w_max in W_MAXS:
        P.add_br_weights(tree, FP, FN, MS_i, w_max, _state=st)
        Y, constr, init, wn, cols = P.get_model_data(tree)
        childs.append(constr.child_int)
        weights.append(wn[:, 0].copy())
        gathers.append(np.array([name_row[c] for c in cols], dtype=np.int32))
    lrt = engine.LrtPlan(ctx, childs)
    lrt.set_weights(weights)
    lrt.set_gather(np.concatenate(gathers), n_rows)
    lrt_out = torch.zeros((len(W_MAXS), 8), dtype=torch.float64, device=dev)

    def step():
        place_step()
        lrt.run(y_dev, lrt_out, stream=stream)
        return lrt_out.cpu()             # the step's result, read back to the host

    pinned = None
    streamer = None
    if not args.no_e2e:
        pinned = torch.empty(packed.shape, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(packed)
        torch.cuda.synchronize()
        # host-buffer path: the packed matrix is streamed from pinned host memory in SNV-row chunks
        # (H2D of chunk c+1 under the GEMMs of chunk c), see engine.StreamingPlacement
        streamer = engine.StreamingPlacement(ctx, n_snv, n_cols, n_rows, n_chunks=args.e2e_chunks or None, pitch=pitch)
        streamer.set_clades_host(bits, words)

    def step_e2e():
        streamer.run(pinned, pitch, active, True, FP, FN, y_dev)
        sharding.allreduce_soft_weights(y_dev)
        lrt.run(y_dev, lrt_out, stream=stream)
        return lrt_out.cpu()

    def timed(fn, steps, warmup, sampler=None, pl=None):
        pl = pl or plan
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        sharding.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        l0 = pl.info()["launches"]
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        gemm = []
        e0.record()
        for _ in range(steps):
            res = fn()
            gemm.append(pl.last_times())
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.stop_flag.set()
        ms = e0.elapsed_time(e1)
        # this library's kernels inside the timed region: plan launches (pre-pass, placement, reduce / operand expansion,
        # two GEMM passes, merge, reduce) + 1 LRT kernel per step
        n_launch = pl.info()["launches"] - l0 + steps
        return sharding.max_over_ranks(ms, dev), gemm, res, n_launch

    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, gemm_times, res, launches = timed(step, args.steps, args.warmup, sampler)
    plan_info = plan.info()
    ms_step = ms_total / args.steps
    placements = float(n_snv) * n_rows * world
    value = placements / (ms_step / 1e3)

    e2e = None
    if not args.no_e2e:
        y_ref = y_dev.clone()
        ms_e2e, _, _, _ = timed(step_e2e, args.steps, min(args.warmup, 1))
        e2e_dev = float((y_dev - y_ref).abs().max() / y_ref.abs().max())     # streamed == resident, to summation order
        e2e = {"value": placements / (ms_e2e / args.steps / 1e3), "unit": "placements/s",
               "h2d_bytes_per_step": int(packed.numel()) * world, "d2h_bytes_per_step": int(lrt_out.numel() * 8 + n_cols * 4 + 8) * world,
               "ms_per_step": ms_e2e / args.steps, "chunks": streamer.n_chunks, "max_rel_dev_vs_resident": e2e_dev}

    # ---- the same workload through the tcgen05 GEMM (north_star's formulation), when the headline ran the interval path
    gemm_leg = None
    if plan_info["algo"] == 2 and not args.no_gemm_leg:
        os.environ["SCS_PLACE_ALGO"] = "gemm"
        plan_g = engine.Placement(ctx, n_snv, n_cols, n_rows)
        plan_g.set_clades_host(bits, words)
        os.environ["SCS_PLACE_ALGO"] = args.algo
        y_g = torch.zeros_like(y_dev)
        y_iv = y_dev.clone()

        def step_gemm():
            plan_g.prepare(packed, pitch, active, True, stream=stream)
            plan_g.run(FP, FN, y_g, stream=stream)
            sharding.allreduce_soft_weights(y_g)
            lrt.run(y_g, lrt_out, stream=stream)
            return lrt_out.cpu()

        ms_g, g_times, _, _ = timed(step_gemm, 2, 1, pl=plan_g)
        gemm_leg = {"ms_per_step": ms_g / 2, "value": placements / (ms_g / 2 / 1e3), "times": np.array(g_times), "info": plan_g.info(),
                    "max_rel_dev_vs_interval": float((y_g - y_iv).abs().max() / y_iv.abs().max())}
        plan_g.close()
        del plan_g, y_g, y_iv

    rep = None
    if args.replicates > 0:
        del pinned, streamer
        plan.close()
        del packed
        torch.cuda.empty_cache()
        rep = replicate_sweep(ctx, dev, rank, world, args.replicates, args.rep_cells, args.rep_snvs, max(2, args.steps), 1,
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
            return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": (NCU_TRAFFIC_BYTES_CACHE if cached else NCU_TRAFFIC_BYTES) if (full_size and info["mode"] == 1) else None,
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
            # interval path: the packed 2-bit genotypes are read once (2 bits per (SNV, cell)); nothing else touches HBM but
            # the per-run partial sums.  The kernel is bound by instruction issue, not by HBM or the tensor pipe -- of the
            # two rooflines the contract names, HBM is the one its only off-chip traffic belongs to.
            t_k = float(g[:, 2].mean())                                      # the placement kernel (g[:, 0]: the cell-order pass before it)
            alg_bytes = float(n_snv) * pitch
            hbm_peak = pk.get("hbm_gbs_sustained", pk.get("hbm_gbs"))
            roofline = {"bound": "hbm", "achieved": alg_bytes / 1e9 / (t_k / 1e3), "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / 1e9 / (t_k / 1e3) / hbm_peak,
                        "traffic": NCU_TRAFFIC_BYTES_INTERVAL if full_size else None,
                        "peak_source": pk_kind + " HBM copy bandwidth",
                        "kernel": "scs_place_interval%s_kernel (prefix counts in shared memory, no GEMM)" % ("3" if plan_info["interval_variant"] == 3 else ""),
                        "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": t_k, "permute_ms": float(g[:, 0].mean()),
                        "reduce_ms": float(g[:, 3].mean()),
                        "entries_per_s": float(n_snv) * n_rows / (t_k / 1e3),
                        "actual_limiter": "instruction issue (ncu: issue slots 67 % busy, ALU pipe 52 %, 8 warps per scheduler, ~12 cycles between "
                                          "a warp's instructions); HBM traffic is ~50 GB/s because the algorithm needs O(cells + clade rows) work "
                                          "per SNV instead of the GEMM's O(cells x clade rows) -- see gemm_path for the tensor-pipe roofline of "
                                          "the same workload through the tcgen05 GEMM",
                        "smem_bytes": plan_info["interval_smem_bytes"], "threads": plan_info["interval_threads"], "grid": plan_info["interval_grid"]}
            if full_size:
              try:
                # the roofline that does bound it: issue slots (4 schedulers per SM, one warp instruction per cycle each)
                sm_count = getattr(ctx, "sm_count", 148)
                clk = ((sampler.summary().get("sm_mhz") if sampler else None) or pk.get("sm_max_mhz", 1965.0)) * 1e6
                inst = NCU_INTERVAL_WARP_INST_PER_SNV * n_snv
                roofline["issue_roofline"] = {"warp_instructions_per_launch": inst, "achieved_per_s": inst / (t_k / 1e3),
                                              "peak_per_s": sm_count * 4 * clk, "frac": inst / (t_k / 1e3) / (sm_count * 4 * clk),
                                              "note": "instruction count from the ncu capture of the same kernel (250k SNVs, scaled)"}
              except Exception as exc:          # (an explanatory extra: never lose the line over it)
                roofline["issue_roofline"] = {"error": repr(exc)}
        else:
            roofline = gemm_roofline(g, plan_info)
        gemm_path = None
        if gemm_leg is not None:
            gemm_path = {"ms_per_step": gemm_leg["ms_per_step"], "value": gemm_leg["value"], "unit": "placements/s", "steps": 2, "warmup": 1,
                         "max_rel_dev_soft_weights_vs_interval": gemm_leg["max_rel_dev_vs_interval"],
                         "roofline": gemm_roofline(gemm_leg["times"], gemm_leg["info"]), "tiling": gemm_leg["info"]}
        out = res.numpy()
        line = {
            "metric": METRIC, "value": value, "unit": "placements/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("u16 prefix counts (exact integers), f32 softmax terms on exact integer differences (LRT f64)" if plan_info["algo"] == 2 else
                      "fp8_e5m2 operands holding exact integer counts, f32 accumulate (epilogue f32, LRT f64)" if plan_info["mode"] == 1 else
                      "u8 operands, s32 accumulate (epilogue f32, LRT f64)"),
            "data": "synthetic",
            "config": {"workload": "coalescent %d cells x %d SNVs per GPU (BASELINE configs[3]), SNV rows sharded, one all-reduce of the branch weights, LRT for 10 w_max"
                                   % (n_cells, n_snv),
                       "cells": n_cells, "snvs_per_gpu": n_snv, "clade_rows": n_rows, "w_max": W_MAXS, "FN": FN_TRUE, "FP": FP_TRUE, "missing": MS_TRUE,
                       "algorithm": "interval path (tree-shaped clade masks: counts = differences of per-SNV prefix counts)" if plan_info["algo"] == 2 else "tcgen05 GEMM",
                       "l2": ("inputs larger than L2 (packed genotypes %.1f GB per GPU, read once per step)" % (float(n_snv) * pitch / 1e9)) if plan_info["algo"] == 2 else
                             ("inputs larger than L2 (operand plane(s) %.1f GB per GPU)" % ((1 if plan_info["mode"] == 1 else 2) * float(plan_info["M_pad"]) * plan_info["K_pad"] / 1e9)),
                       "tiling": plan_info},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "gemm_path": gemm_path,
            "pt_tests": {"per_step": len(W_MAXS), "LR": [float(v) for v in out[:, 0]], "p_value": [float(v) for v in out[:, 3]],
                         "newton_iterations": [int(v) for v in out[:, 6]], "status": [int(v) for v in out[:, 7]]},
            "clocks": sampler.summary() if sampler else None,
            "replicate_sweep": rep,
        }
        if world == 1 and not args.no_cpu_baseline:
            S = engine.unpack_clades(bits, n_cols)[:, :n_cells].astype(np.uint8)
            line["cpu_baseline"] = cpu_port_baseline(tree_s, np.ascontiguousarray(S), n_cells, args.cpu_seconds)
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    sharding.barrier()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
