#!/bin/bash
# round 2, session 8: refactored bench (strong scaling headline, configs[2] leg, real-reference CPU arms), quick shake-out then the full line
mkdir -p gpurun_out
python bench.py --cells 2000 --snvs 50000 --steps 2 --warmup 1 --replicates 64 --cpu-seconds 2 > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err; echo "small bench rc=$?"; tail -3 gpurun_out/r2_bench_small.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench5.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic", "ms_per_launch", "prepass_ms", "reduce_ms")}, d["roofline"].get("issue_roofline"))
print("pt", d["pt_tests"]["newton_iterations"], d["pt_tests"]["status"])
print("small", d["single_test_1k_cells"])
r = d["replicate_sweep"]
print({k: r[k] for k in r if k not in ("tiling",)})
print("gemm", d["gemm_path"] and {k: d["gemm_path"][k] for k in ("ms_per_step", "max_rel_dev_soft_weights_vs_interval")})
print("cpu", d.get("cpu_baseline"))
PY
