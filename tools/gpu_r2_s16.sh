#!/bin/bash
# round 2, session 16: the whole GPU suite, smoke(), the full bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rf > gpurun_out/r2_pytest16.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest16.log | cut -c1-400
python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench6.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench6.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "frac", "traffic", "ms_per_launch", "prepass_ms", "prepass_GBps", "reduce_ms")}, d["roofline"].get("issue_roofline"))
print("small", d["single_test_1k_cells"])
r = d["replicate_sweep"]
print({k: r[k] for k in r if k not in ("tiling", "e2e", "cpu_baseline")}); print(r["e2e"]); print(r.get("cpu_baseline"))
print("gemm", d["gemm_path"] and {k: d["gemm_path"][k] for k in ("ms_per_step", "max_rel_dev_soft_weights_vs_interval")})
print("cpu", d.get("cpu_baseline"))
PY
