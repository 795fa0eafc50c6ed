#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"; tail -5 gpurun_out/r2_bench_n$N.err
python - $N <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_bench_n%s.json" % sys.argv[1]))
print("N", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "launches", d["gpu_launches"])
print("roofline", {k: d["roofline"][k] for k in ("ms_per_launch", "prepass_ms", "reduce_ms")})
print("weak", d["weak_scaling"])
r = d["replicate_sweep"]
print({k: r[k] for k in ("ms_per_step", "pt_tests_per_s")}, r["e2e"]["value"], r["e2e"]["ms_per_step"])
PY
