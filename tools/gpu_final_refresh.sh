#!/bin/bash
# last call of the round: the host-side changes on the GPU, one ncu capture of the final interval kernel
# (250k SNVs) and the N=1 bench line of the final code
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "batch or tsv or cli or streaming" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_default.json)"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:scs_place_interval -s 1 -c 1 -o gpurun_out/prof_interval_final250 -f python tools/check_pass2_modes.py 10000 250000 > gpurun_out/ncu_iv.log 2>&1; echo "ncu rc=$?"
