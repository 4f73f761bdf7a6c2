#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short --timeout 120 -x -k "not cfg5_full" > gpurun_out/tma_tests.log 2>&1
tail -4 gpurun_out/tma_tests.log
B="--steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
for cfg in 2:2 2:4 2:1; do
  IFS=: read gb gtb <<< "$cfg"
  LSTMP_B200_TMA_GROUPS_BWD=$gb LSTMP_B200_TMA_GROUP_TILES_BWD=$gtb timeout -s KILL 150 python bench.py $B > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - $cfg <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_x.json"))
    print("Gb:gtb", sys.argv[1], round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items() if "recurrent" in k}, d["engine"]["bwd_ctas"])
except Exception as e:
    print("failed", e)
PY
done
