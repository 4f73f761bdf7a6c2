#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short --timeout 200 -x > gpurun_out/gemm_tests.log 2>&1
tail -5 gpurun_out/gemm_tests.log
B="--steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
for cfg in 2:4 2:8 1:8; do
  LSTMP_B200_TMA_GROUPS=${cfg%%:*} LSTMP_B200_BWD_KP=${cfg##*:} timeout -s KILL 150 python bench.py $B > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - $cfg <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_x.json"))
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"], 3), {k: (round(v["us_per_launch"], 1), v["launches_per_step"]) for k, v in d["kernels"].items()}, d["engine"])
except Exception as e:
    print("failed", e)
PY
done
timeout -s KILL 600 python -m pytest tests/test_parity_gpu.py tests/test_tail_gpu.py -m gpu -q --tb=short --timeout 300 > gpurun_out/parity_gemm2.log 2>&1
tail -5 gpurun_out/parity_gemm2.log
