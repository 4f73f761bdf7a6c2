#!/bin/bash
# GEMM changes A/B: GEMM / tail / parity tests, then the N=1 bench with the dual weight-gradient launch on and off.
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py tests/test_tail_gpu.py tests/test_parity_gpu.py -m gpu -q --tb=short --timeout 300 -k "not cfg5_full" > gpurun_out/gemm_ab_tests.log 2>&1
tail -6 gpurun_out/gemm_ab_tests.log
for dual in 1 0; do
  LSTMP_B200_DUAL_WGRAD=$dual timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/gemm_ab_dual$dual.json 2> gpurun_out/gemm_ab_dual$dual.err
  tail -2 gpurun_out/gemm_ab_dual$dual.err
  python - <<PY
import json
d = json.load(open("gpurun_out/gemm_ab_dual$dual.json"))
print("dual=$dual value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4))
print({k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
PY
done
