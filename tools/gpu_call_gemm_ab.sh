#!/bin/bash
# GEMM changes A/B: GEMM / tail / parity tests, then the N=1 bench with the grouped backward GEMMs on / off and with the
# tiled transposed split for every transposed operand of a group.
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py tests/test_tail_gpu.py tests/test_parity_gpu.py -m gpu -q --tb=short --timeout 300 -k "not cfg5_full" > gpurun_out/gemm_ab_tests.log 2>&1
tail -12 gpurun_out/gemm_ab_tests.log
run() {
  tag=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/gemm_ab_$tag.json 2> gpurun_out/gemm_ab_$tag.err
  tail -2 gpurun_out/gemm_ab_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/gemm_ab_$tag.json"))
print("$tag value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "launches", d.get("gpu_launches"))
print({k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
PY
}
run group1 LSTMP_B200_GROUP_GEMMS=1
run group1_tiled0 LSTMP_B200_GROUP_GEMMS=1 LSTMP_B200_SPLIT_TILED_MIN=1
run group0 LSTMP_B200_GROUP_GEMMS=0
