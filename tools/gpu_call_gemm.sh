#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py tests/test_tail_gpu.py -m gpu -q --tb=short --timeout 200 -x > gpurun_out/gemm_tests.log 2>&1
tail -5 gpurun_out/gemm_tests.log
B="--steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
timeout -s KILL 200 python bench.py $B > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_x.json"))
print(round(d["value"]), round(d["ms_per_step"], 3), {k: (round(v["us_per_launch"], 1), v["launches_per_step"]) for k, v in d["kernels"].items()})
PY
timeout -s KILL 300 python bench.py --workload cfg4 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg4_x.json 2>/dev/null
cut -c1-200 gpurun_out/bench_cfg4_x.json
timeout -s KILL 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short --timeout 300 > gpurun_out/parity_gemm2.log 2>&1
tail -4 gpurun_out/parity_gemm2.log
