#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_parity_gpu.py tests/test_cpp_component.py tests/test_tail_gpu.py tests/test_dispatch_gpu.py -m gpu -q --tb=short --timeout 200 -k "not cfg5_full" > gpurun_out/tma_tests.log 2>&1
tail -6 gpurun_out/tma_tests.log
B="--steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
timeout -s KILL 150 python bench.py $B > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_x.json"))
print(round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
PY
