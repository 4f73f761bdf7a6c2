#!/bin/bash
# GPU call: validates the TMA-fed tcgen05 time loops (tests), then A/B benches.
#   gpurun --timeout 900 -- 'bash tools/gpu_call_tma.sh'
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short --timeout 120 -x -k "tma or cluster or tiny or cfg3_layer1 or cfg2_recipe" > gpurun_out/tma_tests.log 2>&1
tail -15 gpurun_out/tma_tests.log
B="--steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
for kp in 4 2 8 1; do
  LSTMP_B200_BWD_KP=$kp timeout -s KILL 150 python bench.py $B > gpurun_out/bench_tma_kp$kp.json 2> gpurun_out/bench_tma_kp$kp.err
done
LSTMP_B200_REC=1 timeout -s KILL 150 python bench.py $B > gpurun_out/bench_rec1.json 2>/dev/null
LSTMP_B200_TC_STAGGER=0 timeout -s KILL 150 python bench.py $B > gpurun_out/bench_tma_nostagger.json 2>/dev/null
python - <<'PY'
import json
for n in ("tma_kp4", "tma_kp2", "tma_kp8", "tma_kp1", "rec1", "tma_nostagger"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % n))
        print(n, round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed:", e)
PY
timeout -s KILL 600 python -m pytest tests -m gpu -q --tb=short --timeout 200 > gpurun_out/all_gpu_tests.log 2>&1
tail -15 gpurun_out/all_gpu_tests.log
