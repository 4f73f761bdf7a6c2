#!/bin/bash
# GPU call: validates the TMA-fed tcgen05 time loops (tests), then A/B benches and stamps.
#   gpurun --timeout 900 -- 'bash tools/gpu_call_tma.sh'
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short --timeout 120 -x -k "tma or cluster or tiny or cfg3_layer1 or cfg2_recipe or determinism or slot" > gpurun_out/tma_tests.log 2>&1
tail -15 gpurun_out/tma_tests.log
B="--steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
for cfg in ${CFGS:-2:4 2:2 1:8 1:4}; do
  set -- ${cfg%%:*} ${cfg##*:}
  LSTMP_B200_TMA_GROUPS=$1 LSTMP_B200_BWD_KP=$2 timeout -s KILL 150 python bench.py $B > gpurun_out/bench_tma_g$1_kp$2.json 2> gpurun_out/bench_tma_g$1_kp$2.err
  python - "$1" "$2" <<'PY'
import json, sys
n = "g%s_kp%s" % (sys.argv[1], sys.argv[2])
try:
    d = json.load(open("gpurun_out/bench_tma_%s.json" % n))
    print(n, round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
except Exception as e:
    print(n, "failed:", e)
PY
done
for g in 1 2; do
LSTMP_B200_TMA_GROUPS=$g LSTMP_B200_DEBUG=4 timeout -s KILL 120 python tools/stamps.py 64 > gpurun_out/stamps_fwd_tma_g$g.txt 2>&1
LSTMP_B200_TMA_GROUPS=$g LSTMP_B200_DEBUG=8 timeout -s KILL 120 python tools/stamps.py 64 > gpurun_out/stamps_bwd_tma_g$g.txt 2>&1
done
tail -40 gpurun_out/stamps_fwd_tma_g1.txt
