#!/bin/bash
# First GPU call of the next round (about 5 minutes of box time): settles the open question of DESIGN.md 3.1b / 6.0.
#   gpurun --timeout 600 -- 'bash tools/round2_first_call.sh'
# Needs: make -C kaldi-lstm_b200/csrc all sts ; the two tools/_build binaries (nvcc lines in their headers).
set -u
mkdir -p gpurun_out
timeout 120 tools/_build/tc_pipeline_bench > gpurun_out/tc_pipeline.log 2>&1; cat gpurun_out/tc_pipeline.log
STS=$PWD/kaldi-lstm_b200/_lib/liblstmp_b200_sts.so
export LSTMP_B200_EXPERIMENTAL=1   # also run the loader-variant tests
LSTMP_B200_LIB=$STS timeout -s KILL 300 python -m pytest tests -m gpu -q --tb=short --timeout 150 > gpurun_out/sts_tests.log 2>&1
tail -3 gpurun_out/sts_tests.log
# GEMM ping-pong loaders (mode 2) are selected per process: own pytest process, then a bench
LSTMP_B200_GEMM_LOADER=2 timeout -s KILL 200 python -m pytest tests/test_gemm_gpu.py tests/test_parity_gpu.py -m gpu -q --tb=short --timeout 150 > gpurun_out/gemm_pp_tests.log 2>&1
tail -3 gpurun_out/gemm_pp_tests.log
LSTMP_B200_GEMM_LOADER=2 timeout -s KILL 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/bench_gemm_pp.json 2>/dev/null
LSTMP_B200_GEMM_LOADER=2 LSTMP_B200_TC_LOADER=2 LSTMP_B200_LIB=$STS timeout -s KILL 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/bench_all_new.json 2>/dev/null
for ld in 0 2; do
  LSTMP_B200_TC_LOADER=$ld timeout -s KILL 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary \
    > gpurun_out/bench_loader_$ld.json 2>/dev/null
  LSTMP_B200_LIB=$STS LSTMP_B200_TC_LOADER=$ld timeout -s KILL 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary \
    > gpurun_out/bench_sts_loader_$ld.json 2>/dev/null
done
for lib in "" "$STS"; do
  LSTMP_B200_LIB=$lib timeout -s KILL 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary \
    > gpurun_out/bench_lib_$( [ -z "$lib" ] && echo default || echo sts ).json 2>/dev/null
done
python - <<'PY'
import json
for n in ("lib_default", "lib_sts", "loader_0", "loader_2", "sts_loader_0", "sts_loader_2", "gemm_pp", "all_new"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % n))
        print(n, round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed:", e)
PY
