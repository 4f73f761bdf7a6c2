#!/bin/bash
set -u
mkdir -p gpurun_out
B="--steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
for gt in default 4 2; do
  if [ $gt = default ]; then unset LSTMP_B200_TMA_GROUP_TILES; else export LSTMP_B200_TMA_GROUP_TILES=$gt; fi
  timeout -s KILL 200 python bench.py $B > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - $gt <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_x.json"))
print("gt", sys.argv[1], round(d["value"]), round(d["ms_per_step"], 3), {k: (round(v["us_per_launch"], 1), v["launches_per_step"]) for k, v in d["kernels"].items()})
PY
done
unset LSTMP_B200_TMA_GROUP_TILES
export LSTMP_B200_BWD_COOP=0
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 60 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_cfg4.log 2>&1
tail -2 gpurun_out/ncu_cfg4.log | cut -c1-100
