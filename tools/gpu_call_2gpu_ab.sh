#!/bin/bash
set -u
mkdir -p gpurun_out
N=${N:-2}
for f in 1 0 1 0; do
LSTMP_B200_FUSED_EXCHANGE=$f timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$f bench.py --gpus $N --steps 100 --warmup 10 --no-e2e > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
python - $f <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
print("fused", sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["secondary"].items()})
PY
done
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/bench_ab1.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ab1.json"))
print("N=1 on the same box: value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3))
PY
