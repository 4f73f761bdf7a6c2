#!/bin/bash
set -u
mkdir -p gpurun_out
N=2
run() {
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 100 --warmup 10 --no-e2e --no-secondary > gpurun_out/nccl_x.json 2> gpurun_out/nccl_x.err
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/nccl_x.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
LSTMP_B200_BENCH_NCCL_ALGO=none run default 29741
run Ring-by-default 29743
