#!/bin/bash
set -u
mkdir -p gpurun_out
N=8
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
run default 29741
NCCL_ALGO=NVLS run NVLS 29742
NCCL_ALGO=Ring run Ring 29743
NCCL_ALGO=Tree run Tree 29744
NCCL_MIN_NCHANNELS=32 run minch32 29745
NCCL_MAX_NCHANNELS=8 run maxch8 29746
