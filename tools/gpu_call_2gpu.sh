#!/bin/bash
# 2-GPU check: parity test (both all-reduce back ends), then the bench lines under torchrun.
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short --timeout 400 > gpurun_out/mgpu_tests.log 2>&1
tail -8 gpurun_out/mgpu_tests.log
N=${N:-2}
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"]))
print("mgpu_parity", d["mgpu_parity"])
print("secondary", d["secondary"])
PY
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --workload cfg4 --steps 50 --warmup 5 > gpurun_out/bench_cfg4_n$N.json 2> gpurun_out/bench_cfg4_n$N.err
tail -3 gpurun_out/bench_cfg4_n$N.err; cut -c1-300 gpurun_out/bench_cfg4_n$N.json
