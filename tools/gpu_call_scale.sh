#!/bin/bash
# Scaling on one 8-GPU box: N = 1, 2, 4, 8 back to back (cfg3, weak), then configs[3] as a whole network at N = 8.
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
done
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --workload cfg4 --steps 50 --warmup 5 > gpurun_out/scale_cfg4_n8.json 2> gpurun_out/scale_cfg4_n8.err
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("gpurun_out/scale_n%d.json" % n).read().strip().splitlines()[-1])
        if n == 1:
            base = d["value"]
        print("N=%d value %.0f ms/step %.3f eff %.3f e2e %.0f mgpu_parity %s" % (n, d["value"], d["ms_per_step"], d["value"] / (n * base), d["e2e"]["value"], (d.get("mgpu_parity") or {}).get("ok")))
        if d.get("secondary"):
            print("   secondary", {k: (round(v.get("value", 0)), round(v.get("ms_per_step", 0), 3)) for k, v in d["secondary"].items()})
    except Exception as e:
        print("N=%d failed: %s" % (n, e))
try:
    d = json.loads(open("gpurun_out/scale_cfg4_n8.json").read().strip().splitlines()[-1])
    print("cfg4 N=8 value %.0f ms/step %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("cfg4 N=8 failed", e)
PY
