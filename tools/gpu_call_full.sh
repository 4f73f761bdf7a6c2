#!/bin/bash
# Full GPU check: smoke(), the whole -m gpu suite, then the default bench line (N=1), the cfg4 whole-network line, xent.
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short --timeout 400 > gpurun_out/all_gpu_tests.log 2>&1
tail -8 gpurun_out/all_gpu_tests.log
timeout -s KILL 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"]), d["e2e"] and d["e2e"]["h2d_bytes_per_step"])
print({k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["single_thread"])
PY
timeout -s KILL 600 python bench.py --workload cfg4 --steps 50 --warmup 5 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err
tail -3 gpurun_out/bench_cfg4_n1.err; cut -c1-260 gpurun_out/bench_cfg4_n1.json
timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_n1.json 2>/dev/null; cut -c1-250 gpurun_out/bench_ref_n1.json
