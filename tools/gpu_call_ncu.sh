#!/bin/bash
# The two ncu passes of B200_PROFILING.md for the default bench workload (1 GPU).  Numbers under ncu are never bench values.
set -u
mkdir -p gpurun_out
R=${R:-2}
export LSTMP_B200_BWD_COOP=0   # ncu cannot replay a cooperative cluster launch; the grid is co-resident by construction
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r$R.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"lstmp_.*_kernel|gemm_hl_kernel|split_hl" -s 20 -c 24 -o gpurun_out/prof_r$R -f $B > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/prof_r$R.ncu-rep
