#!/bin/bash
# ncu --set full of two whole cfg3 steps (13 launches each of the repo's kernels), 1 GPU.  Numbers under ncu are never bench values.
set -u
mkdir -p gpurun_out
R=${R:-2}
export LSTMP_B200_BWD_COOP=0   # ncu cannot replay a cooperative cluster launch; the grid is co-resident by construction
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
timeout -s KILL ${NCU_TIMEOUT:-220} ncu --set full --clock-control none --import-source on \
  -k regex:"lstmp_.*_kernel|gemm_hl_kernel|split_multi|split_pair|splitk_reduce_multi" -s 26 -c 26 \
  -o gpurun_out/prof_r$R -f $B > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/prof_r$R.ncu-rep
