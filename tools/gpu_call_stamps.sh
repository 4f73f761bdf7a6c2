#!/bin/bash
# clock64 stamps of CTA 0 in the TMA-fed time loops (S=64): per-phase cycle budget of a timestep
set -u
mkdir -p gpurun_out
for g in 1 2; do
LSTMP_B200_TMA_GROUPS=$g LSTMP_B200_DEBUG=4 timeout -s KILL 120 python tools/stamps.py 64 > gpurun_out/stamps_fwd_tma_g$g.txt 2>&1
LSTMP_B200_TMA_GROUPS=$g LSTMP_B200_DEBUG=8 timeout -s KILL 120 python tools/stamps.py 64 > gpurun_out/stamps_bwd_tma_g$g.txt 2>&1
done
tail -60 gpurun_out/stamps_fwd_tma_g2.txt
