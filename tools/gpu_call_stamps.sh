#!/bin/bash
# clock64 stamps of CTA 0 in the bulk-copy-fed time loops (S=64, engine defaults): per-phase cycle budget of a timestep
set -u
mkdir -p gpurun_out
T=${T:-4}
LSTMP_B200_DEBUG=4 timeout -s KILL 120 python tools/stamps.py 64 $T > gpurun_out/stamps_fwd_tma.txt 2>&1
LSTMP_B200_DEBUG=8 timeout -s KILL 120 python tools/stamps.py 64 $T > gpurun_out/stamps_bwd_tma.txt 2>&1
grep -c . gpurun_out/stamps_bwd_tma.txt
