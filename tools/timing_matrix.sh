#!/bin/bash
# timing experiments: which part of a step costs what (results with LSTMP_B200_DEBUG != 0 are numerically wrong)
mkdir -p gpurun_out
for wl in cfg2 cfg3-layer1; do
 for dbg in 0 1 2 3; do
  for ctas in 148 74 37; do
   if [ "$wl" = "cfg3-layer1" ] && [ "$ctas" != "148" ]; then continue; fi
   out=$(LSTMP_B200_DEBUG=$dbg LSTMP_B200_MAX_CTAS=$ctas timeout -s KILL 120 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print('ms/step %.3f fwd %.1f bwd %.1f ngroups %d' % (d['ms_per_step'], k['fwd_recurrent']['us_per_launch'], k['bwd_recurrent']['us_per_launch'], d['config']['decomposition']['ngroups']))")
   echo "$wl dbg=$dbg ctas=$ctas : $out"
  done
 done
done
