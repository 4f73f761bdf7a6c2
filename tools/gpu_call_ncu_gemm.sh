export LSTMP_B200_BWD_COOP=0
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file gpurun_out/launches_gemm2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_gemm2.log 2>&1
tail -2 gpurun_out/ncu_gemm2.log
