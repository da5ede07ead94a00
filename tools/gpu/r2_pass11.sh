#!/bin/bash
# round 2, GPU pass 11: launch-geometry knobs with four captures in flight (quick legs)
mkdir -p gpurun_out
run() { echo "$1" | tee -a gpurun_out/r2_p11_knobs.log; env $1 BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep -E "bench quick" | cut -c1-140 | tee -a gpurun_out/r2_p11_knobs.log; }
run "X=0"
run "DVBT_B200_VIT_SM_DIV=2"
run "DVBT_B200_VIT_SM_DIV=4"
run "DVBT_B200_DEMOD_THREADS=128"
run "BENCH_STREAMS=3"
run "BENCH_STREAMS=6"
run "DVBT_B200_VIT_SM_DIV=2 BENCH_STREAMS=6"
