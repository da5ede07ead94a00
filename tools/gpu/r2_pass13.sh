#!/bin/bash
# round 2, GPU pass 13 (1 GPU): the ncu launch list of the bench command (per-launch times, serialised) for the final build
mkdir -p gpurun_out
BENCH_QUICK=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_p13_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_p13_ncu.log 2>&1
tail -3 gpurun_out/r2_p13_ncu.log
wc -l gpurun_out/r2_p13_launches.csv
