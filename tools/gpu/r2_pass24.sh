#!/bin/bash
# round 2, GPU pass 24 (1 GPU): ncu launch list (per-launch times, serialised) of the bench command
mkdir -p gpurun_out
BENCH_QUICK=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_p24_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_p24_ncu.log 2>&1
grep -E "demod_stage1|demod_symbol|demod_scan|demod_vote" gpurun_out/r2_p24_launches.csv | head -12 | cut -d, -f5,15- | cut -c1-160
