#!/bin/bash
# round 2, GPU pass 21: ncu --set full of the mid-size kernels of the step (inner codes, stage 1, acquisition tables, RS)
mkdir -p gpurun_out
BENCH_QUICK=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"rx_inner_codes|demod_stage1|acq_pass2|acq_lambda|rs_decode" --launch-skip 10 -c 5 -o gpurun_out/r2_p21_mid -f python bench.py --steps 2 --warmup 3 > gpurun_out/r2_p21_ncu.log 2>&1
tail -2 gpurun_out/r2_p21_ncu.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/r2_p21_mid.ncu-rep "ncu --set full --clock-control none, mid-size kernels of the configs[1] step (tools/gpu/r2_pass21.sh)" > gpurun_out/r2_p21_mid_ncu_summary.txt 2>&1
grep -E "^==|gpu__time_duration|dram__bytes|dram__thr|issue_active|warps_active|stalled_|registers|l1tex__thr|bank_conflicts|inst_executed.sum" gpurun_out/r2_p21_mid_ncu_summary.txt
