#!/bin/bash
# round 2, GPU pass 4: parity with the fused demod kernel, quick bench, ncu of the new kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2_p04_pytest.log
BENCH_VERBOSE=1 BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep -E "bench quick|stages:" | cut -c1-600 | tee gpurun_out/r2_p04_quick.log
BENCH_NO_CONFIGS=1 BENCH_NO_VITERBI_SWEEP=1 BENCH_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"demod_symbol|demod_vote|demod_scan" --launch-skip 9 -c 3 -o gpurun_out/r2_p04_demod -f python bench.py --steps 2 --warmup 3 > gpurun_out/r2_p04_ncu.log 2>&1
tail -2 gpurun_out/r2_p04_ncu.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/r2_p04_demod.ncu-rep "ncu --set full, fused demod_symbol_kernel (tools/gpu/r2_pass04.sh)" > gpurun_out/r2_p04_demod_ncu_summary.txt 2>&1
grep -E "^==|gpu__time_duration|dram__bytes|issue_active|warps_active|stalled_(long|barrier|short)|registers|l1tex__thr" gpurun_out/r2_p04_demod_ncu_summary.txt
