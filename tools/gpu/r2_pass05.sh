#!/bin/bash
# round 2, GPU pass 5: stage1 + TMA-fed symbol kernel (default) vs the fully fused kernel; parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2_p05_pytest.log
for f in 0 1; do
  echo "DVBT_B200_DEMOD_FUSED=$f" | tee -a gpurun_out/r2_p05_quick.log
  DVBT_B200_DEMOD_FUSED=$f BENCH_VERBOSE=1 BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep -E "bench quick|stages:" | cut -c1-330 | tee -a gpurun_out/r2_p05_quick.log
done
BENCH_NO_CONFIGS=1 BENCH_NO_VITERBI_SWEEP=1 BENCH_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"demod_symbol|demod_stage1" --launch-skip 6 -c 2 -o gpurun_out/r2_p05_demod -f python bench.py --steps 2 --warmup 3 > gpurun_out/r2_p05_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_p05_demod.ncu-rep "ncu --set full, demod_stage1_kernel + demod_symbol_kernel<false> (tools/gpu/r2_pass05.sh)" > gpurun_out/r2_p05_demod_ncu_summary.txt 2>&1
grep -E "^==|gpu__time_duration|dram__bytes|issue_active|warps_active|stalled_(long|barrier|short)|registers|inst_executed.sum" gpurun_out/r2_p05_demod_ncu_summary.txt
