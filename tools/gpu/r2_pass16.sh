#!/bin/bash
# round 2, GPU pass 16 (1 GPU): demod scan on a high-priority side stream, placed before the wide kernel (A/B)
mkdir -p gpurun_out
for v in 0 1 0 1; do
  echo "DVBT_B200_DEMOD_SIDE_SCAN=$v" | tee -a gpurun_out/r2_p16_side_scan.log
  DVBT_B200_DEMOD_SIDE_SCAN=$v BENCH_VERBOSE=1 BENCH_QUICK=1 timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | grep -E "bench quick|stages:" | cut -c1-200 | tee -a gpurun_out/r2_p16_side_scan.log
done
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/r2_p16_pytest.log
