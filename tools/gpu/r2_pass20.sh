#!/bin/bash
# round 2, GPU pass 20 (1 GPU): inner-codes kernel with table lookups - parity, quick bench x2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/r2_p20_pytest.log
for i in 1 2; do
  BENCH_VERBOSE=1 BENCH_QUICK=1 timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | grep -E "bench quick|stages:" | cut -c1-200 | tee -a gpurun_out/r2_p20_quick.log
done
