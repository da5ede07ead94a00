#!/bin/bash
# round 2, GPU pass 19: compute-sanitizer memcheck over the WHOLE GPU parity suite
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q > gpurun_out/r2_p19_memcheck_full.log 2>&1
echo "memcheck rc=$?" | tee gpurun_out/r2_p19_memcheck.log
grep -E "passed|failed|ERROR SUMMARY|Invalid|at .* in .*\.cu" gpurun_out/r2_p19_memcheck_full.log | sort | uniq -c | sort -rn | head -20 | tee -a gpurun_out/r2_p19_memcheck.log
