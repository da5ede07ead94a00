#!/bin/bash
# round 2, GPU pass 17b: memcheck over the stream / re-synchronisation tests, full log
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_stream_resync_gpu.py -m gpu -q -x > gpurun_out/r2_p17b_memcheck_full.log 2>&1
echo "memcheck rc=$?"
grep -n "Invalid\|at 0x\|by thread\|Address\|=========     in \|FAILED\|Error\|error" gpurun_out/r2_p17b_memcheck_full.log | head -40
