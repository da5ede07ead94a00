#!/bin/bash
# round 2, GPU pass 17 (1 GPU): compute-sanitizer memcheck + racecheck over the soft-decision and side-stream paths
mkdir -p gpurun_out
T="tests/test_soft_chain_gpu.py::test_noise_free_soft_chain_gives_the_hard_chain_ts tests/test_soft_chain_gpu.py::test_soft_stream_in_pieces_equals_one_shot tests/test_soft_decision_gpu.py::test_soft_repair_path_is_exact tests/test_stream_resync_gpu.py"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2_p17_memcheck.log
echo "memcheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/r2_p17_memcheck.log
T2="tests/test_soft_chain_gpu.py::test_soft_stream_in_pieces_equals_one_shot tests/test_soft_decision_gpu.py::test_soft_repair_path_is_exact"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest $T2 -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2_p17_racecheck.log
echo "racecheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/r2_p17_racecheck.log
