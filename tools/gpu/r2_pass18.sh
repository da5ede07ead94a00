#!/bin/bash
# round 2, GPU pass 18: memcheck after the acquisition-history fix (stream, acquisition, soft chain), then the whole parity suite
mkdir -p gpurun_out
T="tests/test_stream_resync_gpu.py tests/test_acq_gpu.py tests/test_soft_chain_gpu.py::test_noise_free_soft_chain_gives_the_hard_chain_ts tests/test_soft_chain_gpu.py::test_soft_stream_in_pieces_equals_one_shot tests/test_configs_capture_gpu.py"
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -q -x > gpurun_out/r2_p18_memcheck_full.log 2>&1
echo "memcheck rc=$?" | tee gpurun_out/r2_p18_memcheck.log
grep -E "passed|failed|ERROR SUMMARY|Invalid" gpurun_out/r2_p18_memcheck_full.log | head -8 | tee -a gpurun_out/r2_p18_memcheck.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_p18_pytest.log
