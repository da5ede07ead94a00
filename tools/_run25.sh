# flakiness hunt: RS large batch repeated, racecheck of the RS kernels, then the suite
for i in 1 2 3 4 5 6 7 8; do timeout 300 python -m pytest tests/test_rs_gpu.py -x -q 2>&1 | tail -1; done | sort | uniq -c | tee gpurun_out/rs_repeat.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_rs_gpu.py -x -q -k "general_work or status" 2>&1 | grep -E "RACECHECK|passed|failed|Error|hazard" | head -20 | tee gpurun_out/rs_racecheck.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_v24.log
