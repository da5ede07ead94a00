# h16 v2 (VIMNMX3 argmax, compile-time ring stride, G/F stores between segments): parity, sweep, bench
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_v23.log
SWEEP_ONLY="384,0,384;512,0,512;256,0,256;384,0,128;384,5,384;384,3,384" SWEEP_REPS=5 timeout 300 python tools/sweep_acs.py 2>&1 | tail -7 | tee gpurun_out/sweep_h16_v3.txt
timeout 600 python bench.py 2>gpurun_out/bench_v23_err.log | tee gpurun_out/bench_rx_v23.json | cut -c1-300
