# radix-16 fused derotation+FFT kernel: parity suite, bench, launch list
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_v25.log
timeout 600 python bench.py 2>gpurun_out/bench_v25_err.log | tee gpurun_out/bench_rx_v25.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/rx_v25_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v25.log 2>&1
