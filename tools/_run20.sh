# round-1 GPU pass v20: parity tests, bench (fused FFT default, cuFFT variant), launch list
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_v20.log
timeout 600 python bench.py 2>gpurun_out/bench_v20_err.log | tee gpurun_out/bench_rx_v20.json | cut -c1-400
DVBT_B200_ACQ_CUFFT=1 timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_v20c_err.log | tee gpurun_out/bench_rx_v20_cufft.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rx_v20_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v20.log 2>&1
tail -3 gpurun_out/ncu_v20.log | cut -c1-200
