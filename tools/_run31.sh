DVBT_B200_ACQ_TRACE=2 BENCH_SEED=2 timeout 600 python bench.py --steps 1 --warmup 3 2>gpurun_out/bench_trace2_seed2_err.log | cut -c1-100
grep "acq sym" gpurun_out/bench_trace2_seed2_err.log | head -45 | cut -c1-330
