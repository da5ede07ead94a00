DVBT_B200_ACQ_TRACE=2 BENCH_QUICK=1 BENCH_SEED=8 timeout 600 python bench.py --steps 1 --warmup 3 > gpurun_out/q8.log 2>&1
grep "acq best" gpurun_out/q8.log | tail -1 | cut -c1-300
