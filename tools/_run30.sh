# acquisition trace for the slow captures (seeds 2, 3)
for sd in 2 3; do
DVBT_B200_ACQ_TRACE=1 BENCH_VERBOSE=1 BENCH_SEED=$sd timeout 600 python bench.py --steps 1 --warmup 3 2>gpurun_out/bench_trace_seed${sd}_err.log | cut -c1-100
grep "acq batch" gpurun_out/bench_trace_seed${sd}_err.log | head -12 | cut -c1-400
done
