timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_v28.log
for sd in 1 3; do
DVBT_B200_ACQ_TRACE=1 BENCH_VERBOSE=1 BENCH_SEED=$sd timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_v28_seed${sd}_err.log > gpurun_out/bench_rx_v28_seed${sd}.json
cut -c1-200 gpurun_out/bench_rx_v28_seed${sd}.json
grep "acq batch" gpurun_out/bench_v28_seed${sd}_err.log | head -2 | cut -c1-330
grep "stages:\|two concurrent" gpurun_out/bench_v28_seed${sd}_err.log | cut -c1-200
done
