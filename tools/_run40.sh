timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_v32.log
for sd in 1 3 8; do
echo "seed $sd"
DVBT_B200_ACQ_TRACE=1 BENCH_QUICK=1 BENCH_VERBOSE=1 BENCH_SEED=$sd timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/q.log 2>&1
grep "acq batch" gpurun_out/q.log | tail -2 | cut -c1-200; grep "stages:" gpurun_out/q.log | cut -c1-160; grep "bench quick" gpurun_out/q.log
done
