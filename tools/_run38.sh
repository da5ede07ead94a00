timeout 600 python -m pytest tests/test_resample_gpu.py -x -q 2>&1 | tail -5
for nq in 1 2 4; do
echo "resampler quads per thread: $nq"
BENCH_QUICK=1 BENCH_VERBOSE=1 DVBT_B200_RESAMPLE_NQ=$nq timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "bench quick\|stages:" | cut -c1-200
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_v31.log
