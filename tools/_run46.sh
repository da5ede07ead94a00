timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_v35.log
BENCH_QUICK=1 BENCH_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "bench quick\|stages:" | cut -c1-260
