timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_v29.log
for cfg in "1 2" "3 3" "1 4"; do
set -- $cfg
DVBT_B200_ACQ_TRACE=1 BENCH_VERBOSE=1 BENCH_SEED=$1 BENCH_STREAMS=$2 timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_v29_seed$1_err.log > gpurun_out/bench_rx_v29_seed$1_s$2.json
cut -c1-200 gpurun_out/bench_rx_v29_seed$1_s$2.json
grep "acq batch" gpurun_out/bench_v29_seed$1_err.log | head -2 | cut -c1-330
grep "stages:\|concurrent" gpurun_out/bench_v29_seed$1_err.log | cut -c1-200
done
