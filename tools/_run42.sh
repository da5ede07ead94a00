timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_v33.log
for sd in 8; do
DVBT_B200_ACQ_TRACE=1 BENCH_QUICK=1 BENCH_VERBOSE=1 BENCH_SEED=$sd timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/q.log 2>&1
grep "acq batch" gpurun_out/q.log | tail -1 | cut -c1-200; grep "stages:" gpurun_out/q.log | cut -c1-160; grep "bench quick" gpurun_out/q.log
done
BENCH_VERBOSE=1 timeout 600 python bench.py 2>gpurun_out/bench_v33_err.log > gpurun_out/bench_rx_v33.json
cut -c1-250 gpurun_out/bench_rx_v33.json; grep "bench rank" gpurun_out/bench_v33_err.log | grep -v per-step | cut -c1-250
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>gpurun_out/bench_v33_ref_err.log > gpurun_out/bench_rx_v33_reference.json
cut -c1-250 gpurun_out/bench_rx_v33_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rx_v33_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v33.log 2>&1
tail -1 gpurun_out/ncu_v33.log | cut -c1-120
