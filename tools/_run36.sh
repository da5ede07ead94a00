# how much shared memory should the ACS kernel leave to the kernels of the other captures in flight?
for d in 0 5 4 3 2; do
echo "ring depth $d"
BENCH_QUICK=1 DVBT_B200_VIT_DEPTH=$d timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "bench quick"
done
BENCH_VERBOSE=1 timeout 600 python bench.py 2>gpurun_out/bench_v30_err.log > gpurun_out/bench_rx_v30.json
cut -c1-300 gpurun_out/bench_rx_v30.json; grep "bench rank" gpurun_out/bench_v30_err.log | cut -c1-300
