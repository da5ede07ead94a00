# is the slower rank 1 of the 2-GPU run a property of its capture (seed 2)?  one GPU, both seeds
for sd in 1 2 3; do
BENCH_VERBOSE=1 BENCH_SEED=$sd timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_seed${sd}_err.log | cut -c1-200
grep "bench rank" gpurun_out/bench_seed${sd}_err.log | cut -c1-900
done
