for smp in 1 0; do
BENCH_VERBOSE=1 BENCH_SAMPLER=$smp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$smp bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_2gpu_s${smp}_err.log | cut -c1-200
grep "bench rank" gpurun_out/bench_2gpu_s${smp}_err.log | head -6 | cut -c1-250
done
