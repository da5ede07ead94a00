# two ranks over NCCL (config broadcast only), 4 captures in flight per rank
BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_2gpu_v33_err.log > gpurun_out/bench_rx_2gpu_v33.json
cut -c1-300 gpurun_out/bench_rx_2gpu_v33.json
grep "bench rank" gpurun_out/bench_2gpu_v33_err.log | grep -v per-step | cut -c1-200
nproc
