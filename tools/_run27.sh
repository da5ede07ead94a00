# two ranks over NCCL (config broadcast only), then the reference arm launched the same way
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_2gpu_err.log | tee gpurun_out/bench_rx_2gpu_v25.json | cut -c1-400
tail -3 gpurun_out/bench_2gpu_err.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 2>gpurun_out/bench_2gpu_ref_err.log | tee gpurun_out/bench_rx_2gpu_reference.json | cut -c1-400
tail -3 gpurun_out/bench_2gpu_ref_err.log | cut -c1-300
