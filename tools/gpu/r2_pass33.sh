#!/bin/bash
# round 2, GPU pass 33 (8 GPUs): longer warm-up of the in-flight leg - the headline legs at N = 8, twice
mkdir -p gpurun_out
for i in 1 2; do
BENCH_QUICK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2974$i bench.py --gpus 8 --steps 20 --warmup 3 2>&1 | grep -E "bench quick" | cut -c1-200 | tee -a gpurun_out/r2_p33_n8_quick.log
done
