#!/bin/bash
# round 2, GPU pass 32 (8 GPUs): one clock sampler per job instead of one per rank - the headline legs at N = 8, twice
mkdir -p gpurun_out
for i in 1 2; do
BENCH_QUICK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2973$i bench.py --gpus 8 --steps 20 --warmup 3 2>&1 | grep -E "bench quick" | cut -c1-200 | tee -a gpurun_out/r2_p32_n8_quick.log
done
