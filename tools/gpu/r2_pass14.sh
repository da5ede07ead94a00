#!/bin/bash
# round 2, GPU pass 14 (8 GPUs): the resident leg at N = 8 with other numbers of captures in flight and both wait modes
mkdir -p gpurun_out
BENCH_QUICK=1 BENCH_QUICK_SWEEP="4s,3s,2s,6s,4b,6b,8b,4s" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
  --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 8 --steps 10 --warmup 3 2>&1 | grep -E "bench quick" | cut -c1-260 | tee gpurun_out/r2_p14_n8_sweep.log
nproc | tee -a gpurun_out/r2_p14_n8_sweep.log
