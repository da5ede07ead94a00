#!/bin/bash
# round 2, GPU pass 3b (8 GPUs): the host->device copy roof of the box at N = 2, 4, 8 and the headline bench at N = 4 and 8
mkdir -p gpurun_out
for n in 1 2 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n tools/gpu/h2d_roof.py 2>&1 | grep -E "h2d roof|d2h" | tee -a gpurun_out/r2_p03_h2d_roof_multi.log
done
nvidia-smi topo -m > gpurun_out/r2_p03_topo8.txt 2>&1
for n in 4 8; do
  BENCH_NO_CONFIGS=1 BENCH_NO_VITERBI_SWEEP=1 BENCH_NO_ROBUSTNESS=1 BENCH_NO_DROPIN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
    --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/r2_p03_bench_n$n.err > gpurun_out/r2_p03_bench_n$n.json
  python - <<P
import json
d = json.load(open('gpurun_out/r2_p03_bench_n$n.json'))
print("N=$n value", d['value'], "e2e", d['e2e']['value'], "roof GB/s", d['e2e']['h2d_roof_gbs'], "achieved", d['e2e']['h2d_achieved_gbs'], "frac", d['e2e']['frac_of_h2d_roof'], "parity", d['parity_check'])
P
done
