#!/bin/bash
# round 2, GPU pass 26 (8 GPUs): the driver's scaling command at N = 8 with default flags (full bench line), both arms
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29726 bench.py --gpus 8 --steps 20 --warmup 3 \
  2>gpurun_out/r2_p26_bench_n8.err > gpurun_out/r2_p26_bench_n8.json ) 2>&1 | tail -3
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p26_bench_n8.json'))
print("N=8 value", d['value'], "per GPU", d['value'] / 8, "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "roof", d['e2e']['h2d_roof_gbs'], "frac", d['e2e']['frac_of_h2d_roof'], "clocks", d['clocks'])
for k, v in d.get('per_config', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'error')})
vs = d['viterbi_sweep']; print("sweep", len(vs['cases']), all(c['parity'] for c in vs['cases']), isinstance(vs.get('soft_cases'), list) and all(c['parity'] for c in vs['soft_cases']))
P
tail -3 gpurun_out/r2_p26_bench_n8.err | cut -c1-300
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29727 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | cut -c1-400 ) 2>&1 | tail -5
