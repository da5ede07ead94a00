#!/bin/bash
# round 2, GPU pass 31 (8 GPUs): the driver's scaling command at N = 8, default flags, with vetted captures and parity over all ranks
mkdir -p gpurun_out
( time BENCH_VERBOSE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --steps 20 --warmup 3 \
  2>gpurun_out/r2_p31_bench_n8.err > gpurun_out/r2_p31_bench_n8.json ) 2>&1 | tail -3
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p31_bench_n8.json'))
print("N=8 value", d['value'], "per GPU", d['value'] / 8, "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "frac", d['e2e']['frac_of_h2d_roof'], "clocks", d['clocks'])
print("one at a time", d['one_capture_at_a_time'])
for k, v in d.get('per_config', {}).items():
    print(k, round(v['value']), round(v['ms_per_capture_one_at_a_time'], 2), round(v['e2e']['value']), v['parity_check'], v.get('capture_reseeds_max_over_ranks'))
vs = d['viterbi_sweep']; print("sweep", len(vs['cases']), all(c['parity'] for c in vs['cases']), isinstance(vs.get('soft_cases'), list) and all(c['parity'] for c in vs['soft_cases']))
P
grep "per-step wall" gpurun_out/r2_p31_bench_n8.err | awk '{mx=0; for(i=7;i<=NF;i++) if($i+0>mx) mx=$i+0; if (mx>3) print $0}' | cut -c1-200
