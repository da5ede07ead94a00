#!/bin/bash
# round 2, GPU pass 27 (8 GPUs): the driver's scaling command at N = 8 again (clock sampler at 200 ms; per-step wall times of every leg)
mkdir -p gpurun_out
( time BENCH_VERBOSE=1 BENCH_NO_VITERBI_SWEEP=1 BENCH_NO_DROPIN=1 BENCH_NO_TX=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29728 bench.py --gpus 8 --steps 20 --warmup 3 \
  2>gpurun_out/r2_p27_bench_n8.err > gpurun_out/r2_p27_bench_n8.json ) 2>&1 | tail -3
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p27_bench_n8.json'))
print("N=8 value", d['value'], "per GPU", d['value'] / 8, "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "frac", d['e2e']['frac_of_h2d_roof'], "clocks", d['clocks'])
for k, v in d.get('per_config', {}).items():
    print(k, round(v['value']), round(v['ms_per_capture_one_at_a_time'], 2), round(v['e2e']['value']), v['parity_check'])
P
grep "per-step wall" gpurun_out/r2_p27_bench_n8.err | awk '{mx=0; for(i=7;i<=NF;i++) if($i>mx) mx=$i; print $3, $4, "n=" NF-6, "max=" mx}' | sort | uniq -c | sort -k5 | tail -12
