#!/bin/bash
# round 2, GPU pass 34 (1 GPU): the default bench line with the final bench.py (library unchanged since pass 30)
mkdir -p gpurun_out
timeout 100 python bench.py 2>gpurun_out/r2_p34_bench_err.log > gpurun_out/r2_p34_bench.json
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p34_bench.json'))
print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "frac", d['e2e'].get('frac_of_h2d_roof'), "clocks", d['clocks'], "launches", d['gpu_launches'])
P
