#!/bin/bash
# round 2, GPU pass 15 (1 GPU): demod scan on a side stream (A/B), clock sampler started before the warm-up; smoke, parity, bench
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/r2_p15_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_p15_pytest.log
for v in 0 1 0 1; do
  echo "DVBT_B200_DEMOD_SIDE_SCAN=$v" | tee -a gpurun_out/r2_p15_side_scan.log
  DVBT_B200_DEMOD_SIDE_SCAN=$v BENCH_QUICK=1 timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | grep -E "bench quick" | cut -c1-150 | tee -a gpurun_out/r2_p15_side_scan.log
done
( time BENCH_VERBOSE=1 timeout 1500 python bench.py 2>gpurun_out/r2_p15_bench_err.log > gpurun_out/r2_p15_bench.json ) 2>&1 | tail -4
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p15_bench.json'))
print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "frac", d['e2e'].get('frac_of_h2d_roof'))
print("one at a time", d['one_capture_at_a_time']['ms_per_capture'], "launches", d['gpu_launches'], "clocks", d['clocks'])
print("stage", d['stage_ms'])
for k, v in d.get('per_config', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'error', 'ms_per_capture_one_at_a_time')}, v.get('e2e', {}).get('value'))
print("soft", json.dumps(d['robustness'].get('soft_decision'))[:600])
P
grep -E "stages:" gpurun_out/r2_p15_bench_err.log | head -2 | cut -c1-300
