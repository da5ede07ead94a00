#!/bin/bash
# round 2, GPU pass 12 (1 GPU): smoke(), parity incl. the soft-decision tests, the bench line with the soft legs
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r2_p12_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_p12_pytest.log
( time BENCH_VERBOSE=1 timeout 1500 python bench.py 2>gpurun_out/r2_p12_bench_err.log > gpurun_out/r2_p12_bench.json ) 2>&1 | tail -4
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p12_bench.json'))
print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "roof", d['e2e']['h2d_roof_gbs'], "frac", d['e2e'].get('frac_of_h2d_roof'))
print("one at a time", d['one_capture_at_a_time']['ms_per_capture'], "launches", d['gpu_launches'])
for k, v in d.get('per_config', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'error', 'ms_per_capture_one_at_a_time')}, v.get('e2e', {}).get('value'))
print("sweep cases", len(d.get('viterbi_sweep', {}).get('cases', [])) if isinstance(d.get('viterbi_sweep'), dict) else d.get('viterbi_sweep'))
for r in d['robustness'].get('awgn_tiled_capture', []): print("awgn", r)
print("soft", json.dumps(d['robustness'].get('soft_decision'))[:1500])
P
tail -5 gpurun_out/r2_p12_bench_err.log
