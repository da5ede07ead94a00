#!/bin/bash
# round 2, GPU pass 10 (1 GPU): smoke(), parity, the bench line of the final build (drop-in leg on the raw C ABI)
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r2_p10_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_p10_pytest.log
( time BENCH_VERBOSE=1 timeout 1200 python bench.py 2>gpurun_out/r2_p10_bench_err.log > gpurun_out/r2_p10_bench.json ) 2>&1 | tail -4
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p10_bench.json'))
print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "roof", d['e2e']['h2d_roof_gbs'], "frac", d['e2e'].get('frac_of_h2d_roof'))
print("one at a time", d['one_capture_at_a_time']['ms_per_capture'], "launches", d['gpu_launches'])
print("tx", {k: d['tx_generator'].get(k) for k in ('ms_per_capture', 'msamples_per_s', 'loop_ts_equals_source', 'error')})
for k, v in d.get('per_config', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'error', 'ms_per_capture_one_at_a_time')}, v.get('e2e', {}).get('value'))
di = d.get('drop_in_blocks', {})
print("drop-in", di.get('error'))
print("drop-in 64:", di.get('pipelined_msamples_per_s'), di.get('serial_msamples_per_s'), {k: (round(v['us_per_call']), round(v['msamples_per_s'])) for k, v in di.get('blocks', {}).items()})
b = di.get('items_per_call_512', {})
print("drop-in 512:", b.get('error'), b.get('pipelined_msamples_per_s'), b.get('serial_msamples_per_s'), {k: (round(v['us_per_call']), round(v['msamples_per_s'])) for k, v in b.get('blocks', {}).items()})
print("roofline", d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['sass'].get('acs_sass_digest'), [(r['kernel'][:20], round(r['frac'], 3)) for r in d['roofline_other']])
print("robust", json.dumps(d['robustness'])[:600])
P
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
