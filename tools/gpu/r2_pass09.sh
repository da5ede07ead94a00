#!/bin/bash
# round 2, GPU pass 9 (1 GPU): blocking vs spinning stream wait (quick legs), then the full bench line of the final build
mkdir -p gpurun_out
for b in 0 1; do
  echo "DVBT_B200_BLOCKING_WAIT=$b" | tee -a gpurun_out/r2_p09_wait.log
  DVBT_B200_BLOCKING_WAIT=$b BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep -E "bench quick" | cut -c1-150 | tee -a gpurun_out/r2_p09_wait.log
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_p09_pytest.log
( time BENCH_VERBOSE=1 timeout 1200 python bench.py 2>gpurun_out/r2_p09_bench_err.log > gpurun_out/r2_p09_bench.json ) 2>&1 | tail -4
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p09_bench.json'))
print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "frac roof", d['e2e'].get('frac_of_h2d_roof'))
print("one at a time", d['one_capture_at_a_time']['ms_per_capture'], "launches", d['gpu_launches'])
print("tx", {k: d['tx_generator'].get(k) for k in ('ms_per_capture', 'msamples_per_s', 'loop_ts_equals_source')})
for k, v in d.get('per_config', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'error', 'ms_per_capture_one_at_a_time')}, v['e2e']['value'])
di = d.get('drop_in_blocks', {})
print("drop-in 64:", di.get('pipelined_msamples_per_s'), di.get('serial_msamples_per_s'), {k: round(v['us_per_call']) for k, v in di.get('blocks', {}).items()})
b = di.get('items_per_call_512', {})
print("drop-in 512:", b.get('pipelined_msamples_per_s'), b.get('serial_msamples_per_s'), {k: round(v['us_per_call']) for k, v in b.get('blocks', {}).items()})
print("roofline", d['roofline']['frac'], d['roofline']['traffic'], [(r['kernel'][:20], round(r['frac'], 3)) for r in d['roofline_other']])
P
