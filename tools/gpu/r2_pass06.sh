#!/bin/bash
# round 2, GPU pass 6: block size of demod_symbol_kernel<false>; full bench line (in-flight stage times, drop-in legs)
mkdir -p gpurun_out
for nt in 128 192 256 384; do
  echo "DVBT_B200_DEMOD_THREADS=$nt" | tee -a gpurun_out/r2_p06_quick.log
  DVBT_B200_DEMOD_THREADS=$nt BENCH_VERBOSE=1 BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep -E "bench quick|stages:" | cut -c1-215 | tee -a gpurun_out/r2_p06_quick.log
done
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_p06_pytest.log
( time BENCH_VERBOSE=1 timeout 1200 python bench.py 2>gpurun_out/r2_p06_bench_err.log > gpurun_out/r2_p06_bench.json ) 2>&1 | tail -4
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p06_bench.json'))
print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "frac roof", d['e2e'].get('frac_of_h2d_roof'))
print("one at a time", d['one_capture_at_a_time']['ms_per_capture'])
print("stage", d['stage_ms']); print("in flight", d['stage_ms_in_flight'])
print("roofline_other", [(r['kernel'][:24], round(r['frac'], 3)) for r in d['roofline_other']])
for k, v in d.get('per_config', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'error', 'ms_per_capture_one_at_a_time')})
di = d.get('drop_in_blocks', {})
print("drop-in 64:", di.get('pipelined_msamples_per_s'), di.get('serial_msamples_per_s'), {k: round(v['us_per_call']) for k, v in di.get('blocks', {}).items()})
print("drop-in 512:", json.dumps(di.get('items_per_call_512'))[:900])
P
