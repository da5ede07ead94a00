#!/bin/bash
# round 2, GPU pass 7 (1 GPU): parity incl. the transmit chain; full bench line (tx_generator, yield-polling waits)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_p07_pytest.log
( time BENCH_VERBOSE=1 timeout 1200 python bench.py 2>gpurun_out/r2_p07_bench_err.log > gpurun_out/r2_p07_bench.json ) 2>&1 | tail -4
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p07_bench.json'))
print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "frac roof", d['e2e'].get('frac_of_h2d_roof'))
print("one at a time", d['one_capture_at_a_time']['ms_per_capture'], "launches", d['gpu_launches'])
print("tx", json.dumps(d.get('tx_generator'))[:700])
for k, v in d.get('per_config', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'error')})
print("sweep ok", all(c['parity'] for c in d['viterbi_sweep']['cases']), [(c['rate'], c['channel_ber'], round(c['ms_per_decode'], 2)) for c in d['viterbi_sweep']['cases'] if c['channel_ber'] >= 1e-2])
P
tail -3 gpurun_out/r2_p07_bench_err.log | cut -c1-300
