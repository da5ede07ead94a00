#!/bin/bash
# round 2, GPU pass 25 (1 GPU): smoke(), parity, the full bench line (soft sweep, soft legs with stage times), reference arm, launch list
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/r2_p25_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_p25_pytest.log
( time BENCH_VERBOSE=1 timeout 1500 python bench.py 2>gpurun_out/r2_p25_bench_err.log > gpurun_out/r2_p25_bench.json ) 2>&1 | tail -4
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p25_bench.json'))
print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], "e2e", d['e2e']['value'], "frac", d['e2e'].get('frac_of_h2d_roof'))
print("one at a time", d['one_capture_at_a_time']['ms_per_capture'], "launches", d['gpu_launches'], "clocks", d['clocks'])
for k, v in d.get('per_config', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'error', 'ms_per_capture_one_at_a_time')}, v.get('e2e', {}).get('value'))
vs = d['viterbi_sweep']
print("sweep", len(vs['cases']), all(c['parity'] for c in vs['cases']))
sc = vs.get('soft_cases')
if isinstance(sc, list):
    for c in sc: print("soft", c['rate'], c['noise_sigma'], round(c['hard_decision_ber_of_the_input'], 4), round(c['ms_per_decode'], 3), round(c['acs_kernel_ms'], 3), c['repaired_chunks'], c['decoded_byte_errors_vs_source'], c['parity'])
else: print("soft", sc)
for r in d['robustness'].get('soft_decision', []): print("robust", {k: r[k] for k in ('snr_db', 'decisions', 'ms_per_capture', 'viterbi_repaired_chunks', 'packets_equal_to_source')}, r.get('stage_ms'))
print("in flight", json.dumps(d.get('in_flight_sweep'))[:400])
P
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
BENCH_QUICK=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_p25_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_p25_ncu.log 2>&1
wc -l gpurun_out/r2_p25_launches.csv
