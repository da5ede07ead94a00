#!/bin/bash
# round 2, GPU pass 2: parity of the streaming / re-synchronising chain, the new bench line, ncu traffic of the ACS kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2_p02_pytest.log
( time BENCH_VERBOSE=1 timeout 1200 python bench.py 2>gpurun_out/r2_p02_bench_err.log > gpurun_out/r2_p02_bench.json ) 2>&1 | tail -4
tail -5 gpurun_out/r2_p02_bench_err.log | cut -c1-400
python - <<'P'
import json
try:
    d = json.load(open('gpurun_out/r2_p02_bench.json'))
    print("value", d['value'], "ms/step", d['ms_per_step'], "parity", d['parity_check'], d.get('parity_detail'), "e2e", d['e2e']['value'], "roof", d['e2e'].get('h2d_roof_gbs'))
    print("one at a time", d['one_capture_at_a_time'])
    print("roofline", json.dumps(d['roofline'])[:900])
    for k, v in d.get('per_config', {}).items():
        print(k, {kk: v.get(kk) for kk in ('value', 'parity_check', 'parity_detail', 'error', 'ms_per_capture_one_at_a_time', 'acs_kernel_ms')}, v.get('e2e'))
    vs = d.get('viterbi_sweep') or {}
    for c in vs.get('cases', []):
        print(c)
    print("robustness", json.dumps(d.get('robustness'))[:2500])
    print("drop_in", json.dumps(d.get('drop_in_blocks'))[:2000])
    print("sweep", json.dumps(d.get('in_flight_sweep'))[:800])
except Exception as e:
    print("bench line unreadable:", e)
P
BENCH_NO_CONFIGS=1 BENCH_NO_VITERBI_SWEEP=1 BENCH_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"vit_acs_kernel|rx_descr_plan|rx_descramble|demod_scan" --launch-skip 12 -c 4 -o gpurun_out/r2_p02_acs -f python bench.py --steps 2 --warmup 3 > gpurun_out/r2_p02_ncu.log 2>&1
tail -2 gpurun_out/r2_p02_ncu.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/r2_p02_acs.ncu-rep "ncu --set full, round-2 streaming build (tools/gpu/r2_pass02.sh)" > gpurun_out/r2_p02_acs_ncu_summary.txt 2>&1
grep -E "^==|gpu__time_duration|dram__bytes" gpurun_out/r2_p02_acs_ncu_summary.txt
