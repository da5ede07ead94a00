timeout 200 python bench.py 2>gpurun_out/bench_v37_err.log > gpurun_out/bench_rx_v37.json
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_rx_v37.json'))
print(d['value'], d['ms_per_step'], d['parity_check'], d['e2e']['value'], {k:round(v,3) for k,v in d['stage_ms'].items()})
for r in d['roofline_other']: print(r['kernel'][:40], round(r['achieved']), round(r['frac'],3), round(r['avg_launch_ms'],4))
P
