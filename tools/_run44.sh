timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_v34.log
BENCH_VERBOSE=1 timeout 600 python bench.py 2>gpurun_out/bench_v34_err.log > gpurun_out/bench_rx_v34.json
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_rx_v34.json'))
print(d['value'], d['ms_per_step'], d['stage_ms'])
for r in d['roofline_other']: print(r['kernel'][:40], round(r['achieved']), round(r['frac'],3), round(r['avg_launch_ms'],4))
P
