BENCH_VERBOSE=1 timeout 600 python bench.py 2>gpurun_out/bench_v36_err.log > gpurun_out/bench_rx_v36.json
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_rx_v36.json'))
print(d['value'], d['ms_per_step'], d['parity_check'], d['e2e']['value'], {k:round(v,3) for k,v in d['stage_ms'].items()})
for r in d['roofline_other']: print(r['kernel'][:40], round(r['achieved']), round(r['frac'],3), round(r['avg_launch_ms'],4))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/rx_v36_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_v36.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"demod_equalise_kernel|acq_fftd_kernel|resample_multi_kernel|acq_pass2_kernel" --launch-skip 12 -c 4 -o gpurun_out/prof_v36_side -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_v36_side.log 2>&1
tail -1 gpurun_out/ncu_v36_side.log | cut -c1-150
