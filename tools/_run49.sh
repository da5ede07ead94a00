# round 2, first GPU pass (prepared at the end of round 1, whose GPU budget was spent before the h16b schedule existed):
#   gpurun --timeout 1500 -- 'bash tools/_run49.sh'
# 1. parity of everything, the opt-in h16b ACS schedule included (xpass = green)
timeout 900 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -40 | tee gpurun_out/pytest_v49.log
# 2. A/B of the ACS schedules inside the RX chain (quick legs: ms per capture, ACS kernel ms, parity against the source TS)
for v in h16 h16b; do
  echo "ACS schedule $v"
  DVBT_B200_VIT_ACS=$v BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "bench quick"
done
DVBT_B200_VIT_ACS=h16b DVBT_B200_VIT_TPSM=512 DVBT_B200_VIT_BD=512 BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "bench quick"
# 3. the bench line (acs_variants and in_flight_sweep ride in it)
BENCH_VERBOSE=1 timeout 900 python bench.py 2>gpurun_out/bench_v49_err.log > gpurun_out/bench_rx_v49.json
python - <<'P'
import json
d = json.load(open('gpurun_out/bench_rx_v49.json'))
print(d['value'], d['ms_per_step'], d['parity_check'], d['e2e']['value'])
print(json.dumps(d.get('acs_variants'), indent=1)[:1500])
print(json.dumps(d.get('in_flight_sweep')))
P
# 4. ncu of the h16b ACS kernel inside the RX step (compare with profiles/r01_rx_acs_h16_v26_ncu_summary.txt)
DVBT_B200_VIT_ACS=h16b BENCH_NO_ACS_AB=1 BENCH_NO_SWEEP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vit_acs_kernel" --launch-skip 4 -c 1 -o gpurun_out/prof_v49_acs_h16b -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v49_acs.log 2>&1
tail -2 gpurun_out/ncu_v49_acs.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/prof_v49_acs_h16b.ncu-rep "ncu --set full of vit_acs_kernel<gring, h16b> inside the RX step (tools/_run49.sh)" > gpurun_out/rx_acs_h16b_v49_ncu_summary.txt 2>&1
head -30 gpurun_out/rx_acs_h16b_v49_ncu_summary.txt
