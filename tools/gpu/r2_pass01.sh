#!/bin/bash
# round 2, GPU pass 1 (gpurun --timeout 1500 -- 'bash tools/gpu/r2_pass01.sh'): baseline of the round-1 build
#  1. parity suite  2. A/B of the ACS schedules in the chain  3. bare H2D roof  4. ncu --set full of the five
#  heaviest kernels with source (read here with ncu -i ... --page source)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -15 | tee gpurun_out/r2_p01_pytest.log
for v in h16 h16b; do
  echo "ACS schedule $v"
  DVBT_B200_VIT_ACS=$v BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "bench quick" | tee -a gpurun_out/r2_p01_acs_ab.log
done
DVBT_B200_VIT_ACS=h16b DVBT_B200_VIT_TPSM=512 DVBT_B200_VIT_BD=512 BENCH_QUICK=1 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep "bench quick" | tee -a gpurun_out/r2_p01_acs_ab.log
timeout 120 python tools/gpu/h2d_roof.py 2>&1 | tee gpurun_out/r2_p01_h2d_roof.log
BENCH_NO_ACS_AB=1 BENCH_NO_SWEEP=1 BENCH_QUICK=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"vit_acs_kernel|demod_equalise|resample_multi|rx_inner_codes|acq_fftd|demod_stage1|acq_pass2|rs_decode" --launch-skip 24 -c 8 \
  -o gpurun_out/r2_p01_top8 -f python bench.py --steps 2 --warmup 3 > gpurun_out/r2_p01_ncu.log 2>&1
tail -3 gpurun_out/r2_p01_ncu.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/r2_p01_top8.ncu-rep "ncu --set full, round-1 build, top kernels of the RX step (tools/gpu/r2_pass01.sh)" > gpurun_out/r2_p01_top8_ncu_summary.txt 2>&1
head -50 gpurun_out/r2_p01_top8_ncu_summary.txt
nvidia-smi topo -m > gpurun_out/r2_p01_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r2_p01_topo.txt; numactl -H >> gpurun_out/r2_p01_topo.txt 2>&1
