timeout 900 ncu --set full --clock-control none --import-source on -k regex:"demod_equalise_kernel|rx_inner_codes_kernel|acq_pass2_kernel" --launch-skip 6 -c 3 -o gpurun_out/prof_v34_side -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_v34_side.log 2>&1
tail -2 gpurun_out/ncu_v34_side.log | cut -c1-200
