timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vit_acs_kernel" --launch-skip 4 -c 1 -o gpurun_out/prof_v21_acs -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v21_acs.log 2>&1
tail -2 gpurun_out/ncu_v21_acs.log | cut -c1-200
