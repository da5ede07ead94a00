for sd in 1 2 3 4; do
DVBT_B200_ACQ_TRACE=2 BENCH_SEED=$sd timeout 600 python bench.py --steps 1 --warmup 3 2>gpurun_out/bench_trace3_seed${sd}_err.log | cut -c1-100
grep "acq best\|acq batch" gpurun_out/bench_trace3_seed${sd}_err.log | head -3 | cut -c1-330
done
# ncu full capture of the current ACS kernel inside the RX step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vit_acs_kernel" --launch-skip 4 -c 1 -o gpurun_out/prof_v26_acs -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v26_acs.log 2>&1
tail -2 gpurun_out/ncu_v26_acs.log | cut -c1-200
