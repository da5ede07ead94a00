python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/sweep_acs.py 2>&1 | tail -14
for cfg in "256,0,128" "384,0,128"; do
  SWEEP_ONLY="$cfg" SWEEP_REPS=4 ncu --set full --clock-control none -k regex:vit_acs_kernel --launch-skip 2 -c 1 -o gpurun_out/prof_acs_${cfg//,/_} -f python tools/sweep_acs.py > gpurun_out/ncu_acs_${cfg//,/_}.log 2>&1
  tail -2 gpurun_out/ncu_acs_${cfg//,/_}.log | cut -c1-200
done
python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_rx_v11.json | cut -c1-300
