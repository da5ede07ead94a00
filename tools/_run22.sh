# h16 (VIADDMNMX.U16x2) ACS schedule: parity, tuning sweep, bench
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_v21.log
SWEEP_REPS=5 timeout 600 python tools/sweep_acs.py 2>&1 | tail -16 | tee gpurun_out/sweep_h16.txt
SWEEP_ONLY="384,0,384;512,0,512;640,0,320;768,0,384" SWEEP_REPS=5 timeout 300 python tools/sweep_acs.py 2>&1 | tail -5 | tee -a gpurun_out/sweep_h16.txt
DVBT_B200_VIT_ACS=swar SWEEP_ONLY="384,0,384" SWEEP_REPS=5 timeout 300 python tools/sweep_acs.py 2>&1 | tail -2 | tee -a gpurun_out/sweep_h16.txt
timeout 600 python bench.py 2>gpurun_out/bench_v21_err.log | tee gpurun_out/bench_rx_v21.json | cut -c1-300
timeout 600 python bench.py --workload viterbi 2>>gpurun_out/bench_v21_err.log | tee gpurun_out/bench_viterbi_v21.json | cut -c1-300
