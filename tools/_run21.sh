# DPX instruction rates + ncu --set full of the HBM-side kernels of the RX step
./tools/ubench/dpx_rate | tee gpurun_out/dpx_rate.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"acq_fftd_kernel|demod_equalise_kernel|resample_quad_kernel|rx_inner_codes_kernel|demod_scan_kernel|acq_compose_kernel" --launch-skip 18 -c 6 -o gpurun_out/prof_v20_side -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v20_side.log 2>&1
tail -3 gpurun_out/ncu_v20_side.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
