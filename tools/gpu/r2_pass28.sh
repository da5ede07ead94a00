#!/bin/bash
# round 2, GPU pass 28 (1 GPU): the capture of rank 6 at configs[3] (seed 107) takes 15.9 ms instead of 2 - which stage?
# + the soft chain with the padded inner kernel
mkdir -p gpurun_out
DVBT_B200_ACQ_TRACE=1 python - <<'P' 2>&1 | tail -40 | cut -c1-400 | tee gpurun_out/r2_p28_seed107.log
import time, json, torch
import bench
for seed in (101, 107):
    w = bench.RxWorkload(0, "configs[3]")
    w.setup_gpu(seed=seed)
    for i in range(3): w.step_resident(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(5): w.step_resident(i)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    print("seed", seed, "ms/capture", round(ms, 3), "check", w.check(), json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in w.info.items()}))
P
timeout 600 python -m pytest tests/test_soft_chain_gpu.py tests/test_soft_decision_gpu.py tests/test_stream_edge_cases_gpu.py -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/r2_p28_pytest.log
BENCH_NO_CONFIGS=1 BENCH_NO_VITERBI_SWEEP=1 BENCH_NO_DROPIN=1 BENCH_NO_TX=1 BENCH_NO_SWEEP=1 timeout 900 python bench.py --steps 8 2>/dev/null > gpurun_out/r2_p28_bench.json
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_p28_bench.json'))
for r in d['robustness'].get('soft_decision', []): print("robust", {k: r[k] for k in ('snr_db', 'decisions', 'ms_per_capture', 'packets_equal_to_source')}, r.get('stage_ms'))
P
