#!/bin/bash
# round 2, GPU pass 3a (1 GPU): parity incl. the guard-interval cases, Viterbi sweep with the parallel repair rounds
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2_p03_pytest.log
for r in 0 2; do
DVBT_B200_VIT_REPAIR_ROUNDS=$r python - <<'P' 2>&1 | tee -a gpurun_out/r2_p03_repair_rounds.log
import os, sys, json
sys.path.insert(0, '.')
import torch, bench, numpy as np
import gr_dvbt_b200 as g
g.capi.check(g.capi.lib().dvbt_b200_set_device(0))
res = bench.viterbi_sweep(g, torch, None, lambda: torch.cuda.synchronize())
print("repair rounds", os.environ.get("DVBT_B200_VIT_REPAIR_ROUNDS"))
for c in res["cases"]:
    if c["channel_ber"] >= 1e-2 and c["m"] == 6:
        print({k: c[k] for k in ("rate", "channel_ber", "ms_per_decode", "acs_kernel_ms", "repaired_chunks", "parity")})
P
done
BENCH_NO_CONFIGS=1 BENCH_NO_VITERBI_SWEEP=1 BENCH_NO_DROPIN=1 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'one', d['one_capture_at_a_time']['ms_per_capture'], 'parity', d['parity_check'])
print(json.dumps(d['robustness'])[:1500])"
