import os, sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import gr_dvbt_b200 as g
from oracle import port as O
for rate, m in [(4,6),(0,4),(2,4)]:
    k,n = O.RATE_KN[rate]
    nblocks = 20000 if rate else 80000
    data = np.random.default_rng(1).integers(0,256,nblocks*96*k,dtype=np.uint8)
    rx = O.conv_encode(data, m, rate)
    d_in = torch.from_numpy(rx).cuda(); nbt = len(data); d_out = torch.zeros(nbt, dtype=torch.uint8, device='cuda')
    for lanes in ("1","2"):
        os.environ["DVBT_B200_VIT_LANES"] = lanes
        dec = g.viterbi_decoder({2:0,4:1,6:2}[m], g.NH, rate)
        for it in range(4):
            nout = dec.decode_dev(d_in.data_ptr(), len(rx), len(rx), 1, d_out.data_ptr(), nbt)
            st = dec.last_stats()
        ok = np.array_equal(d_out[:nout].cpu().numpy(), data[:nout])
        print(f"rate {rate} m {m} lanes {lanes}: acs {st['acs_kernel_ms']:.3f} ms -> {nout*8/st['acs_kernel_ms']/1e6:.1f} Gbit/s; chunks {st['chunks']} repaired {st['repaired']} ok={ok}")
