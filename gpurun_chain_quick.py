import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import gr_dvbt_b200 as g
from oracle import refchain as R
from dvbt_testlib import tx_frequency_domain, channel, ofdm_modulate
con, cr, tm = R.QAM64, R.C7_8, R.T2k
tx = tx_frequency_domain(con, cr, tm, 1088, 3)
X0 = tx['X'][:1088]
X = np.tile(channel(X0), (16,1)); nsym = X.shape[0]
rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
dX = torch.from_numpy(X).cuda(); dts = torch.zeros(nsym*1512, dtype=torch.uint8, device='cuda')
for it in range(3):
    torch.cuda.synchronize(); t=time.time()
    n = rx.run_freq_dev(dX.data_ptr(), nsym, dts.data_ptr(), dts.numel())
    torch.cuda.synchronize(); dt=time.time()-t
info = rx.info()
print('FREQ nsym', nsym, 'ts bytes', n, 'wall %.2f ms'%(dt*1e3), {k:(round(v,3) if isinstance(v,float) else v) for k,v in info.items()})
print('symbols/s %.0f  => equivalent 10Msps-domain Msamples/s: %.0f'%(nsym/dt, nsym*2310/dt/1e6))
x = ofdm_modulate(np.tile(X0,(16,1)), tm, offset=777)
dx = torch.from_numpy(x).cuda()
for it in range(3):
    torch.cuda.synchronize(); t=time.time()
    n = rx.run_baseband_dev(dx.data_ptr(), len(x), dts.data_ptr(), dts.numel())
    torch.cuda.synchronize(); dt=time.time()-t
info = rx.info()
print('BASEBAND samples', len(x), 'ts bytes', n, 'wall %.2f ms'%(dt*1e3), {k:(round(v,3) if isinstance(v,float) else v) for k,v in info.items()})
print('Msamples/s (64/7 domain) %.0f ; 10Msps-domain equivalent %.0f'%(len(x)/dt/1e6, len(x)*70/64/dt/1e6))
ts = dts[:n].cpu().numpy(); src = tx['ts']
print('ts ok prefix:', np.array_equal(ts[:188*1000], src[1328*188:1328*188+188*1000]))
