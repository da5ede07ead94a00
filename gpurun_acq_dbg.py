import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench
w = bench.RxWorkload(4)
w.setup_gpu(1)
for i in range(3): w.step_resident(i)
print({k:v for k,v in w.info.items()})
print('check', w.check())
