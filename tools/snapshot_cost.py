import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
n=1<<20
cfg=cb.default_config(8)
amp,freq,phase,pose7,twist6=wl.c3_instances(n,1)
for every in (0, 100, 10, 1000):
    g=cb.CdprBatch(cfg,n)
    g.set_platform_state(pose7,twist6); g.set_sine_cmd(amp,freq,phase)
    if every:
        buf=torch.empty((1000//every,13,n),dtype=torch.float64,device="cuda"); torch.cuda.synchronize()
    ms=[]
    for _ in range(3):
        if every: g.set_snapshots(every, buf.data_ptr(), buf.shape[0])
        g.step(1000); ms.append(g.last_kernel_ms)
    print("snapshot every", every, "kernel ms", ms)
    g.close()
