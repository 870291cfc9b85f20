import sys,time,os; sys.path.insert(0,'.')
import numpy as np, torch
import dplug_b200 as b2
from dplug_b200 import synth
W,H=3840,2160
d,m,z=[torch.from_numpy(a).pin_memory().numpy() for a in synth.frame(W,H)]
wfb=torch.empty((H,W,4),dtype=torch.uint8).pin_memory().numpy()
c=b2.PBRCompositor(0); c.resizeBuffers(W,H); c.setSkybox(synth.skybox())
tiles=b2.BoxArray(b2.tileAreas([(0,0,W,H)],64,64)); whole=b2.BoxArray([(0,0,W,H)])
for _ in range(3): c.compositeTile(wfb,tiles,d,m,z,updateRects=whole)
t0=time.perf_counter()
for _ in range(20): c.compositeTile(wfb,tiles,d,m,z,updateRects=whole)
dt=(time.perf_counter()-t0)/20
print(os.environ.get('B2_PIPE_CHUNKS'),'%.3f ms  %.0f Mpix/s'%(dt*1e3,W*H/dt/1e6))
